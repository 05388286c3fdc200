#!/bin/bash
# Round 2, visit L: timeline of one render (ME_RENDER_TRACE) and a pure-write bandwidth probe.
mkdir -p gpurun_out
python - <<'PY'
import torch
a=torch.empty(2<<30,dtype=torch.float32,device='cuda')  # 8 GB
for name,fn in (('fill',lambda: a.fill_(1.0)),('zero',lambda: a.zero_())):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    best=1e9
    for _ in range(5):
        s.record(); fn(); e.record(); torch.cuda.synchronize(); best=min(best,s.elapsed_time(e))
    print(name, 'write GB/s', a.numel()*4/best/1e6)
b=torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize()
s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
s.record(); b.copy_(a); e.record(); torch.cuda.synchronize(); print('copy GB/s (r+w)', 2*a.numel()*4/s.elapsed_time(e)/1e6)
s.record(); x=a.sum(); e.record(); torch.cuda.synchronize(); print('read GB/s', a.numel()*4/s.elapsed_time(e)/1e6)
PY
ME_RENDER_TRACE=1 ME_WALK_SUBWINDOW_TILES=3 ME_BENCH_DEBUG=1 timeout 300 python bench.py --workload resonator --steps 2 --warmup 3 --no-cpu-baseline --no-parity 2>&1 >/dev/null | tail -22
