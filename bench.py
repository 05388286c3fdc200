#!/usr/bin/env python
"""bench.py — headline benchmark of the modal hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[4], the polyphonic resonator bank — 1024 voices x 500 modes x 10 s at
48 kHz, one strike per voice at frame 0 plus Poisson re-strikes — the one metric BASELINE.json quotes at 1/2/4/8 B200.
A "step" is one offline render of the whole 10 s timeline. Voices are sharded over the ranks (strong scaling: the total
stays 1024 voices) and the per-rank mono mixes are summed with one NCCL all-reduce over NVLink.

`value`  = mode-samples/s with the bank resident in HBM, timed on the device (CUDA events, max over ranks).
`e2e`    = the same metric through the C ABI with the strike timeline in HOST memory and the mix read back to HOST.
`roofline` = FP32 issue ceiling (the resonator is FP32-ALU bound); `roofline_hbm` = mandatory bytes.
`cpu_baseline` = the UNMODIFIED reference RenderModal (oracle/_ref) on the host cores, on a bounded sample.
`--impl reference` runs only that CPU arm and prints it in the same format.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOICES, MODES, SECONDS, RATE, BLOCK = 1024, 500, 10.0, 48000.0, 512
METRIC, UNIT = "resonator mode-samples/s", "mode-samples/s"
OPS_PER_MODE_SAMPLE = 2.75  # FP32 lane-operations per mode-sample of the K=4 kernel: (2K+3)/K
REF_OPS_PER_MODE_SAMPLE = 7.0  # the reference's loop, SURVEY.md §8(d)


def env_int(name, default):
    return int(os.environ.get(name, default))


def tensor_traffic(which):
    """DRAM bytes per launch of the tensor-form kernels from the committed ncu --set full capture, when there is one."""
    try:
        with open(os.path.join(ROOT, "profiles", "resonator_traffic.json")) as f:
            return json.load(f).get(f"{which}_dram_bytes_per_launch")
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """One `nvidia-smi -lms` process sampling clocks and throttle reasons while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            self.rows = [[x.strip() for x in line.split(",")] for line in out.splitlines() if line.count(",") >= 6]

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        # "Under load" = samples drawing more than the idle floor; the median SM clock over those.
        power = [float(r[2]) for r in self.rows]
        loaded = [r for r, p in zip(self.rows, power) if p >= 0.5 * max(power)] or self.rows
        sm = sorted(float(r[0]) for r in loaded)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "power_w_max": max(power), "samples": len(self.rows), "samples_under_load": len(loaded), "reasons": reasons}


def cpu_reference(steps, warmup, voices=256, seconds=1.0, quiet=False):
    """The reference's own RenderModal (oracle/_ref/libme_ref_audio.so, built from /root/reference sources) on the host
    cores, RenderThreads = min(16, cores) (its cap, src/audio/AudioSystem.cpp:60), 512-frame blocks."""
    from mesheditor_b200 import workloads as wl
    from oracle import resonator as orc

    kind = "reference" if orc.have_ref() else "port"
    cores = os.cpu_count() or 1
    threads = min(16, cores) if kind == "reference" else 1
    modes = wl.c5_modes(MODES)
    blocks = int(seconds * RATE) // BLOCK
    events, ev_frames, _ = wl.c5_timeline(voices, blocks * BLOCK)
    times, live = [], None
    for it in range(warmup + steps):
        scene = (orc.RefScene if kind == "reference" else orc.PortBank)(RATE, threads)
        for _ in range(voices):
            scene.add_modes(modes)
        scene.install()
        out = np.zeros(blocks * BLOCK, np.float32)
        k = 0
        t0 = time.perf_counter()
        for b in range(blocks):
            while k < len(events) and ev_frames[k] == b * BLOCK:
                v, impulse, ex = events[k]
                scene.enqueue(orc.impact_event(v, impulse, ex))
                k += 1
            scene.render(out[b * BLOCK:(b + 1) * BLOCK])
        dt = time.perf_counter() - t0
        live = scene.object_column("LiveModeCount")
        if it >= warmup:
            times.append(dt)
    mode_samples = voices * MODES * blocks * BLOCK
    ms = 1e3 * sum(times) / len(times)
    return {
        "value": mode_samples / (ms * 1e-3), "unit": UNIT, "cores": threads, "kind": kind, "ms_per_step": ms,
        "sample": f"{voices} voices x {MODES} modes x {blocks * BLOCK} frames ({blocks} blocks of {BLOCK}), same strike recipe; LiveModeCount {int(live.min())}..{int(live.max())} of {MODES} (no culling)",
    }


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    base = cpu_reference(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus), "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {
        "workload": f"BASELINE.json configs[4]: polyphonic resonator bank {VOICES} voices x {MODES} modes x {SECONDS:g} s at {RATE:g} Hz, strike per voice at frame 0 + 2 Hz Poisson re-strikes (MT19937 12345), 512-frame blocks",
        "voices": VOICES, "modes_per_voice": MODES, "frames": int(SECONDS * RATE), "sample_rate": RATE, "block_frames": BLOCK,
        "parallelism": f"voices sharded over {n_gpus} GPU(s), NCCL all-reduce of the mono mix" if n_gpus > 1 else "single GPU",
        "l2_policy": "working set > L2: one step writes and reads 8 GB of block-start states (tensor-core form) or 3.9 GB of per-warp partial mixes (sample loop); nothing is reused across steps",
    }


def run_ours(args):
    import torch
    import torch.distributed as dist

    global VOICES
    if args.voices:
        VOICES = args.voices

    from mesheditor_b200 import ModalBank, build, measure_fp32_fma_rate
    from mesheditor_b200 import workloads as wl

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    frames = int(SECONDS * RATE)
    lo, hi = wl.shard_voices(VOICES, world, rank)
    modes = wl.c5_modes(MODES)
    all_events, all_frames, all_voice = wl.c5_timeline(VOICES, frames)
    mine = (all_voice >= lo) & (all_voice < hi)
    events = [wl.impact(v - lo, impulse, ex) for (v, impulse, ex), keep in zip(all_events, mine) if keep]
    ev_frames = all_frames[mine]
    events = ModalBank.pack_events(events, ev_frames)  # contiguous MeModalEvent[] + frames, built once

    bank = ModalBank(RATE, local)
    for _ in range(hi - lo):
        bank.add_modes(modes)
    bank.install(0)
    bank.set_render_path({"auto": 0, "loop": 1, "tensor": 2}[args.render_path])

    out = torch.zeros(frames, dtype=torch.float32, device="cuda")
    host_out = torch.zeros(frames, dtype=torch.float32).pin_memory()
    # One explicit stream carries the library's kernels, the NCCL all-reduce and the timing events.
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    launches = [0]

    def step_device():
        bank.render_offline_device(events, ev_frames, frames, BLOCK, out.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(out)

    def step_e2e():
        # Strike timeline from host memory in, final mix back in host memory (rank 0 keeps it).
        bank.render_offline_device(events, ev_frames, frames, BLOCK, out.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(out)
        host_out.copy_(out, non_blocking=True)
        stream.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Each step re-renders the same 10 s from a silent bank, so every step does identical work.
    def reset():
        bank.install(0)

    for _ in range(max(args.warmup, 3)):
        reset()
        step_device()
    barrier()

    kernel_ms, walk_ms, mix_ms, stats = [], [], [], None
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    elapsed_ms = 0.0
    with ClockSampler(local) as clocks:
        for _ in range(args.steps):
            reset()
            barrier()
            start.record(stream)
            step_device()
            stop.record(stream)
            barrier()
            elapsed_ms += start.elapsed_time(stop)
            stats = bank.stats()
            if os.environ.get("ME_BENCH_DEBUG"):
                print(f"[bench] step {start.elapsed_time(stop):.2f} ms, library total {stats['total_device_ms']:.2f} ms, walk {stats['walk_kernel_ms']:.2f}, mix {stats['tensor_mix_kernel_ms']:.2f}, host plan {stats['host_plan_ms']:.2f}, pulses {stats['pulse_kernels_ms']:.2f}, stage {stats['resonator_kernel_ms']:.2f}", file=sys.stderr)
            kernel_ms.append(stats["resonator_kernel_ms"])
            walk_ms.append(stats["walk_kernel_ms"])
            mix_ms.append(stats["tensor_mix_kernel_ms"])
            launches[0] += stats["kernel_launches"]
    # e2e: wall clock around the host-buffer path.
    e2e_s = 0.0
    for _ in range(args.steps):
        reset()
        barrier()
        t0 = time.perf_counter()
        step_e2e()
        e2e_s += time.perf_counter() - t0
        launches[0] += bank.stats()["kernel_launches"]
    barrier()

    live = [bank.object_status(v)["LiveModeCount"] for v in range(hi - lo)]
    fallbacks = stats["scan_fallbacks"]
    t = torch.tensor([elapsed_ms, e2e_s, float(launches[0]), float(min(live)) - 1e6 * fallbacks], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tmin = t.clone()
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        elapsed_ms, e2e_s, total_launches, min_live = float(tmax[0]), float(tmax[1]), int(tsum[2]), int(tmin[3])
    else:
        total_launches, min_live = launches[0], int(min(live))

    if rank == 0:
        total_mode_samples = VOICES * MODES * frames
        ms_per_step = elapsed_ms / args.steps
        value = total_mode_samples / (ms_per_step * 1e-3)
        e2e_value = total_mode_samples / (e2e_s / args.steps)
        pk, pk_kind = peaks()
        fma_peak = measure_fp32_fma_rate(local, 0, 10)       # scalar FFMA, the nominal FP32 issue ceiling
        fma_peak_fresh = measure_fp32_fma_rate(local, 5, 10)  # FFMA2 with three fresh register pairs (register-file bound)
        k_ms = sum(kernel_ms) / len(kernel_ms)
        rank_mode_samples = (hi - lo) * MODES * frames
        achieved = OPS_PER_MODE_SAMPLE * rank_mode_samples / (k_ms * 1e-3)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "resonator_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        # Mandatory bytes of one launch on this rank: 32 B per mode of state/coefficients + the per-warp
        # partial rows written once (4 B per warp-sample) — the latter is this design's own traffic, counted as such.
        mandatory = (hi - lo) * (32 * MODES) + 4 * frames
        base = cpu_reference(1, 0) if not args.no_cpu_baseline else None
        tensor_form = stats["tensor_windows"] > 0
        if tensor_form:
            # Tensor-core form. Dominant kernels: the tcgen05 mix (3xTF32 GEMM) and the state walk that
            # feeds it. Issued flops and written bytes follow from the padded layout: 4 voices of 63 chunks per 256-chunk
            # group, 4096 reduction elements per group, 128 time blocks of 256 frames per tile.
            chunks = -(-MODES // 8)
            groups = -(-(hi - lo) // (256 // chunks))
            tiles = -(-frames // 32768)
            product_flops = 2.0 * 256 * 128 * 4096 * groups * tiles  # one of the three products of the 3xTF32 split
            # head x head runs as kind::tf32, the two cross products as kind::f16 on BF16 copies at twice that rate: the
            # tensor-pipe time at peak is (1 + 2/2) products at the TF32 rate, so "achieved" is quoted in TF32-equivalent flop/s.
            issued_flops = 3 * product_flops
            tf32_equivalent = 2 * product_flops
            walk_bytes = 4096 * 4.0 * groups * -(-frames // 256)
            m_ms, w_ms = sum(mix_ms) / len(mix_ms), sum(walk_ms) / len(walk_ms)
            tf32_peak = pk.get("bf16_tflops", 2250.0) / 2
            stages = groups * 256 * tiles
            smem_bytes = stages * 106496.0  # per 16-element stage: 40 KB of TMA writes, 16 KB of splitter traffic, 48 KB of MMA operand reads
            roofline = {
                "bound": "tensor", "kernel": "TensorMixKernel<128,4> (tcgen05.mma: head x head kind::tf32, cross products kind::f16 on BF16 copies; FP32 register folds)",
                "achieved": tf32_equivalent / (m_ms * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TF32-equivalent TFLOP/s", "frac": tf32_equivalent / (m_ms * 1e-3) / 1e12 / tf32_peak,
                "traffic": tensor_traffic("mix"), "kernel_ms_per_launch": m_ms, "launches_per_step": stats["tensor_windows"],
                "issued_flops_per_step": issued_flops, "issued_tflops": issued_flops / (m_ms * 1e-3) / 1e12, "share_of_step": m_ms / ms_per_step,
                "peak_source": ("half of MEASURED_PEAKS.json bf16_tflops (TF32 runs at half the bf16 rate; nominal 1125)" if "bf16_tflops" in pk else "nominal dense TF32 1125 TFLOP/s"),
                "reference_fma_equivalent": REF_OPS_PER_MODE_SAMPLE * rank_mode_samples / (m_ms * 1e-3) / 1e12,
                "shared_memory": {"bytes_per_launch": smem_bytes, "achieved_bytes_per_clk_per_sm": smem_bytes / (m_ms * 1e-3) / (148 * (pk.get("sm_max_mhz", 1965.0) * 1e6)), "peak_bytes_per_clk_per_sm": 128,
                                  "note": "what actually bounds the kernel: SS-mode MMAs read both operands from shared memory (DESIGN.md 5.2, profiles/r01_resonator.md)"},
            }
            roofline_extra = {
                "roofline_walk": {"bound": "hbm", "kernel": "ResonatorKernel<1,2,true> (state walk: c^256 steps, FP32 state rows)", "achieved": walk_bytes / (w_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                  "frac": walk_bytes / (w_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": tensor_traffic("walk"), "kernel_ms_per_launch": w_ms, "algorithmic_bytes": walk_bytes, "share_of_step": w_ms / ms_per_step, "peak_source": pk_kind},
                "fp32_pipe_equivalent": {"note": "the same mode-samples per second on the FP32 pipe would need this multiple of the measured scalar-FFMA ceiling (reference loop: 7 lane-ops per mode-sample; this repo's sample loop: 2.75)",
                                         "reference_loop": REF_OPS_PER_MODE_SAMPLE * rank_mode_samples / (k_ms * 1e-3) / fma_peak, "sample_loop": achieved / fma_peak, "ffma_peak_tlane_ops": fma_peak / 1e12},
            }
        else:
            roofline = {
                "bound": "fp32_fma", "kernel": "ResonatorKernel<4,2>", "achieved": achieved / 1e12, "peak": fma_peak / 1e12, "unit": "TFMA-lane-op/s",
                "frac": achieved / fma_peak, "traffic": traffic, "kernel_ms_per_launch": k_ms,
                "ops_per_mode_sample": OPS_PER_MODE_SAMPLE, "peak_source": "FFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP32 figure)",
                "peak_three_fresh_operands": fma_peak_fresh / 1e12, "frac_of_register_file_bound": achieved / fma_peak_fresh,
                "reference_op_equivalent_frac": REF_OPS_PER_MODE_SAMPLE * rank_mode_samples / (k_ms * 1e-3) / fma_peak,
            }
            roofline_extra = {}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (3xTF32 products, FP32 accumulation)" if tensor_form else "f32", "data": "synthetic",
            "config": dict(workload_config(world), live_mode_count_min=min_live, culling_triggered=min_live < MODES, time_segments=stats["time_segments"],
                           render_path="tensor-core form: state walk + tcgen05 mix" if tensor_form else "FP32 sample loop", partial_rows=stats["partial_rows"]),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(stats["h2d_bytes"]), "d2h_bytes_per_step": frames * 4, "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": total_launches,
            "roofline": roofline,
            **roofline_extra,
            "roofline_hbm": {"bound": "hbm", "achieved": mandatory / (k_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": mandatory / (k_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "peak_source": pk_kind,
                             "note": "mandatory bytes of the reference formulation only (SURVEY.md F9)"},
            "clocks": clocks.summary(),
        }
        if base:
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ analysis workloads
SOLVE_METRIC, SOLVE_UNIT = "modal solve seconds per mesh", "s/mesh"
C3_EDGE, C3_MODES, C3_ORDER = 55, 200, 1  # BASELINE.json configs[2]: 55^3 Kuhn block = 998,250 tets, lowest 200 modes, P1
CPU_SAMPLE_EDGE = 16                       # the CPU arm's bounded sample: 16^3 cells = 24,576 tets (~20-30 s of host work), same material / modes / element order


def solve_config():
    return {
        "workload": f"BASELINE.json configs[2]: synthetic {C3_EDGE}^3-cell Kuhn block = {6 * C3_EDGE ** 3:,} tets (0.3 m steel cube), linear (P1) elements (the reference's P2 does not fit one GPU at this size, SURVEY.md F1), "
                    f"lowest {C3_MODES} modes (NumFemModes {C3_MODES + 15}, ncv {C3_MODES + 35}), FP64 shift-invert Lanczos, 10 excitation points",
        "tets": 6 * C3_EDGE ** 3, "element_order": C3_ORDER, "num_modes": C3_MODES,
        "l2_policy": "working set > L2: the factor is 7.3 GB and the Lanczos basis 1 GB, both streamed from HBM every step",
    }


def cpu_solve_reference(edge=CPU_SAMPLE_EDGE, modes=C3_MODES, order=C3_ORDER):
    """The oracle's restatement of the reference path (numpy assembly + ARPACK shift-invert over SuperLU; the reference binary
    itself needs Eigen + Apple Accelerate and cannot be built here) on a bounded sample of the workload."""
    from mesheditor_b200 import workloads as wl
    from oracle import modal as om

    points, tets = wl.kuhn_block(edge, edge, edge, (0.3, 0.3, 0.3))
    cfg = om.SolverConfig(num_modes=modes, num_fem_modes=modes + 15, max_mode_freq=1e9)
    t0 = time.perf_counter()
    r = om.mesh2modes(points, tets, om.MATERIALS["Steel"], wl.bench_excitations(points), config=cfg, order=order, tol=1e-8)
    dt = time.perf_counter() - t0
    return {"value": dt, "unit": SOLVE_UNIT, "cores": os.cpu_count() or 1, "kind": "port", "tets": len(tets), "dofs": r["dofs"],
            "sample": f"{edge}^3-cell Kuhn block = {len(tets):,} tets ({len(tets) / (6 * C3_EDGE ** 3):.1%} of the workload's tets), P{order}, {modes} modes, oracle/modal.py (numpy assembly + scipy ARPACK/SuperLU, BLAS threads on all host cores): "
                      f"{dt:.1f} s for this sample; the full 998,250-tet mesh does not finish on the host in the bench's time budget"}


def solve_once(points, tets, ex, cfg):
    from mesheditor_b200 import mesh2modes

    t0 = time.perf_counter()
    r = mesh2modes(points, tets, "Steel", ex, config=cfg)
    return time.perf_counter() - t0, r


def run_solve(args):
    """`--workload solve`: one 1M-tet mesh on one GPU (a single eigensolve does not shard: with --gpus N every rank solves a
    replica and the slowest is reported)."""
    import torch
    import torch.distributed as dist

    from mesheditor_b200 import Factor, FemSystem, build, measure_fp64_rate, solver_config
    from mesheditor_b200 import workloads as wl

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    points, tets = wl.kuhn_block(C3_EDGE, C3_EDGE, C3_EDGE, (0.3, 0.3, 0.3))
    ex = wl.bench_excitations(points)
    cfg = solver_config(num_modes=C3_MODES, element_order=C3_ORDER, max_mode_freq=1e9, device=local)
    for _ in range(max(1, min(args.warmup, 3))):
        solve_once(points, tets, ex, cfg)
    times, profiles = [], []
    with ClockSampler(local) as clocks:
        for _ in range(args.steps):
            torch.cuda.synchronize()
            dt, r = solve_once(points, tets, ex, cfg)  # host mesh in, host modal model out: this IS the end-to-end call
            assert r.status == 0 and len(r.freqs) == C3_MODES
            times.append(dt)
            profiles.append(r.profile)
    t = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(t[0])
        prof = {k: (sum(p[k] for p in profiles) / len(profiles) if isinstance(profiles[0][k], float) else profiles[-1][k]) for k in profiles[0]}
        device_s = sec - prof["mass_props"]  # everything but the (host) lumped mass properties runs on, or waits for, the device
        pk, pk_kind = peaks()
        # Stage rooflines, measured in this run through the stage entry points of the C ABI.
        fem = FemSystem(points, tets, "Steel", C3_ORDER, local)
        i = fem.info
        x = np.random.default_rng(0).standard_normal(i["dofs"])
        fem.spmv("K", x, 50)
        spmv_bytes = 12 * 9 * i["node_blocks_full"] + 20 * i["dofs"] + 4
        spmv = {"kernel": "SpmvBsr3Kernel (y = K x)", "bound": "hbm", "achieved": spmv_bytes / (fem.last_spmv_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "ms": fem.last_spmv_ms, "algorithmic_bytes": spmv_bytes}
        spmv["frac"] = spmv["achieved"] / spmv["peak"]
        asm_bytes = 16 * len(tets) + 24 * len(points) + 12 * (i["nnz_stiffness"] + i["nnz_mass"]) + 8 * (i["dofs"] + 1)
        asm = {"kernel": "AssembleKernel", "bound": "hbm", "achieved": asm_bytes / (i["assemble_kernel_ms"] * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "ms": i["assemble_kernel_ms"], "algorithmic_bytes": asm_bytes}
        asm["frac"] = asm["achieved"] / asm["peak"]
        f = Factor(fem, -((2 * np.pi * 20.0) ** 2))
        fi = f.info
        dmma = measure_fp64_rate(local, 1, 3)
        factor = {"kernel": "SyrkScatterKernel + PanelTrsmKernel + FactorDiagKernel (numeric Cholesky)", "bound": "tensor", "achieved": fi["factor_flops"] / (fi["factor_device_ms"] * 1e-3) / 1e12, "peak": dmma / 1e12,
                  "unit": "TFLOP/s", "ms": fi["factor_device_ms"], "flops": fi["factor_flops"], "peak_source": "FP64 DMMA (mma.sync.m8n8k4.f64) issue-rate probe measured in this run; MEASURED_PEAKS.json has no FP64 figure"}
        factor["frac"] = factor["achieved"] / factor["peak"]
        b = np.random.default_rng(1).standard_normal(i["dofs"])
        for _ in range(3):
            f.solve(b)
        solve_ms = f.info["last_solve_device_ms"]
        sweep_bytes = 16 * fi["factor_nonzeros"] + 16 * i["dofs"]
        ops = prof["op_applications"]
        line = {
            "metric": SOLVE_METRIC, "value": sec, "unit": SOLVE_UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": 1e3 * sec, "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(solve_config(), parallelism="replicas only: one mesh's eigensolve runs on one GPU" if world > 1 else "single GPU"),
            "e2e": {"value": sec, "unit": SOLVE_UNIT, "h2d_bytes_per_step": int(points.nbytes + tets.nbytes + ex.nbytes), "d2h_bytes_per_step": int(8 * (C3_MODES + 15) + 4 * 3 * 10 * (C3_MODES + 15) + 4 * 3 * 10),
                    "note": "me_modal_solve takes the mesh from HOST memory and returns the modal model to HOST memory; value and e2e are the same call"},
            "device_seconds": device_s, "gpu_launches": int(sum(p["kernel_launches"] for p in profiles)),
            "profile": {k: prof[k] for k in ("mass_props", "assemble", "sample_excite", "factorize", "analyse", "iterate", "op_solve", "extract", "dofs", "stiffness_nonzeros", "op_applications", "restarts", "factor_nonzeros", "supernodes", "levels")},
            "roofline": {"bound": "hbm", "kernel": "triangular-solve sweep of the shift-invert operator (DiagSolve/PanelForward/PanelBackward, one CUDA-graph replay per operator application)",
                         "achieved": sweep_bytes / (solve_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": sweep_bytes / (solve_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None,
                         "ms_per_launch": solve_ms, "algorithmic_bytes": sweep_bytes, "share_of_step": prof["op_solve"] / sec, "applications_per_step": ops, "peak_source": pk_kind},
            "roofline_spmv": spmv, "roofline_assembly": asm, "roofline_factor": factor,
            "clocks": clocks.summary(),
        }
        if not args.no_cpu_baseline:
            base = cpu_solve_reference()
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_batch(args):
    """`--workload batch`: BASELINE.json configs[3], 64 Kuhn blocks of 10k..500k tets, dealt biggest-first to the least-loaded
    rank (the reference bench's biggest-file-first rule, tests/ModalSolverBench.cpp:475-478); no collective on the data path."""
    import torch
    import torch.distributed as dist

    from mesheditor_b200 import build, solver_config
    from mesheditor_b200 import workloads as wl

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = wl.config4_dims()
    owner = wl.lpt_assign([(6.0 * d ** 3) ** (4.0 / 3.0) for d in dims], world)
    mine = [d for d, o in zip(dims, owner) if o == rank]
    cfg = solver_config(num_modes=30, element_order=1, max_mode_freq=1e9, device=local)
    meshes = [wl.kuhn_block(d, d, d, (0.3, 0.3, 0.3)) for d in mine]

    def step():
        n = 0
        for points, tets in meshes:
            _, r = solve_once(points, tets, wl.bench_excitations(points), cfg)
            assert r.status == 0
            n += r.profile["kernel_launches"]
        return n

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(1, min(args.warmup, 3)) if args.warmup else 0):
        step()
    launches, total = 0, 0.0
    with ClockSampler(local) as clocks:
        for _ in range(args.steps):
            barrier()
            t0 = time.perf_counter()
            launches += step()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            total += float(dt[0])
    lt = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(lt)
    if rank == 0:
        sec = total / args.steps
        print(json.dumps({
            "metric": "batch modal solve meshes/s", "value": len(dims) / sec, "unit": "meshes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE.json configs[3]: 64 Kuhn-block meshes, {6 * dims[0] ** 3:,}..{6 * dims[-1] ** 3:,} tets (log-spaced), P1, lowest 30 modes each", "parallelism": f"meshes dealt biggest-first over {world} GPU(s), no collective"},
            "e2e": {"value": len(dims) / sec, "unit": "meshes/s", "h2d_bytes_per_step": int(sum(p.nbytes + t.nbytes for p, t in meshes)), "d2h_bytes_per_step": 0}, "gpu_launches": int(lt[0]), "clocks": clocks.summary()}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_solve_reference(args):
    if env_int("RANK", 0) != 0:
        return
    base = cpu_solve_reference()
    print(json.dumps({
        "impl": "reference", "metric": SOLVE_METRIC, "value": base["value"], "unit": SOLVE_UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": 1e3 * base["value"], "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": solve_config(), "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": SOLVE_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
        "note": "value is the time of the bounded SAMPLE (see cpu_baseline.sample), not of the full 998,250-tet mesh"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--voices", type=int, default=0, help="resonator: experiment with another voice count (the bench line then names it; default = BASELINE.json configs[4])")
    ap.add_argument("--render-path", default="auto", choices=["auto", "loop", "tensor"], help="resonator: kernels of the free-running bank (auto = tensor-core form for this workload)")
    ap.add_argument("--workload", default="resonator", choices=["resonator", "solve", "batch"],
                    help="resonator: configs[4] (default, the metric quoted at 1/2/4/8 GPUs); solve: configs[2], one 1M-tet mesh; batch: configs[3], 64 meshes sharded")
    args = ap.parse_args()
    if args.workload == "solve":
        if args.steps == 5:
            args.steps = 2
        run_solve_reference(args) if args.impl == "reference" else run_solve(args)
    elif args.workload == "batch":
        if args.steps == 5:
            args.steps = 1
        if args.impl == "reference":
            run_solve_reference(args)
        else:
            run_batch(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
