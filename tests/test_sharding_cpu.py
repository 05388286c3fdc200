"""World-size-2 `gloo` coverage of the two sharded paths, on CPU.

The data path of each rank is the oracle here (this is a test; the product renders on the device): what is checked is
the HOST logic bench.py runs under torchrun — the voice split, the per-rank event filter and re-indexing, the all-reduce
of the mono mix against a single-rank render at the reference's own 1e-5*peak gate (tests/ModalRenderTest.cpp:48), and
the biggest-first deal of the mesh batch with its gather of per-mesh results."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

VOICES, MODES, BLOCKS, BLOCK = 6, 64, 24, 512


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _render_voices(lo, hi):
    """Oracle render of voices [lo, hi) of the shared timeline, voices re-indexed from 0 as bench.py does per rank."""
    from mesheditor_b200 import workloads as wl
    from oracle import resonator as orc

    orc.build_port()
    frames = BLOCKS * BLOCK
    events, ev_frames, ev_voice = wl.c5_timeline(VOICES, frames, restrike_hz=20.0)
    bank = (orc.RefScene if orc.have_ref() else orc.PortBank)(48000.0, 1)
    modes = wl.make_modes(MODES, 0.5)
    for _ in range(hi - lo):
        bank.add_modes(modes)
    bank.install()
    out = np.zeros(frames, np.float32)
    for b in range(BLOCKS):
        for (v, impulse, ex), f in zip(events, ev_frames):
            if f == b * BLOCK and lo <= v < hi:
                bank.enqueue(orc.impact_event(v - lo, impulse, ex))
        bank.render(out[b * BLOCK:(b + 1) * BLOCK])
    return out


def _mix_worker(rank, world, port, result):
    from mesheditor_b200 import workloads as wl

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = wl.shard_voices(VOICES, world, rank)
    mix = torch.from_numpy(_render_voices(lo, hi))
    dist.all_reduce(mix)  # the one collective of the synthesis path: sum of the per-rank mono mixes
    spans = [None] * world
    dist.all_gather_object(spans, (lo, hi))
    if rank == 0:
        result["mix"] = mix.numpy().copy()
        result["spans"] = spans
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mix_matches_single_rank_render(built_lib):
    world, port = 2, _free_port()
    with mp.Manager() as manager:
        result = manager.dict()
        mp.spawn(_mix_worker, args=(world, port, result), nprocs=world, join=True)
        mix, spans = np.array(result["mix"]), list(result["spans"])
    assert spans == [(0, 3), (3, 6)]  # contiguous, disjoint, covering
    whole = _render_voices(0, VOICES)
    peak = float(np.abs(whole).max())
    assert peak > 0
    assert float(np.abs(mix - whole).max()) <= 1e-5 * peak


def _batch_worker(rank, world, port, result):
    from mesheditor_b200 import workloads as wl

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dims = wl.config4_dims(12, 48, 3000)
    owner = wl.lpt_assign([(6.0 * d ** 3) ** (4.0 / 3.0) for d in dims], world)
    mine = [i for i in range(len(dims)) if owner[i] == rank]
    # stand-in for the per-mesh solve: the mesh's own size table, enough to check that every unit is done exactly once
    local = {i: (6 * dims[i] ** 3, (dims[i] + 1) ** 3) for i in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    work = torch.tensor([float(sum(t for t, _ in local.values()))], dtype=torch.float64)
    total = work.clone()
    dist.all_reduce(total)
    if rank == 0:
        result["gathered"] = gathered
        result["total"] = float(total[0])
        result["dims"] = dims
    dist.barrier()
    dist.destroy_process_group()


def test_batch_deal_covers_every_mesh_once(built_lib):
    world, port = 2, _free_port()
    with mp.Manager() as manager:
        result = manager.dict()
        mp.spawn(_batch_worker, args=(world, port, result), nprocs=world, join=True)
        gathered, total, dims = list(result["gathered"]), result["total"], list(result["dims"])
    seen = sorted(i for part in gathered for i in part)
    assert seen == list(range(len(dims)))  # no mesh dropped, none solved twice
    assert total == float(sum(6 * d ** 3 for d in dims))
    loads = [sum((6.0 * dims[i] ** 3) ** (4.0 / 3.0) for i in part) for part in gathered]
    assert max(loads) / (sum(loads) / world) < 1.25


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shard_voices_partitions(world):
    from mesheditor_b200 import workloads as wl

    spans = [wl.shard_voices(1024, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == 1024
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1
