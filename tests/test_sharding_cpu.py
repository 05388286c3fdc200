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


class OracleRank:
    """The ModalBank calls ShardedModalBank makes, answered by the CPU oracle (this is a test: the product's rank is a
    CUDA ModalBank). Events arrive as MeModalEvent and are handed to the oracle in its own struct."""

    def __init__(self, sample_rate, device):
        from oracle import resonator as orc

        orc.build_port()
        self.orc, self.sample_rate = orc, sample_rate
        self.bank = (orc.RefScene if orc.have_ref() else orc.PortBank)(sample_rate, 1)

    def add_object(self, *a):
        return self.bank.add_object(*a)

    def install(self, discard_frames=512):
        self.bank.install(discard_frames)

    def _event(self, e):
        return self.orc.Event(e.kind, e.object, e.ex_pos, e.jx, e.jy, e.jz, e.pulse_step, e.pulse_gamma, e.accel_amp, e.click_b0, e.click_a1, e.click_a2)

    def enqueue(self, e):
        self.bank.enqueue(self._event(e))
        return True

    def render(self, out):
        self.bank.render(out)

    def render_offline(self, packed, _frames, total, block):
        arr, frames, n = packed
        out, k = np.zeros(total, np.float32), 0
        for begin in range(0, total, block):
            while k < n and frames[k] == begin:
                self.enqueue(arr[k])
                k += 1
            self.bank.render(out[begin:begin + block])
        return out


def _unequal_bank(rank, world, all_reduce=None):
    """Six objects of unequal mode counts (the deal matters), the second one carrying a sustained voice in its cost."""
    from mesheditor_b200 import ShardedModalBank
    from mesheditor_b200 import workloads as wl

    bank = ShardedModalBank(48000.0, 0, rank, world, bank_factory=OracleRank, all_reduce=all_reduce)
    counts = [64, 8, 40, 64, 16, 24]
    for i, n in enumerate(counts):
        bank.add_modes(wl.make_modes(n, 0.5), voices=1 if i == 1 else 0)
    bank.install()
    frames = BLOCKS * BLOCK
    events, ev_frames, _ = wl.c5_timeline(len(counts), frames, restrike_hz=20.0)
    return bank, [wl.impact(v, impulse, ex) for v, impulse, ex in events], ev_frames, frames


def _mix_worker(rank, world, port, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bank, events, ev_frames, frames = _unequal_bank(rank, world)  # the default reduce: torch.distributed.all_reduce
    mix = bank.render_offline(events, ev_frames, frames, BLOCK)
    # streaming form: the same strikes enqueued block by block on every rank (SPMD), one all-reduce per block
    stream_bank, _, _, _ = _unequal_bank(rank, world)
    stream, k = np.zeros(frames, np.float32), 0
    for b in range(BLOCKS):
        while k < len(events) and ev_frames[k] == b * BLOCK:
            stream_bank.enqueue(events[k])
            k += 1
        stream_bank.render(stream[b * BLOCK:(b + 1) * BLOCK])
    owned = [None] * world
    dist.all_gather_object(owned, bank.owned())
    if rank == 0:
        result["mix"], result["stream"], result["owned"] = mix.copy(), stream.copy(), owned
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mix_matches_single_rank_render(built_lib):
    """ShardedModalBank over two gloo ranks == the same bank on one rank, at the reference's own gate for 1 vs N renderers
    (tests/ModalRenderTest.cpp:40-49); the deal is the reference's (heaviest first, least-loaded renderer)."""
    world, port = 2, _free_port()
    with mp.Manager() as manager:
        result = manager.dict()
        mp.spawn(_mix_worker, args=(world, port, result), nprocs=world, join=True)
        mix, stream, owned = np.array(result["mix"]), np.array(result["stream"]), list(result["owned"])
    # costs 64, 16, 40, 64, 16, 24 -> order 0, 3, 2, 5, 1, 4 -> loads (64, 64) (104, 64) (104, 88) (104, 104) (120, 104)
    assert owned == [[0, 2, 4], [1, 3, 5]]
    bank, events, ev_frames, frames = _unequal_bank(0, 1)
    whole = bank.render_offline(events, ev_frames, frames, BLOCK)
    peak = float(np.abs(whole).max())
    assert peak > 0
    assert float(np.abs(mix - whole).max()) <= 1e-5 * peak
    assert float(np.abs(stream - whole).max()) <= 1e-5 * peak


def _deal_restated(costs, count):
    """DealObjects, src/audio/ModalAudio.cpp:430-461, statement by statement (test-side restatement)."""
    if count == 1:
        return [0] * len(costs)
    order = sorted(range(len(costs)), key=lambda o: (-costs[o], o))
    load, owner = [0] * count, [0] * len(costs)
    for o in order:
        least = load.index(min(load))
        load[least] += costs[o]
        owner[o] = least
    return owner


@pytest.mark.parametrize("seed,n,world", [(0, 1, 1), (1, 7, 2), (2, 64, 8), (3, 1024, 8), (4, 33, 4), (5, 5, 8), (6, 0, 3)])
def test_deal_objects_is_the_reference_deal(built_lib, seed, n, world):
    from mesheditor_b200 import deal_objects

    rng = np.random.default_rng(seed)
    costs = (rng.integers(1, 60, n) * rng.integers(1, 4, n) * 8).astype(np.uint64) if seed != 3 else np.full(n, 500, np.uint64)
    owner, local = deal_objects(costs, world)
    assert list(owner) == _deal_restated([int(c) for c in costs], world)
    for r in range(world):  # each renderer takes its objects in bank order (:459)
        assert list(local[owner == r]) == list(range(int((owner == r).sum())))
    if seed == 3:  # identical voices: an even split, 128 per GPU
        assert np.bincount(owner, minlength=world).tolist() == [128] * 8


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
