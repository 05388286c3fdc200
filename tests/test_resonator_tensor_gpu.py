"""Parity of the tensor-core form of the resonator bank (walk kernel + tcgen05 mix, me_bank_set_render_path(2)) against
the oracle, on the cases of tests/test_resonator_gpu.py plus long offline timelines that span several 16384-frame tiles.
Same bar as the sample loop: 1e-5 of peak amplitude (BASELINE.json north_star; ModalRenderTest.cpp:48).
"""
import numpy as np
import pytest

from oracle import resonator as orc
from test_resonator_gpu import TOL, build_pair, gpu_bank, oracles, rel_err, to_me

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("oracle_cls", oracles())
@pytest.mark.parametrize("n_obj,n_modes,blocks,t60", [(1, 64, 8, 0.2), (16, 64, 32, 0.2), (3, 500, 16, 2.0), (5, 30, 12, 1.0), (2, 7, 4, 0.5)])
def test_streaming_blocks_match_oracle(oracle_cls, n_obj, n_modes, blocks, t60):
    """Block-by-block rendering: every 512-frame call is one ragged tile; culling is applied by the walk kernel itself."""
    o, g = build_pair(oracle_cls, n_obj, orc.make_modes(n_modes, t60))
    g.set_render_path(2)
    for obj in range(n_obj):
        ev = orc.impact_event(obj, 1.0 + 0.1 * obj, ex_pos=obj % 4)
        o.enqueue(ev)
        g.enqueue(to_me(ev))
    ro, rg = o.render_blocks(blocks), g.render_blocks(blocks)
    assert g.stats()["tensor_windows"] == 1
    assert np.abs(ro).max() > 0
    assert rel_err(rg, ro) <= TOL
    for col in ("StateRe", "StateIm"):
        a, b = o.mode_column(col), g.mode_column(col)
        scale = max(np.abs(o.mode_column("StateRe")).max(), np.abs(o.mode_column("StateIm")).max(), 1e-30)
        assert np.abs(a - b).max() <= 2e-5 * scale, col
    assert_same_status(o, g, n_obj)


def assert_same_status(o, g, n_obj):
    """LiveModeCount / Ringing (ModalAudio.h:126-127,133): the culling state the next render starts from."""
    live, ringing = o.object_column("LiveModeCount"), o.object_column("Ringing")
    for obj in range(n_obj):
        status = g.object_status(obj)
        assert (status["LiveModeCount"], status["Ringing"]) == (int(live[obj]), int(ringing[obj])), obj


def timeline(rng, n_obj, blocks, rate):
    events, frames = [], []
    for b in range(blocks):
        for obj in range(n_obj):
            if b == 0 or rng.random() < rate:
                events.append(orc.impact_event(obj, float(rng.uniform(0.2, 1.0)), int(rng.integers(0, 4)), float(1.0 / rng.integers(40, 1500))))
                frames.append(b * 512)
    return events, frames


def oracle_timeline(o, events, frames, total, block=512):
    ref = np.zeros(total, np.float32)
    k = 0
    for begin in range(0, total, block):
        while k < len(events) and frames[k] == begin:
            o.enqueue(events[k])
            k += 1
        o.render(ref[begin:min(begin + block, total)])
    return ref


@pytest.mark.parametrize("oracle_cls", oracles())
@pytest.mark.parametrize("n_obj,n_modes,t60,blocks,tail", [(6, 96, 1.5, 100, 0), (40, 500, 3.0, 70, 200), (3, 8, 2.0, 33, 77)])
def test_offline_timeline_over_several_tiles(oracle_cls, n_obj, n_modes, t60, blocks, tail):
    """Re-strikes on block boundaries, pulses of 40..1500 samples ending anywhere inside a time block, a ragged tail."""
    rng = np.random.default_rng(11)
    modes = orc.make_modes(n_modes, t60)
    o, g = build_pair(oracle_cls, n_obj, modes)
    g1 = gpu_bank()
    for _ in range(n_obj):
        g1.add_modes(modes)
    g1.install()
    g.set_render_path(2), g1.set_render_path(1)
    events, frames = timeline(rng, n_obj, blocks, 0.05)
    total = blocks * 512 + tail
    ref = oracle_timeline(o, events, frames, total)
    me_events = [to_me(e) for e in events]
    tensor = g.render_offline(me_events, frames, total, 512)
    stats = g.stats()
    loop = g1.render_offline(me_events, frames, total, 512)
    assert stats["tensor_windows"] >= 1  # (a small bank's seeded walk may have been repeated sequentially: scan_fallbacks)
    assert g1.stats()["tensor_windows"] == 0
    assert rel_err(tensor, ref) <= TOL
    # two approximations of the reference's own float recurrence, each within TOL of it
    assert rel_err(tensor, loop) <= 2.5 * TOL
    assert o.active_impacts() == g.active_impacts()
    scale = max(np.abs(o.mode_column("StateRe")).max(), np.abs(o.mode_column("StateIm")).max())
    for col in ("StateRe", "StateIm"):
        assert np.abs(o.mode_column(col) - g.mode_column(col)).max() <= 2e-5 * scale, col
    # a second call continues from the adopted state and the surviving impacts
    more, more_frames = timeline(rng, n_obj, 40, 0.05)
    ref2 = oracle_timeline(o, more, more_frames, 40 * 512)
    tensor2 = g.render_offline([to_me(e) for e in more], more_frames, 40 * 512, 512)
    assert rel_err(tensor2, ref2) <= TOL


@pytest.mark.parametrize("oracle_cls", oracles())
def test_culling_inside_the_timeline(oracle_cls):
    """Short T60s: objects fall silent and chunks drop out of the audible prefix mid-timeline. The walk kernel is
    sequential in time, so it applies those decisions itself and the window stays in the tensor-core form."""
    rng = np.random.default_rng(3)
    n_obj, blocks = 8, 80
    modes = orc.make_modes(64, 0.05)
    o, g = build_pair(oracle_cls, n_obj, modes)
    g.set_render_path(2)
    events, frames = timeline(rng, n_obj, blocks, 0.01)
    ref = oracle_timeline(o, events, frames, blocks * 512)
    out = g.render_offline([to_me(e) for e in events], frames, blocks * 512, 512)
    assert rel_err(out, ref) <= TOL
    assert_same_status(o, g, n_obj)


def test_muted_object_and_gain_changes():
    modes = orc.make_modes(120, 1.0)
    o, g = build_pair(orc.PortBank, 4, modes)
    g.set_render_path(2)
    o.set_gain(1, 0.0, 1.0), g.set_gain(1, 0.0, 1.0)
    o.set_gain(2, 0.5, 0.25), g.set_gain(2, 0.5, 0.25)
    rng = np.random.default_rng(2)
    events, frames = timeline(rng, 4, 64, 0.05)
    ref = oracle_timeline(o, events, frames, 64 * 512)
    out = g.render_offline([to_me(e) for e in events], frames, 64 * 512, 512)
    assert g.stats()["tensor_windows"] >= 1
    assert rel_err(out, ref) <= TOL


@pytest.mark.parametrize("oracle_cls", oracles())
@pytest.mark.parametrize("t60,shape_scale,expect_fallback", [(4000.0, 100.0, False), (0.05, 1.0, True)])
def test_seeded_walk_of_a_small_bank(oracle_cls, t60, shape_scale, expect_fallback):
    """Few chunk groups: the walk is split into seeded time segments (me_bank_set_time_segments forces 5 here); with
    short T60s culling falls inside them and the walk is repeated sequentially — the window stays in the tensor-core form."""
    rng = np.random.default_rng(17)
    n_obj, blocks = 5, 90
    # T60_k = t60 / k: the first case is the bench's recipe (no chunk ever falls under SilentEnergy), the second dies fast
    o, g = build_pair(oracle_cls, n_obj, orc.make_modes(200, t60, shape_scale))
    g.set_render_path(2)
    g.set_time_segments(5)
    events, frames = timeline(rng, n_obj, blocks, 0.02)
    ref = oracle_timeline(o, events, frames, blocks * 512)
    out = g.render_offline([to_me(e) for e in events], frames, blocks * 512, 512)
    stats = g.stats()
    assert stats["tensor_windows"] >= 1
    assert (stats["scan_fallbacks"] > 0) == expect_fallback
    assert stats["time_segments"] == (1 if expect_fallback else 5)
    assert rel_err(out, ref) <= TOL
    assert_same_status(o, g, n_obj)


@pytest.mark.parametrize("oracle_cls", oracles())
@pytest.mark.parametrize("rate,block", [(96000.0, 256), (48000.0, 1024), (44100.0, 2048)])
def test_clicks_silence_and_other_block_sizes(oracle_cls, rate, block):
    """Click-carrying impacts (they survive across blocks until their filter rings out), a Silence event in the middle
    of the timeline (it cuts the timeline into two spans and clears one object), other sample rates and block lengths."""
    from mesheditor_b200 import silence_event

    rng = np.random.default_rng(23)
    n_obj, blocks = 6, 48
    modes = orc.make_modes(80, 2000.0, 50.0)
    o, g = build_pair(oracle_cls, n_obj, modes, rate)
    g.set_render_path(2)
    click = orc.RefScene(rate, 1).click_filter(0.05, 4.0 / 3.0 * np.pi * 0.05**3, 1.0, rate) if orc.have_ref() else (0.0001865190570242703, -1.716620683670044, 0.7520560026168823)
    events, frames = [], []
    for b in range(blocks):
        for obj in range(n_obj):
            if b == 0 or rng.random() < 0.06:
                step = np.float32(1.0 / rng.integers(40, 900))
                events.append(orc.Event(0, obj, int(rng.integers(0, 4)), float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), float(rng.uniform(0.2, 1)), step, 2 * step, 0.5 * rate, *click))
                frames.append(b * block)
        if b == blocks // 2:
            events.append(orc.Event(1, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0))
            frames.append(b * block)
    total = blocks * block
    ref = oracle_timeline(o, events, frames, total, block)
    me_events = [silence_event(e.Object) if e.Kind == 1 else to_me(e) for e in events]
    out = g.render_offline(me_events, frames, total, block)
    assert g.stats()["tensor_windows"] >= 1
    assert rel_err(out, ref) <= TOL
    assert o.active_impacts() == g.active_impacts()
    assert_same_status(o, g, n_obj)


def test_object_larger_than_a_chunk_group():
    """2100 modes = 263 chunks: the object straddles chunk groups (no culling applies to it, as in the sample loop)."""
    modes = orc.make_modes(2100, 9000.0, 20.0)
    o, g = build_pair(orc.PortBank, 2, modes)
    g.set_render_path(2)
    rng = np.random.default_rng(4)
    events, frames = timeline(rng, 2, 70, 0.03)
    ref = oracle_timeline(o, events, frames, 70 * 512)
    out = g.render_offline([to_me(e) for e in events], frames, 70 * 512, 512)
    assert g.stats()["tensor_windows"] >= 1
    assert rel_err(out, ref) <= TOL


def test_automatic_path_choice():
    """Long spans of a large bank go to the tensor-core form, block-sized real-time calls stay on the sample loop."""
    modes = orc.make_modes(500, 10000.0, 100.0)
    g = gpu_bank()
    for _ in range(40):
        g.add_modes(modes)
    g.install()
    ev = [to_me(orc.impact_event(v, 1.0, v % 4)) for v in range(40)]
    g.render_offline(ev, [0] * 40, 64 * 32768, 512)  # 10 groups x 64 tiles
    assert g.stats()["tensor_windows"] >= 1
    g.render_blocks(2)
    assert g.stats()["tensor_windows"] == 0
    g.render_offline(ev, [0] * 40, 20 * 512, 512)  # too short to fill the SMs
    assert g.stats()["tensor_windows"] == 0
