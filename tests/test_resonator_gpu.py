"""Parity of the CUDA resonator bank (through the C ABI) against the oracle.

The cases mirror the reference's own harness and tests: tests/ModalBench.h (ModalScene, MakeModes, ImpactEvent)
and tests/ModalRenderTest.cpp:21-68. Tolerance: 1e-5 of peak amplitude in FP32 (BASELINE.json north_star; the
reference's own bar for thread-count independence, ModalRenderTest.cpp:48).
"""
import numpy as np
import pytest

from oracle import resonator as orc

pytestmark = pytest.mark.gpu

TOL = 1e-5


def oracles():
    out = [orc.PortBank]
    if orc.have_ref():
        out.append(orc.RefScene)
    return out


def gpu_bank(sample_rate=48000.0):
    from mesheditor_b200 import ModalBank

    return ModalBank(sample_rate, 0)


def to_me(ev):
    from mesheditor_b200 import MeModalEvent

    return MeModalEvent(ev.Kind, ev.Object, ev.ExPos, ev.Jx, ev.Jy, ev.Jz, ev.PulseStep, ev.PulseGamma, ev.AccelAmp, ev.ClickB0, ev.ClickA1, ev.ClickA2)


def build_pair(oracle_cls, n_obj, modes, sample_rate=48000.0):
    o = oracle_cls(sample_rate, 1)
    g = gpu_bank(sample_rate)
    for _ in range(n_obj):
        o.add_modes(modes)
        g.add_modes(modes)
    o.install()
    g.install()
    return o, g


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("oracle_cls", oracles())
def test_tuning_is_bit_exact(oracle_cls):
    """TuneModalObject runs on the host with the reference's libm calls: columns must be identical."""
    o, g = build_pair(oracle_cls, 3, orc.make_modes(67, 0.3))
    for col in ("CoeffRe", "CoeffIm", "RadiationGain", "RadiationArea", "OutPhaseIm", "OutPhaseRe", "DeflectionGain", "QuadCompliance", "QuadDriveScale"):
        assert np.array_equal(o.mode_column(col), g.mode_column(col)), col
    assert g.object_layout(1)["RadiantRadius"] == o.object_column("RadiantRadius")[1]


@pytest.mark.parametrize("oracle_cls", oracles())
@pytest.mark.parametrize("n_obj,n_modes,blocks,t60", [(1, 64, 8, 0.2), (16, 64, 32, 0.2), (3, 500, 16, 2.0), (5, 30, 12, 1.0), (2, 7, 4, 0.5)])
def test_strike_matches_oracle(oracle_cls, n_obj, n_modes, blocks, t60):
    o, g = build_pair(oracle_cls, n_obj, orc.make_modes(n_modes, t60))
    for obj in range(n_obj):
        ev = orc.impact_event(obj, 1.0 + 0.1 * obj, ex_pos=obj % 4)
        o.enqueue(ev)
        g.enqueue(to_me(ev))
    ro, rg = o.render_blocks(blocks), g.render_blocks(blocks)
    assert np.abs(ro).max() > 0
    assert rel_err(rg, ro) <= TOL


@pytest.mark.parametrize("oracle_cls", oracles())
def test_state_matches_oracle_while_ringing(oracle_cls):
    o, g = build_pair(oracle_cls, 2, orc.make_modes(100, 3.0))
    for obj in range(2):
        ev = orc.impact_event(obj, 1.0)
        o.enqueue(ev)
        g.enqueue(to_me(ev))
    o.render_blocks(6), g.render_blocks(6)
    for col in ("StateRe", "StateIm"):
        a, b = o.mode_column(col), g.mode_column(col)
        scale = max(np.abs(o.mode_column("StateRe")).max(), np.abs(o.mode_column("StateIm")).max())
        assert np.abs(a - b).max() <= 2e-5 * scale, col


@pytest.mark.parametrize("oracle_cls", oracles())
def test_superposition(oracle_cls):
    """tests/ModalRenderTest.cpp:21-37."""
    both = [orc.impact_event(0, 1.0, 0, 1.0 / 300.0), orc.impact_event(0, -0.4, 1, 1.0 / 90.0)]

    def render(events):
        g = gpu_bank()
        g.add_modes(orc.make_modes(64, 0.2))
        g.install()
        for e in events:
            g.enqueue(to_me(e))
        return g.render_blocks(8)

    a, b, together = render(both[:1]), render(both[1:]), render(both)
    assert np.abs(a).max() > 0 and np.abs(b).max() > 0
    assert np.abs(together - (a + b)).max() <= np.abs(together).max() * TOL
    # and against the oracle rendering both
    o = oracle_cls(48000.0, 1)
    o.add_modes(orc.make_modes(64, 0.2))
    o.install()
    for e in both:
        o.enqueue(e)
    assert rel_err(together, o.render_blocks(8)) <= TOL


@pytest.mark.parametrize("oracle_cls", oracles())
@pytest.mark.parametrize("rate", [48000.0, 96000.0])
def test_click(oracle_cls, rate):
    """tests/ModalRenderTest.cpp:53-68: the acceleration-noise click, here sample for sample."""
    tau, radius, mass, impulse = 5e-4, 0.05, 1.0, 0.5
    volume = 4.0 / 3.0 * np.pi * radius**3
    if orc.have_ref():
        click = orc.RefScene(rate, 1).click_filter(radius, volume, mass, rate)
    else:
        click = (0.0001865190570242703, -1.716620683670044, 0.7520560026168823)
    o, g = build_pair(oracle_cls, 1, orc.make_modes(64, 0.2), rate)
    step = np.float32(1.0 / (tau * rate))
    ev = orc.Event(0, 0, 0, 0.0, 0.0, 0.0, step, 2 * step, impulse * rate, *click)
    o.enqueue(ev)
    g.enqueue(to_me(ev))
    blocks = int(np.ceil(4 * tau * rate / 512))
    ro, rg = o.render_blocks(blocks + 2), g.render_blocks(blocks + 2)
    assert np.abs(ro).max() > 0
    assert rel_err(rg, ro) <= TOL
    assert o.active_impacts() == g.active_impacts()


@pytest.mark.parametrize("oracle_cls", oracles())
def test_offline_equals_streaming_and_oracle(oracle_cls):
    """me_bank_render_offline == the block loop; re-strikes land on block boundaries; impacts carry across calls."""
    rng = np.random.default_rng(5)
    n_obj, blocks = 6, 40
    modes = orc.make_modes(96, 1.5)
    o, g = build_pair(oracle_cls, n_obj, modes)
    g2 = gpu_bank()
    for _ in range(n_obj):
        g2.add_modes(modes)
    g2.install()
    events, frames = [], []
    for b in range(blocks):
        for obj in range(n_obj):
            if b == 0 or rng.random() < 0.08:
                events.append(orc.impact_event(obj, float(rng.uniform(0.2, 1.0)), int(rng.integers(0, 4)), float(1.0 / rng.integers(40, 1500))))
                frames.append(b * 512)
    ref = np.zeros(blocks * 512, np.float32)
    stream = np.zeros(blocks * 512, np.float32)
    k = 0
    for b in range(blocks):
        while k < len(events) and frames[k] == b * 512:
            o.enqueue(events[k])
            g.enqueue(to_me(events[k]))
            k += 1
        o.render(ref[b * 512:(b + 1) * 512])
        g.render(stream[b * 512:(b + 1) * 512])
    offline = g2.render_offline([to_me(e) for e in events], frames, blocks * 512, 512)
    assert rel_err(stream, ref) <= TOL
    assert rel_err(offline, ref) <= TOL
    assert o.active_impacts() == g.active_impacts() == g2.active_impacts()


@pytest.mark.parametrize("oracle_cls", oracles())
def test_silence_event_and_retune(oracle_cls):
    from mesheditor_b200 import silence_event

    o, g = build_pair(oracle_cls, 2, orc.make_modes(40, 2.0))
    for obj in range(2):
        ev = orc.impact_event(obj, 1.0)
        o.enqueue(ev)
        g.enqueue(to_me(ev))
    ro, rg = [o.render_blocks(3)], [g.render_blocks(3)]
    o.enqueue(orc.Event(1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0))
    g.enqueue(silence_event(0))
    m2 = orc.make_modes(40, 1.0)
    o.retune(1, m2["freqs"] * 1.5, m2["t60s"])
    g.retune(1, m2["freqs"] * 1.5, m2["t60s"])
    ro.append(o.render_blocks(3)), rg.append(g.render_blocks(3))
    ro, rg = np.concatenate(ro), np.concatenate(rg)
    assert rel_err(rg, ro) <= TOL
    assert np.array_equal(o.mode_column("CoeffRe"), g.mode_column("CoeffRe"))


def test_max_impacts_cap_and_queue_overflow():
    o, g = build_pair(orc.PortBank, 4, orc.make_modes(16, 0.5))
    o.set_max_impacts(3), g.set_max_impacts(3)
    for i in range(6):
        ev = orc.impact_event(i % 4, 1.0 + i)
        o.enqueue(ev), g.enqueue(to_me(ev))
    ro, rg = o.render_blocks(4), g.render_blocks(4)
    assert rel_err(rg, ro) <= TOL
    accepted = sum(g.enqueue(to_me(orc.impact_event(0, 1.0))) for _ in range(300))
    assert accepted == 256 and g.events_dropped() == 44
    for _ in range(300):
        o.enqueue(orc.impact_event(0, 1.0))
    assert o.events_dropped() == 44


def test_events_queued_before_first_render_are_flushed():
    """InstallModalBank flags queued events stale; the adopting render drops them (ModalAudio.cpp:280,496-498)."""
    g = gpu_bank()
    g.add_modes(orc.make_modes(8, 0.5))
    g.install(discard_frames=0)
    g.enqueue(to_me(orc.impact_event(0, 1.0)))
    assert np.abs(g.render_blocks(2)).max() == 0.0
    g.enqueue(to_me(orc.impact_event(0, 1.0)))
    assert np.abs(g.render_blocks(2)).max() > 0.0


def test_errors_are_reported_not_swallowed():
    from mesheditor_b200 import MeError

    g = gpu_bank()
    with pytest.raises(MeError):
        g.render(np.zeros(16, np.float32))  # not installed
    g.add_modes(orc.make_modes(8, 0.5))
    g.install()
    with pytest.raises(MeError):
        g.tune(7, [1.0], [1.0])
    with pytest.raises(MeError):
        g.render_offline([to_me(orc.impact_event(0, 1.0))], [100], 1024, 512)  # not on a block boundary


def test_objects_added_after_install_need_a_reinstall():
    """The reference never grows a live bank (RebuildModalBank builds the next one, InstallModalBank swaps it in,
    ModalAudio.cpp:277-289): a slot added after me_bank_install makes rendering fail loudly until the next install."""
    from mesheditor_b200 import MeError

    modes = orc.make_modes(40, 0.3)
    g = gpu_bank()
    g.add_modes(modes)
    g.install()
    g.enqueue(to_me(orc.impact_event(0, 1.0)))
    first = g.render_blocks(2)
    assert np.abs(first).max() > 0
    g.add_modes(orc.make_modes(500, 0.3))  # would need more chunks than the device buffers were sized for
    with pytest.raises(MeError):
        g.render_blocks(1)
    g.install()
    o = orc.PortBank(48000.0, 1)
    o.add_modes(modes), o.add_modes(orc.make_modes(500, 0.3))
    o.install()
    for bank, conv in ((o, lambda e: e), (g, to_me)):
        bank.enqueue(conv(orc.impact_event(1, 0.7, ex_pos=2)))
    assert rel_err(g.render_blocks(6), o.render_blocks(6)) <= TOL


def test_strike_outside_the_excitable_points_is_dropped():
    """TriggerModalStrike returns early for an excitable index past the object's points; the C ABI takes raw events, so the
    guard sits in the bank: the event is ignored (never a shape read out of bounds), later events still render."""
    modes = orc.make_modes(64, 0.2, sample_points=4)
    o, g = build_pair(orc.PortBank, 2, modes)
    bad = orc.impact_event(0, 1.0, ex_pos=4)      # one past the last point
    worse = orc.impact_event(1, 1.0, ex_pos=10**6)
    g.enqueue(to_me(bad)), g.enqueue(to_me(worse))
    assert np.abs(g.render_blocks(2)).max() == 0.0 and g.active_impacts() == 0
    good = orc.impact_event(1, 0.8, ex_pos=3)
    o.render_blocks(2)
    o.enqueue(good), g.enqueue(to_me(good))
    assert rel_err(g.render_blocks(8), o.render_blocks(8)) <= TOL
