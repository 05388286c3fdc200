#!/usr/bin/env python
"""bench.py — headline benchmark of the modal hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload all|resonator|solve|batch]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

BASELINE.json's metric has two halves; the default run (`--workload all`) measures both and prints ONE JSON line:

  * the line itself = the half quoted at 1/2/4/8 B200: BASELINE.json configs[4], the polyphonic resonator bank — 1024 voices x
    500 modes x 10 s at 48 kHz, one strike per voice at frame 0 plus Poisson re-strikes. A "step" is one offline render of the
    whole 10 s timeline. Voices are dealt over the ranks by the reference's DealObjects rule (mesheditor_b200.ShardedModalBank;
    strong scaling: the total stays 1024 voices) and the per-rank mono mixes are summed with one NCCL all-reduce over NVLink.
      `value`    = mode-samples/s with the bank resident in HBM, timed on the device (CUDA events, max over ranks).
      `e2e`      = the same metric through the C ABI with the strike timeline in HOST memory and the mix read back to HOST.
      `roofline` = the dominant kernel of the step (tensor-core form: the tcgen05 mix; sample loop: FP32 issue).
      `parity`   = the output of THIS configuration checked outside the timed region: the full 1024-voice mix against the
                   reference's own render of the same timeline, and a 16-voice slice against the reference and the FP64 arbiter.
      `cpu_baseline` = the UNMODIFIED reference RenderModal (oracle/_ref) on the host cores, one full-size step.
  * `"solve"`  = the other half, "modal solve s/mesh @1M tets": BASELINE.json configs[2] through me_modal_solve (host mesh in,
    host modal model out) with the reference bench's per-stage table (tests/ModalSolverBench.cpp:413-449) and the stage
    rooflines (the 8-wide triangular-solve sweeps that carry the timed solve, SpMV, assembly, numeric factorisation).
    A single eigensolve does not shard: with --gpus N every rank solves a replica ("replicas only") and the slowest is reported.
  * `"batch"`  = BASELINE.json configs[3]: 64 meshes (10k..500k tets) dealt biggest-first over the ranks, no collective.

`--impl reference` runs the CPU arm alone (the reference's RenderModal over the same full configuration) in the same format.
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
PARITY_SLICE_VOICES = 16


def env_int(name, default):
    return int(os.environ.get(name, default))


def profiled_traffic():
    """DRAM bytes per launch of the tensor-form kernels from the committed `ncu --set full` capture of the N=1 step, with the
    launch shape it was taken at (profiles/resonator_traffic.json), so that a rank's figure can be scaled to ITS launch."""
    try:
        with open(os.path.join(ROOT, "profiles", "resonator_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """One `nvidia-smi -lms` process sampling clocks and throttle reasons while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, period_ms=None):
        self.index, self.proc, self.rows = index, None, []
        self.period_ms = int(os.environ.get("ME_CLOCK_SAMPLE_MS", period_ms or 50))

    def __enter__(self):
        if self.period_ms <= 0:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


# ------------------------------------------------------------------------------------------------ synthesis: the CPU arm
def workload_config(n_gpus):
    """The same dict in both arms (the driver compares them): nothing run-dependent in here."""
    return {
        "workload": f"BASELINE.json configs[4]: polyphonic resonator bank {VOICES} voices x {MODES} modes x {SECONDS:g} s at {RATE:g} Hz, strike per voice at frame 0 + 2 Hz Poisson re-strikes (MT19937 12345), 512-frame blocks",
        "voices": VOICES, "modes_per_voice": MODES, "frames": int(SECONDS * RATE), "sample_rate": RATE, "block_frames": BLOCK,
        "parallelism": f"voices dealt over {n_gpus} GPU(s) (DealObjects rule), one NCCL all-reduce of the mono mix" if n_gpus > 1 else "single GPU",
        "l2_policy": "working set > L2: one step writes and reads 8 GB of block-start states (tensor-core form) or 3.9 GB of per-warp partial mixes (sample loop); nothing is reused across steps",
    }


RING = 256  # ModalAudio::EventCapacity (ModalAudio.h:275): strikes the reference can take between two RenderModal calls


class ReferenceScenes:
    """The reference bank(s) of the CPU arm. The workload strikes all 1024 voices in the same block, and the reference's event
    ring holds 256 events between two RenderModal calls (the rest are dropped and counted, ModalAudio.cpp:419-422), so the voices
    live in ceil(voices / 256) ModalAudio instances of 256 voices each, rendered one after the other INTO THE SAME buffer
    (RenderModal adds into `out`): the same code, the same total work, no strike lost."""

    def __init__(self, voices, threads):
        from mesheditor_b200 import workloads as wl
        from oracle import resonator as orc

        self.orc = orc
        self.kind = "reference" if orc.have_ref() else "port"
        cls = orc.RefScene if self.kind == "reference" else orc.PortBank
        modes = wl.c5_modes(MODES)
        self.scenes = []
        for first in range(0, voices, RING):
            scene = cls(RATE, threads if self.kind == "reference" else 1)
            for _ in range(min(RING, voices - first)):
                scene.add_modes(modes)
            scene.install()
            self.scenes.append(scene)

    def step(self, events, ev_frames, frames):
        """One step of the reference arm: RenderModal over the whole timeline in 512-frame blocks, the strikes of a block enqueued
        before it (tests/ModalBench.h:76-80; the offline loop of src/audio/AudioSystem.cpp:1155-1159)."""
        out, k = np.zeros(frames, np.float32), 0
        t0 = time.perf_counter()
        for begin in range(0, frames, BLOCK):
            while k < len(events) and ev_frames[k] == begin:
                v, impulse, ex = events[k]
                self.scenes[v // RING].enqueue(self.orc.impact_event(v % RING, impulse, ex))
                k += 1
            for scene in self.scenes:
                scene.render(out[begin:min(begin + BLOCK, frames)])
        return time.perf_counter() - t0, out

    def live_mode_counts(self):
        return np.concatenate([s.object_column("LiveModeCount") for s in self.scenes])

    def events_dropped(self):
        return int(sum(s.events_dropped() for s in self.scenes))


def reference_threads():
    """The renderer count the CPU arm uses: the reference's pool takes up to the core count (ModalRenderPool::SetSize clamps to
    hardware_concurrency, ModalAudio.cpp:238-243; its UI offers 1..16, AudioSystem.cpp:60). Both are timed on a 1 s prefix of the
    workload and the faster one is used, so the baseline is the reference at its best on this host."""
    from mesheditor_b200 import workloads as wl
    from oracle import resonator as orc

    cores = os.cpu_count() or 1
    if not orc.have_ref():
        return 1, {}
    frames = int(RATE)
    events, ev_frames, _ = wl.c5_timeline(VOICES, frames)
    rates = {}
    for threads in sorted({min(16, cores), cores}):
        dt, _ = ReferenceScenes(VOICES, threads).step(events, ev_frames, frames)
        rates[threads] = VOICES * MODES * frames / dt
    return max(rates, key=rates.get), {str(k): float(f"{v:.4g}") for k, v in rates.items()}


def cpu_reference(steps, warmup):
    """The reference's own RenderModal (oracle/_ref/libme_ref_audio.so, built from /root/reference sources) on the host cores over
    the FULL configuration: every step is 1024 voices x 480,000 frames. Returns the figures and the last step's mix."""
    from mesheditor_b200 import workloads as wl

    frames = int(SECONDS * RATE)
    threads, calibration = reference_threads()
    events, ev_frames, _ = wl.c5_timeline(VOICES, frames)
    times, out, live, kind, dropped = [], None, None, "port", 0
    for it in range(warmup + steps):
        scenes = ReferenceScenes(VOICES, threads)
        dt, out = scenes.step(events, ev_frames, frames)
        live, kind, dropped = scenes.live_mode_counts(), scenes.kind, scenes.events_dropped()
        if it >= warmup:
            times.append(dt)
    assert dropped == 0, f"the reference dropped {dropped} strikes"
    ms = 1e3 * sum(times) / len(times)
    return {
        "value": VOICES * MODES * frames / (ms * 1e-3), "unit": UNIT, "cores": threads, "kind": kind, "ms_per_step": ms, "host_cores": os.cpu_count() or 1,
        "sample": f"the full workload per step: {VOICES} voices x {MODES} modes x {frames} frames in {-(-frames // BLOCK)} blocks of {BLOCK}, same strike timeline, {steps} timed step(s) after {warmup} warm-up; "
                  f"{-(-VOICES // RING)} ModalAudio banks of {RING} voices rendered into one buffer (the reference's event ring holds {RING} strikes per block; none dropped); "
                  f"RenderPool of {threads} renderer(s) (faster of the UI's 16 and the pool's own ceiling = the core count, 1 s calibration in mode-samples/s: {calibration}); LiveModeCount {int(live.min())}..{int(live.max())} of {MODES} (no culling)",
    }, out


def run_reference(args):
    if env_int("RANK", 0) != 0:
        return
    base, _ = cpu_reference(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus), "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ synthesis: the GPU arm
class Ranks:
    """torch.distributed plumbing shared by the three workloads (one process per GPU, NCCL)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        from mesheditor_b200 import build

        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
        build.build()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN, "sum": self.dist.ReduceOp.SUM}[op])
        return [float(x) for x in t]

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def sharded_bank(ranks, voices, render_path):
    from mesheditor_b200 import ShardedModalBank
    from mesheditor_b200 import workloads as wl

    bank = ShardedModalBank(RATE, ranks.local, ranks.rank, ranks.world)
    modes = wl.c5_modes(MODES)
    for _ in range(voices):
        bank.add_modes(modes)
    bank.install(0)
    bank.set_render_path({"auto": 0, "loop": 1, "tensor": 2}[render_path])
    return bank


def parity_report(ranks, full_mix, reference_mix, tensor_form, time_segments):
    """Outside the timed region: is what was just timed the right audio? (a) the whole N-rank mix of configs[4] against the
    reference's render of the same timeline; (b) a 16-voice slice of the same timeline, full length, dealt over the same ranks,
    against the reference AND the FP64 arbiter (oracle/slices.py), which tells whose rounding moved."""
    from mesheditor_b200 import workloads as wl
    from oracle import slices

    torch = ranks.torch
    frames = int(SECONDS * RATE)
    events, ev_frames = slices.c5_slice_timeline(PARITY_SLICE_VOICES, frames)
    bank = sharded_bank(ranks, PARITY_SLICE_VOICES, "tensor" if tensor_form else "loop")
    if time_segments == 1:
        bank.set_time_segments(1)  # the walk the timed 1024-voice bank took
    routed = bank.route_events([wl.impact(v, impulse, ex) for v, impulse, ex in events], ev_frames)
    out = torch.zeros(frames, dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()  # an explicit stream: the library's kernels and the all-reduce must share one (NULL would mean the bank's own)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        bank.render_offline_device(routed, frames, BLOCK, out, stream.cuda_stream)
    stream.synchronize()
    if ranks.rank != 0:
        return None
    gpu_slice = out.cpu().numpy()
    ref, kind = slices.reference_render(PARITY_SLICE_VOICES, frames, threads=min(16, os.cpu_count() or 1), events=events, ev_frames=ev_frames)
    exact = slices.exact_render(PARITY_SLICE_VOICES, frames, events=events, ev_frames=ev_frames)
    rep = {"gate": slices.GATE, "unit": "fraction of the reference render's peak amplitude", "oracle": kind,
           "slice": dict(slices.compare(gpu_slice, ref, exact), voices=PARITY_SLICE_VOICES, frames=frames, ranks=ranks.world)}
    if reference_mix is not None:
        peak = float(np.abs(reference_mix).max())
        rep["full_config"] = {"voices": VOICES, "frames": frames, "peak": peak, "gpu_vs_reference": float(np.abs(full_mix.astype(np.float64) - reference_mix).max() / peak),
                              "note": "the N-rank mix of the timed configuration against the reference's float32 RenderModal of the same timeline (the cpu_baseline leg's own output)"}
    s = rep["slice"]
    rep["verdict"] = ("within the gate of the exact recurrence over all 10 s" if s["gpu_vs_exact"] <= slices.GATE else "OUTSIDE the gate of the exact recurrence") + \
                     f"; the reference's own float32 recurrence is {s['reference_vs_exact']:.2e} of peak from the exact value by the end, so |gpu - reference| = {s['gpu_vs_reference']:.2e} is the reference's drift, not the GPU's"
    return rep


def run_resonator(args, ranks):
    from mesheditor_b200 import ModalBank, measure_fp32_fma_rate  # noqa: F401
    from mesheditor_b200 import workloads as wl

    torch, dist = ranks.torch, ranks.dist
    rank, world, local = ranks.rank, ranks.world, ranks.local
    voices = args.voices or VOICES
    frames = int(SECONDS * RATE)
    bank = sharded_bank(ranks, voices, args.render_path)
    mine = len(bank.owned())
    all_events, all_frames, _ = wl.c5_timeline(voices, frames)
    routed = bank.route_events([wl.impact(v, impulse, ex) for v, impulse, ex in all_events], all_frames)  # contiguous MeModalEvent[] + frames, built once

    out = torch.zeros(frames, dtype=torch.float32, device="cuda")
    host_out = torch.zeros(frames, dtype=torch.float32).pin_memory()
    # One explicit stream carries the library's kernels, the NCCL all-reduce and the timing events.
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def step_device():
        bank.render_offline_device(routed, frames, BLOCK, out, stream.cuda_stream)

    def step_e2e():
        # Strike timeline from host memory in, final mix back in host memory.
        bank.render_offline_device(routed, frames, BLOCK, out, stream.cuda_stream)
        host_out.copy_(out, non_blocking=True)
        stream.synchronize()

    # Each step re-renders the same 10 s from a silent bank, so every step does identical work. (Tuning the bank — the power
    # stages of the tensor-core form, 2.7 ms on the device — happens once at setup, like the reference's TuneModalObject.)
    def reset():
        bank.local.install(0)

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        reset()
        step_device()
    ranks.barrier()

    kernel_ms, walk_ms, mix_ms, pulse_ms, plan_ms, stats, launches = [], [], [], [], [], None, 0
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    elapsed_ms = 0.0
    with ClockSampler(local) as clocks:
        for _ in range(args.steps):
            reset()
            ranks.barrier()
            start.record(stream)
            step_device()
            stop.record(stream)
            ranks.barrier()
            elapsed_ms += start.elapsed_time(stop)
            stats = bank.stats()
            if os.environ.get("ME_BENCH_DEBUG"):
                print(f"[bench] step {start.elapsed_time(stop):.2f} ms, library total {stats['total_device_ms']:.2f} ms, walk {stats['walk_kernel_ms']:.2f}, mix {stats['tensor_mix_kernel_ms']:.2f}, host plan {stats['host_plan_ms']:.2f}, pulses {stats['pulse_kernels_ms']:.2f}, stage {stats['resonator_kernel_ms']:.2f}", file=sys.stderr)
            kernel_ms.append(stats["resonator_kernel_ms"]), walk_ms.append(stats["walk_kernel_ms"]), mix_ms.append(stats["tensor_mix_kernel_ms"])
            pulse_ms.append(stats["pulse_kernels_ms"]), plan_ms.append(stats["host_plan_ms"])
            launches += stats["kernel_launches"]
    # e2e: wall clock around the host-buffer path.
    e2e_s = 0.0
    for _ in range(args.steps):
        reset()
        ranks.barrier()
        t0 = time.perf_counter()
        step_e2e()
        e2e_s += time.perf_counter() - t0
        launches += bank.stats()["kernel_launches"]
    ranks.barrier()
    full_mix = host_out.numpy().copy()

    live = [bank.local.object_status(v)["LiveModeCount"] for v in range(mine)] or [MODES]
    fallbacks = stats["scan_fallbacks"]
    elapsed_ms, e2e_s = ranks.reduce([elapsed_ms, e2e_s], "max")
    total_launches = int(ranks.reduce([float(launches)], "sum")[0])
    min_live = int(ranks.reduce([float(min(live)) - 1e6 * fallbacks], "min")[0])
    tensor_form = stats["tensor_windows"] > 0

    line = None
    if rank == 0:
        avg = lambda xs: sum(xs) / len(xs)  # noqa: E731
        total_mode_samples = voices * MODES * frames
        ms_per_step = elapsed_ms / args.steps
        value = total_mode_samples / (ms_per_step * 1e-3)
        e2e_value = total_mode_samples / (e2e_s / args.steps)
        pk, pk_kind = peaks()
        fma_peak = measure_fp32_fma_rate(local, 0, 10)       # scalar FFMA, the nominal FP32 issue ceiling
        fma_peak_fresh = measure_fp32_fma_rate(local, 5, 10)  # FFMA2 with three fresh register pairs (register-file bound)
        k_ms = avg(kernel_ms)
        rank_mode_samples = mine * MODES * frames
        achieved = OPS_PER_MODE_SAMPLE * rank_mode_samples / (k_ms * 1e-3)
        # Mandatory bytes of one launch on this rank: 32 B per mode of state/coefficients + the mix written once.
        mandatory = mine * (32 * MODES) + 4 * frames
        base, reference_mix = (None, None) if args.no_cpu_baseline else cpu_reference(1, 0)
        if tensor_form:
            # Tensor-core form. Dominant kernels: the tcgen05 mix (FP16-split GEMM) and the state walk that
            # feeds it. Issued flops and written bytes follow from the padded layout: 4 voices of 63 chunks per 256-chunk
            # group, 4096 reduction elements per group, 128 time blocks of 256 frames per tile.
            chunks = -(-MODES // 8)
            groups = -(-mine // (256 // chunks))
            tiles = -(-frames // 32768)
            product_flops = 2.0 * 256 * 128 * 4096 * groups * tiles  # one of the three products of the split (hi*hi, hi*lo, lo*hi)
            # All three products run as kind::f16 MMAs (FP16 operands, FP32 accumulation): "achieved" is issued FP16 flop/s
            # against the measured dense 16-bit tensor peak.
            issued_flops = 3 * product_flops
            walk_bytes = 4096 * 4.0 * groups * -(-frames // 256)
            m_ms, w_ms = avg(mix_ms), avg(walk_ms)
            f16_peak = pk.get("bf16_tflops", 2250.0)
            stages = groups * 256 * tiles
            copy_bytes = stages * 24576.0  # per 16-element stage: 8 KB of states + 16 KB of power images through the SM's L2 port
            prof = profiled_traffic()
            # The ncu figure belongs to the launch it was captured on; a rank's launch covers (its groups x tiles) of that.
            scale = groups * tiles / max(1, prof.get("groups", 256) * prof.get("tiles", 15))
            traffic = lambda which: (prof[f"{which}_dram_bytes_per_launch"] * scale if f"{which}_dram_bytes_per_launch" in prof else None)  # noqa: E731
            roofline = {
                "bound": "tensor", "kernel": "TensorMixKernel<128,8> (tcgen05.mma kind::f16: hi*hi + hi*lo + lo*hi of the two-term FP16 split, FP32 accumulation in TMEM, FP32 register folds)",
                "achieved": issued_flops / (m_ms * 1e-3) / 1e12, "peak": f16_peak, "unit": "TFLOP/s (FP16 operands)", "frac": issued_flops / (m_ms * 1e-3) / 1e12 / f16_peak,
                "traffic": traffic("mix"), "traffic_source": f"ncu --set full dram__bytes of the {prof.get('groups', 256)}-group x {prof.get('tiles', 15)}-tile launch (profiles/), scaled by this rank's groups x tiles = {groups} x {tiles}",
                "kernel_ms_per_launch": m_ms, "launches_per_step": stats["tensor_windows"],
                "issued_flops_per_step": issued_flops, "algorithmic_flops_per_step": product_flops, "share_of_step": m_ms / ms_per_step,
                "peak_source": ("MEASURED_PEAKS.json bf16_tflops (FP16 and BF16 share the tensor rate; nominal 2250)" if "bf16_tflops" in pk else "nominal dense FP16 2250 TFLOP/s"),
                "reference_fma_equivalent": REF_OPS_PER_MODE_SAMPLE * rank_mode_samples / (m_ms * 1e-3) / 1e12,
                "l2_to_sm": {"bytes_per_launch": copy_bytes, "achieved_bytes_per_clk_per_sm": copy_bytes / (m_ms * 1e-3) / (148 * (pk.get("sm_max_mhz", 1965.0) * 1e6)), "peak_bytes_per_clk_per_sm": 64,
                             "note": "what the kernel runs into next to the tensor pipe: 24 KB per stage of 384 tensor cycles is the SM's whole 64 B/clk L2 read port (at the ~1.55 GHz the SM holds inside this kernel the figure is 1.27x higher; DESIGN.md 5.2)"},
            }
            extra = {
                "roofline_walk": {"bound": "hbm", "kernel": "ResonatorKernel<1,2,true> (state walk: c^256 steps, scaled FP16 hi/lo state rows)", "achieved": walk_bytes / (w_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                  "frac": walk_bytes / (w_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": traffic("walk"), "kernel_ms_per_launch": w_ms, "algorithmic_bytes": walk_bytes, "share_of_step": w_ms / ms_per_step, "peak_source": pk_kind,
                                  "note": "its sub-window launches run beside the pulse kernels of the next sub-window (both slower for it): kernel_ms is the sum of the launches' own spans"},
                "fp32_pipe_equivalent": {"note": "the same mode-samples per second on the FP32 pipe would need this multiple of the measured scalar-FFMA ceiling (reference loop: 7 lane-ops per mode-sample; this repo's sample loop: 2.75)",
                                         "reference_loop": REF_OPS_PER_MODE_SAMPLE * rank_mode_samples / (k_ms * 1e-3) / fma_peak, "sample_loop": achieved / fma_peak, "ffma_peak_tlane_ops": fma_peak / 1e12},
            }
        else:
            roofline = {
                "bound": "fp32_fma", "kernel": "ResonatorKernel<4,2>", "achieved": achieved / 1e12, "peak": fma_peak / 1e12, "unit": "TFMA-lane-op/s",
                "frac": achieved / fma_peak, "traffic": None, "kernel_ms_per_launch": k_ms,
                "ops_per_mode_sample": OPS_PER_MODE_SAMPLE, "peak_source": "FFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP32 figure)",
                "peak_three_fresh_operands": fma_peak_fresh / 1e12, "frac_of_register_file_bound": achieved / fma_peak_fresh,
                "reference_op_equivalent_frac": REF_OPS_PER_MODE_SAMPLE * rank_mode_samples / (k_ms * 1e-3) / fma_peak,
            }
            extra = {}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (two-term FP16 split products on the tensor cores, FP32 accumulation; 22 significant bits per operand)" if tensor_form else "f32", "data": "synthetic",
            "config": workload_config(world) if voices == VOICES else dict(workload_config(world), voices=voices, workload=f"EXPERIMENT: {voices} voices (not the BASELINE.json configuration)"),
            "run": {"live_mode_count_min": min_live, "culling_triggered": min_live < MODES, "time_segments": stats["time_segments"], "voices_on_rank0": mine,
                    "render_path": "tensor-core form: state walk + tcgen05 mix" if tensor_form else "FP32 sample loop", "partial_rows": stats["partial_rows"],
                    "step_breakdown_ms_rank0": {"walk": avg(walk_ms), "tensor_mix": avg(mix_ms), "force_and_pulse_kernels": avg(pulse_ms), "host_planning": avg(plan_ms), "whole_step": ms_per_step},
                    "outside_the_timed_step": "bank tuning (TuneModalObject on the host, power stages of the tensor-core form: PowerTableKernel, once per tuning) and me_bank_install, as in the reference where tuning happens at setup"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(stats["h2d_bytes"]), "d2h_bytes_per_step": frames * 4, "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": total_launches,
            "roofline": roofline,
            **extra,
            "roofline_hbm": {"bound": "hbm", "achieved": mandatory / (k_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": mandatory / (k_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "peak_source": pk_kind,
                             "note": "mandatory bytes of the reference formulation only (SURVEY.md F9)"},
            "clocks": clocks.summary(),
        }
        if base:
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    else:
        reference_mix = None
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())
    if not args.no_parity and voices == VOICES:
        rep = parity_report(ranks, full_mix, reference_mix, tensor_form, stats["time_segments"])
        if rank == 0:
            line["parity"] = rep
    return line


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
                      f"{dt:.1f} s for this sample; the full 998,250-tet mesh does not finish on the host in the bench's time budget (SuperLU's fill grows ~n^(4/3): a stage-wise extrapolation would be a guess, so none is given)"}


def solve_once(points, tets, ex, cfg):
    from mesheditor_b200 import mesh2modes

    t0 = time.perf_counter()
    r = mesh2modes(points, tets, "Steel", ex, config=cfg)
    return time.perf_counter() - t0, r


STAGES = ("mass_props", "assemble", "sample_excite", "factorize", "analyse", "iterate", "op_solve", "extract", "dofs", "stiffness_nonzeros", "op_applications", "restarts", "factor_nonzeros", "supernodes", "levels")


def sweep_traffic(factor_nonzeros):
    """DRAM bytes of one panel application from the committed ncu capture, if it was taken on a factor of this size."""
    try:
        with open(os.path.join(ROOT, "profiles", "solve_traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_per_panel_application"]) if int(t["factor_nonzeros"]) == int(factor_nonzeros) else None
    except (OSError, KeyError, ValueError):
        return None


def run_solve(args, ranks, steps, warmup, cpu_baseline=True):
    """configs[2]: one 1M-tet mesh on one GPU (a single eigensolve does not shard: every rank solves a replica and the slowest
    is reported). Returns the record (rank 0) with the reference bench's stage table (tests/ModalSolverBench.cpp:413-449)."""
    from mesheditor_b200 import Factor, FemSystem, measure_fp64_rate, solver_config
    from mesheditor_b200 import workloads as wl

    torch = ranks.torch
    rank, world, local = ranks.rank, ranks.world, ranks.local
    points, tets = wl.kuhn_block(C3_EDGE, C3_EDGE, C3_EDGE, (0.3, 0.3, 0.3))
    ex = wl.bench_excitations(points)
    cfg = solver_config(num_modes=C3_MODES, element_order=C3_ORDER, max_mode_freq=1e9, device=local)
    for _ in range(warmup):
        solve_once(points, tets, ex, cfg)
    times, profiles = [], []
    with ClockSampler(local) as clocks:
        for _ in range(steps):
            ranks.barrier()
            dt, r = solve_once(points, tets, ex, cfg)  # host mesh in, host modal model out: this IS the end-to-end call
            assert r.status == 0 and len(r.freqs) == C3_MODES
            times.append(dt)
            profiles.append(r.profile)
    sec = ranks.reduce([sum(times) / len(times)], "max")[0]
    if rank != 0:
        return None
    prof = {k: (sum(p[k] for p in profiles) / len(profiles) if isinstance(profiles[0][k], float) else profiles[-1][k]) for k in profiles[0]}
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
    # The timed solve applies the operator to PANELS of 8 Krylov vectors: the kernels on its path are the 8-wide sweeps
    # (solve_panel, CholeskyShiftInvert.cpp:55-62). Timed here with CUDA events around one panel application, 5 repetitions.
    sweep_bytes = 16 * fi["factor_nonzeros"] + 16 * i["dofs"] * 8
    b8 = np.asfortranarray(np.random.default_rng(1).standard_normal((i["dofs"], 8)))
    panel_ms = []
    for _ in range(6):
        f.solve(b8)
        panel_ms.append(f.info["last_solve_device_ms"])
    panel_ms = sum(panel_ms[1:]) / 5
    b1 = b8[:, 0].copy()
    single_ms = []
    for _ in range(4):
        f.solve(b1)
        single_ms.append(f.info["last_solve_device_ms"])
    single_ms = sum(single_ms[1:]) / 3
    ops = prof["op_applications"]
    panels = -(-ops // 8)
    rec = {
        "metric": SOLVE_METRIC, "value": sec, "unit": SOLVE_UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * sec, "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(solve_config(), parallelism="replicas only: one mesh's eigensolve runs on one GPU; every rank solves the same mesh, the slowest is reported" if world > 1 else "single GPU"),
        "e2e": {"value": sec, "unit": SOLVE_UNIT, "h2d_bytes_per_step": int(points.nbytes + tets.nbytes + ex.nbytes), "d2h_bytes_per_step": int(8 * (C3_MODES + 15) + 4 * 3 * 10 * (C3_MODES + 15) + 4 * 3 * 10),
                "note": "me_modal_solve takes the mesh from HOST memory and returns the modal model to HOST memory; value and e2e are the same call"},
        "device_seconds": sec - prof["mass_props"], "gpu_launches": int(sum(p["kernel_launches"] for p in profiles)),
        "seconds_each": times, "profile": {k: prof[k] for k in STAGES},
        "roofline": {"bound": "hbm", "kernel": "WideSweepKernel<0> + WideSweepKernel<1> (forward + backward triangular sweeps over the factor for a panel of 8 right-hand sides; WideBegin / MarkUnsolved / WidePermuteOut ride in the same event pair, ~1 % of it)",
                     "achieved": sweep_bytes / (panel_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": sweep_bytes / (panel_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": sweep_traffic(fi["factor_nonzeros"]),
                     "traffic_source": "ncu --set full dram__bytes of one forward + one backward panel sweep of this factor (profiles/solve_traffic.json); null for any other factor",
                     "ms_per_launch": panel_ms, "algorithmic_bytes": sweep_bytes, "share_of_step": prof["op_solve"] / sec, "panel_applications_per_step": panels, "operator_applications_per_step": ops,
                     "op_solve_ms_per_panel_in_the_timed_solve": 1e3 * prof["op_solve"] / panels, "peak_source": pk_kind,
                     "bytes": "16 * nnz(L) (the factor read once forward and once backward, 8 B each) + 16 * n * 8 (eight right-hand sides in and out)"},
        "roofline_single_sweep": {"kernel": "SweepKernel<0/1> (one right-hand side; NOT on the timed path, kept for comparison)", "bound": "hbm", "ms_per_launch": single_ms, "achieved": (16 * fi["factor_nonzeros"] + 16 * i["dofs"]) / (single_ms * 1e-3) / 1e9,
                                  "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": (16 * fi["factor_nonzeros"] + 16 * i["dofs"]) / (single_ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
        "roofline_spmv": spmv, "roofline_assembly": asm, "roofline_factor": factor,
        "clocks": clocks.summary(),
    }
    rec["other_configs"] = other_solve_configs(local)
    if cpu_baseline:
        base = cpu_solve_reference()
        rec["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    return rec


def other_solve_configs(device):
    """The analysis configurations of BASELINE.json next to configs[2], one warm call each after one warm-up call: configs[0]
    (the reference's own CPU-runnable case, quadratic elements) and configs[1] (the ~200k-tet torus, lowest 100 modes) in the
    linear elements BASELINE.json names AND in the reference's quadratic elements (both orders)."""
    from mesheditor_b200 import mesh2modes, solver_config
    from mesheditor_b200 import workloads as wl

    out = []
    for label, mesh, material, order, modes in (
        ("configs[0]: tetrahedralised icosphere, steel, lowest 30 modes, quadratic (P2) elements", wl.config1_mesh()[:2], "Steel", 2, 30),
        ("configs[1]: synthetic torus, lowest 100 modes, linear (P1) elements", wl.torus_mesh(), "Ceramic", 1, 100),
        ("configs[1] in the reference's quadratic (P2) elements", wl.torus_mesh(), "Ceramic", 2, 100),
    ):
        points, tets = mesh
        ex = wl.bench_excitations(points)
        cfg = solver_config(num_modes=modes, element_order=order, max_mode_freq=1e9, device=device)
        mesh2modes(points, tets, material, ex, config=cfg)
        t0 = time.perf_counter()
        r = mesh2modes(points, tets, material, ex, config=cfg)
        dt = time.perf_counter() - t0
        p = r.profile
        out.append({"workload": label, "tets": int(len(tets)), "element_order": order, "num_modes": modes, "status": int(r.status), "kept_modes": int(len(r.freqs)), "value": dt, "unit": SOLVE_UNIT,
                    "profile": {k: p[k] for k in ("assemble", "factorize", "analyse", "iterate", "op_solve", "dofs", "op_applications", "restarts", "factor_nonzeros", "supernodes", "levels")}})
    return out


def run_batch(args, ranks, steps, warmup):
    """BASELINE.json configs[3], 64 Kuhn blocks of 10k..500k tets, dealt biggest-first to the least-loaded rank (the reference
    bench's biggest-file-first rule, tests/ModalSolverBench.cpp:475-478); no collective on the data path."""
    from mesheditor_b200 import solver_config
    from mesheditor_b200 import workloads as wl

    rank, world, local = ranks.rank, ranks.world, ranks.local
    dims = wl.config4_dims()
    owner = wl.lpt_assign([(6.0 * d ** 3) ** (4.0 / 3.0) for d in dims], world)
    mine = sorted((d for d, o in zip(dims, owner) if o == rank), reverse=True)
    cfg = solver_config(num_modes=30, element_order=1, max_mode_freq=1e9, device=local)
    meshes = [wl.kuhn_block(d, d, d, (0.3, 0.3, 0.3)) for d in mine]

    # Solves in flight per GPU: every me_modal_solve runs on a stream of its own, so a few of them on host threads overlap one
    # solve's host stages (symbolic analysis, the Rayleigh-Ritz problems) with the other's kernels. The rank's meshes are taken
    # biggest first from one queue.
    workers = max(1, int(os.environ.get("ME_BATCH_WORKERS", "3")))
    from concurrent.futures import ThreadPoolExecutor

    def solve_mesh(mesh):
        points, tets = mesh
        _, r = solve_once(points, tets, wl.bench_excitations(points), cfg)
        assert r.status == 0 and len(r.freqs) == 30
        return r.profile["kernel_launches"]

    def step():
        busy = time.perf_counter()
        if workers == 1:
            n = sum(solve_mesh(m) for m in meshes)
        else:
            with ThreadPoolExecutor(max_workers=workers) as pool:
                n = sum(pool.map(solve_mesh, meshes))
        return n, time.perf_counter() - busy

    for _ in range(warmup):
        step()
    launches, total, busy_s = 0, 0.0, 0.0
    with ClockSampler(local) as clocks:
        for _ in range(steps):
            ranks.barrier()
            t0 = time.perf_counter()
            n, busy = step()
            launches += n
            busy_s += busy
            total += ranks.reduce([time.perf_counter() - t0], "max")[0]
    n_launches = int(ranks.reduce([float(launches)], "sum")[0])
    busy_all = ranks.reduce([busy_s / steps], "sum")[0]
    if rank != 0:
        return None
    sec = total / steps
    return {
        "metric": "batch modal solve meshes/s", "value": len(dims) / sec, "unit": "meshes/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BASELINE.json configs[3]: 64 Kuhn-block meshes, {6 * dims[0] ** 3:,}..{6 * dims[-1] ** 3:,} tets (log-spaced), P1, lowest 30 modes each", "tets_total": int(sum(6 * d ** 3 for d in dims)),
                   "parallelism": f"meshes dealt biggest-first over {world} GPU(s) by tets^(4/3), no collective on the data path; {workers} solve(s) in flight per GPU (host threads, a stream per solve)"},
        "seconds_per_batch": sec, "tets_per_second": float(sum(6 * d ** 3 for d in dims)) / sec, "load_balance": busy_all / (world * sec),
        "e2e": {"value": len(dims) / sec, "unit": "meshes/s", "h2d_bytes_per_step": int(sum(p.nbytes + t.nbytes for p, t in meshes)), "d2h_bytes_per_step": 0,
                "note": "each solve is an me_modal_solve call: host mesh in, host modal model out"},
        "gpu_launches": n_launches, "clocks": clocks.summary(),
    }


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
    ap.add_argument("--no-parity", action="store_true", help="skip the parity report of the resonator line (it renders a 16-voice slice on the host twice: ~10 s)")
    ap.add_argument("--voices", type=int, default=0, help="resonator: experiment with another voice count (the bench line then names it; default = BASELINE.json configs[4])")
    ap.add_argument("--render-path", default="auto", choices=["auto", "loop", "tensor"], help="resonator: kernels of the free-running bank (auto = tensor-core form for this workload)")
    ap.add_argument("--workload", default="all", choices=["all", "resonator", "solve", "batch"],
                    help="all (default): the resonator line (configs[4], the metric quoted at 1/2/4/8 GPUs) carrying the `solve` (configs[2]) and `batch` (configs[3]) records; or one of them alone")
    ap.add_argument("--solve-steps", type=int, default=2, help="timed me_modal_solve calls of the solve record (after one warm-up call)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_solve_reference(args) if args.workload in ("solve", "batch") else run_reference(args)
    ranks = Ranks(args)
    try:
        if args.workload == "solve":
            line = run_solve(args, ranks, args.steps if args.steps != 5 else args.solve_steps, max(1, min(args.warmup, 3)), not args.no_cpu_baseline)
        elif args.workload == "batch":
            line = run_batch(args, ranks, args.steps if args.steps != 5 else 1, 1 if args.warmup else 0)
        else:
            line = run_resonator(args, ranks)
            if args.workload == "all":
                # The other half of BASELINE.json's metric and its sharded batch, in the line the driver parses.
                solve = run_solve(args, ranks, args.solve_steps, 1, not args.no_cpu_baseline)
                batch = run_batch(args, ranks, 1, 1)  # (one warm-up batch: the first concurrent solves of a process pay for its memory pool)
                if ranks.rank == 0:
                    line["solve"], line["batch"] = solve, batch
        if ranks.rank == 0:
            print(json.dumps(line))
    finally:
        ranks.close()


if __name__ == "__main__":
    main()
