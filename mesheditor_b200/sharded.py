"""One resonator bank spread over the GPUs of a node: one process per GPU, objects dealt between the ranks the way the
reference deals them between its render threads, and the per-rank mono mixes summed with ONE all-reduce per render
(NCCL over NVLink through torch.distributed).

Reference: ModalRenderPool / DealObjects / the fixed-order sum of the renderers' buffers (src/audio/ModalAudio.cpp:189-273,
430-461, 553-555). A renderer there is a thread with its own Out buffer; here it is a rank with its own ModalBank on its own
B200. The deal is the reference's (me_deal_objects: cost = modes x (1 + voices), heaviest first, ties by object, onto the
least-loaded renderer), computed identically on every rank from the same object list, so no rank ever talks to another
except in the final sum. The sum order differs from the reference's fixed renderer order; like the reference's own 1-vs-4
thread test (tests/ModalRenderTest.cpp:40-49) the result is held to 1e-5 of peak against the single-device render.

Every rank calls every method with the same arguments (SPMD), exactly as every rank of a torchrun job runs the same script.
"""
from __future__ import annotations

import numpy as np

from ._lib import MeModalEvent, check, lib
from .audio import ModalBank


def deal_objects(costs, n_renderers):
    """DealObjects (ModalAudio.cpp:430-461) -> (owner[i], slot of object i inside its owner's bank)."""
    c = np.ascontiguousarray(costs, np.uint64)
    owner, local = np.zeros(len(c), np.uint32), np.zeros(len(c), np.uint32)
    check(lib().me_deal_objects(c.ctypes.data, len(c), int(n_renderers), owner.ctypes.data, local.ctypes.data))
    return owner, local


class ShardedModalBank:
    """The ModalBank interface over `world` ranks. `group` is a torch.distributed process group (None = the default group);
    with world == 1 no collective is issued and torch is not needed at all.

    add_object()/add_modes() only record the object: which rank owns it is decided at install(), when all costs are known
    (the reference deals at render time for the same reason). `voices` is the number of sustained voices the caller expects
    on the object (SurfaceVoiceCount in the reference's cost); strikes alone leave it 0.
    """

    kind = "cuda-sharded"

    def __init__(self, sample_rate=48000.0, device=0, rank=0, world=1, group=None, bank_factory=None, all_reduce=None):
        self.rank, self.world, self.group = int(rank), int(world), group
        self.local = (bank_factory or ModalBank)(sample_rate, device)
        self._all_reduce = all_reduce  # tests inject one; default: torch.distributed.all_reduce on the device buffer
        self._objects, self._costs = [], []
        self.owner = np.zeros(0, np.uint32)
        self.local_slot = np.zeros(0, np.uint32)

    # --- building ---------------------------------------------------------------------------------------------------
    def add_object(self, freqs, t60s, shapes, positions, indices, out_gain=1.0, radius_scale=1.0, voices=0):
        freqs = np.ascontiguousarray(freqs, np.float32)
        self._objects.append((freqs, t60s, shapes, positions, indices, out_gain, radius_scale))
        # TunedModeCount is what the reference weighs (muted trailing modes cost nothing): frequencies TuneModalObject mutes.
        nyquist = 0.5 * self.local.sample_rate - 1 if hasattr(self.local, "sample_rate") else np.inf
        t = np.ascontiguousarray(t60s, np.float32)[:len(freqs)]
        audible = np.nonzero(np.isfinite(freqs[:len(t)]) & np.isfinite(t) & (freqs[:len(t)] > 0) & (freqs[:len(t)] < nyquist) & (t > 0))[0]
        tuned = int(audible[-1]) + 1 if len(audible) else 0
        self._costs.append(tuned * (1 + int(voices)))
        return len(self._objects) - 1

    def add_modes(self, modes, out_gain=1.0, radius_scale=1.0, voices=0):
        return self.add_object(modes["freqs"], modes["t60s"], modes["shapes"], modes["positions"], modes["indices"], out_gain, radius_scale, voices)

    def install(self, discard_frames=512):
        """Deal, build this rank's bank from the objects it owns (in bank order, ModalAudio.cpp:459), install it."""
        if self.owner.size != len(self._objects):
            first_new = self.owner.size
            if first_new:
                raise RuntimeError("objects were added after install(): build a new ShardedModalBank (the reference rebuilds its bank too)")
            self.owner, self.local_slot = deal_objects(self._costs, self.world)
            for i, obj in enumerate(self._objects):
                if self.owner[i] == self.rank:
                    slot = self.local.add_object(*obj)
                    assert slot == self.local_slot[i]
        self.local.install(discard_frames)

    def object_count(self):
        return len(self._objects)

    def owned(self):
        return [i for i in range(len(self._objects)) if self.owner[i] == self.rank]

    def _mine(self, obj):
        return obj < len(self.owner) and self.owner[obj] == self.rank

    def tune(self, obj, freqs, t60s, radius_scale=1.0):
        if self._mine(obj):
            self.local.tune(int(self.local_slot[obj]), freqs, t60s, radius_scale)

    def set_gain(self, obj, out_gain, listener_gain=1.0):
        if self._mine(obj):
            self.local.set_gain(int(self.local_slot[obj]), out_gain, listener_gain)

    def set_render_path(self, path):
        self.local.set_render_path(path)

    def set_time_segments(self, n):
        self.local.set_time_segments(n)

    # --- events -----------------------------------------------------------------------------------------------------
    def _localise(self, ev):
        """The event re-aimed at the owner's slot, or None when another rank owns the object (or nobody: DrainEvents :71)."""
        if not self._mine(ev.object):
            return None
        return MeModalEvent(ev.kind, int(self.local_slot[ev.object]), ev.ex_pos, ev.jx, ev.jy, ev.jz, ev.pulse_step, ev.pulse_gamma, ev.accel_amp, ev.click_b0, ev.click_a1, ev.click_a2)

    def enqueue(self, ev):
        local = self._localise(ev)
        return True if local is None else self.local.enqueue(local)

    def route_events(self, events, frames):
        """The slice of a global timeline this rank renders: (MeModalEvent array, frames, count), slots localised."""
        frames = np.asarray(frames, np.uint64)
        keep = [(self._localise(e), f) for e, f in zip(events, frames)]
        keep = [(e, f) for e, f in keep if e is not None]
        return ModalBank.pack_events([e for e, _ in keep], np.asarray([f for _, f in keep], np.uint64))

    # --- rendering --------------------------------------------------------------------------------------------------
    def _reduce_host(self, mix):
        if self.world == 1:
            return mix
        if self._all_reduce is not None:
            return self._all_reduce(mix)
        import torch
        import torch.distributed as dist

        t = torch.from_numpy(mix)
        if dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def render(self, out):
        """RenderModal on every rank + the sum of the renderers' buffers (:553-555): ADDS the full mix into `out` on every rank."""
        mix = np.zeros(out.size, np.float32)
        self.local.render(mix)
        out += self._reduce_host(mix)

    def render_blocks(self, blocks, frames=512):
        out = np.zeros(blocks * frames, np.float32)
        for b in range(blocks):
            self.render(out[b * frames:(b + 1) * frames])
        return out

    def render_offline(self, events, frames, total_frames, block_frames=512, routed=None):
        """Offline render of a global timeline; returns the full mix (host) on every rank."""
        routed = routed if routed is not None else self.route_events(events, frames)
        mix = self.local.render_offline(routed, None, total_frames, block_frames)  # a packed (array, frames, count) timeline
        return self._reduce_host(mix)

    def render_offline_device(self, routed, total_frames, block_frames, out, stream_ptr=None):
        """The throughput path: this rank's share rendered into the torch CUDA tensor `out` on `stream_ptr` (no host copy),
        then one NCCL all-reduce of the mono mix on torch's current stream (the caller makes that the same stream)."""
        self.local.render_offline_device(routed, None, total_frames, block_frames, out.data_ptr(), stream_ptr)
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(out, group=self.group)
        return out

    # --- introspection ----------------------------------------------------------------------------------------------
    def stats(self):
        return self.local.stats()

    def object_status(self, obj):
        """LiveModeCount / Ringing of a global object, on the rank that owns it (None elsewhere)."""
        return self.local.object_status(int(self.local_slot[obj])) if self._mine(obj) else None
