"""mesheditor_b200 — B200-native linear modal analysis / synthesis (the hot path of khiner/MeshEditor).

The product is libme_modal.so: hand-written sm_100a CUDA kernels behind the C ABI in include/me_modal.h.
This package is only the ctypes face of that ABI. Importing it never imports anything under oracle/.
"""
from ._lib import LIB_PATH, MeError, MeModalEvent, MeRetune, lib  # noqa: F401
from .audio import ModalBank, impact_event, listener_gain, measure_fp32_fma_rate, modal_out_gain, monitor_frames, retune_modes, retuning, silence_event, uniform_scale_ratio, wav_bytes, wav_frames  # noqa: F401
from .modal import symbolic_analyse, Factor, FemSystem, ModalResult, material, measure_fp64_rate, mesh2modes, postprocess_modes, solver_config  # noqa: F401,E402
from .sharded import ShardedModalBank, deal_objects  # noqa: F401,E402
