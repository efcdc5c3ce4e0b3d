"""gdkvm_b200 -- B200-native GDKVM memory op (LKVA readout + Gated Delta Rule state update).

Only the hot path of wangrui2025/GDKVM named by BASELINE.json ``north_star`` lives here:
``csrc/`` (hand-written sm_100a CUDA behind the C ABI of ``include/gdkvm_gdr.h``) and the
host-side mirror of the reference memory module's call surface.
"""
from .ops import chunk_gated_delta_rule, gdr_lkva, gdr_lkva_out, gdr_lkva_varlen, gdr_lkva_varlen_out, check_inputs, l2norm, launch_count, plan, plan_reason, plan_segments, plan_units, qkvgb_project, qkvgb_project_reference, train_unsupported_reason  # noqa: F401
from .memory import GDRMemory  # noqa: F401
from .model import GDKVMSkeleton  # noqa: F401
from ._cabi import FLAG_FLAT_CHUNKS, FLAG_FORCE_CHUNKED, FLAG_FORCE_RECURRENT, FLAG_FRAME_CHUNKS, FLAG_SEGMENTS  # noqa: F401

__version__ = "0.1.0"
