"""ctypes binding of the C ABI declared in include/pfb_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc,
sm_100a).  There is no CPU fallback: if the library is missing, importing this
module raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PFB_B200_LIB selects another build of the same library (kernel experiments); the default is the in-tree build
LIB_PATH = os.environ.get("PFB_B200_LIB") or os.path.join(_HERE, "libpfb_b200.so")

EXPORTS = [
    "pfb_last_error", "pfb_version", "pfb_launch_count",
    "pfb_analysis_plan_create", "pfb_analysis_plan_destroy", "pfb_analysis_nblocks",
    "pfb_analysis_nchan_out", "pfb_analysis_execute", "pfb_analysis_execute_segment",
    "pfb_inverse_plan_create", "pfb_inverse_plan_destroy", "pfb_inverse_nblocks",
    "pfb_inverse_out_len", "pfb_inverse_execute", "pfb_inverse_set_scratch_bytes",
    "pfb_inverse_execute_stage", "pfb_inverse_set_profiling", "pfb_inverse_get_profile",
    "pfb_inverse_set_spectral_taper", "pfb_generate_test_vector", "pfb_purity_metrics",
    "pfb_analysis_execute_batched", "pfb_inverse_execute_batched", "pfb_generate_test_vector_strided",
    "pfb_generate_test_vectors", "pfb_unpack_payload", "pfb_pack_payload", "pfb_inverse_reserve", "pfb_inverse_execute_vectors", "pfb_spectrum_vectors", "pfb_purity_metrics_pairs", "pfb_analysis_execute_segment_peer", "pfb_roundtrip_execute", "pfb_roundtrip_execute_vectors",
]

PFB_DTYPE_C64, PFB_DTYPE_C128 = 0, 1
VARIANTS = {"polyphase_analysis": 0, "bunton": 0,
            "polyphase_analysis_padded": 1, "padded": 1,
            "polyphase_analysis_lowcbf": 2, "lowcbf": 2}


class PfbError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    c_ll, c_int, c_vp = ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p
    dp = ctypes.POINTER(ctypes.c_double)
    lib.pfb_last_error.restype = ctypes.c_char_p
    lib.pfb_version.restype = c_int
    lib.pfb_launch_count.restype = c_ll
    lib.pfb_analysis_plan_create.argtypes = [ctypes.POINTER(c_vp), c_int, c_int, c_int, c_int, dp, c_int, c_int, c_int]
    lib.pfb_analysis_plan_destroy.argtypes = [c_vp]
    lib.pfb_analysis_nblocks.argtypes = [c_vp, c_ll, c_int]
    lib.pfb_analysis_nblocks.restype = c_ll
    lib.pfb_analysis_nchan_out.argtypes = [c_vp]
    lib.pfb_analysis_execute.argtypes = [c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_vp]
    lib.pfb_analysis_execute_segment.argtypes = [c_vp, c_vp, c_ll, c_ll, c_vp, c_int, c_ll, c_ll, c_int, c_vp]
    lib.pfb_analysis_execute_segment_peer.argtypes = [c_vp, c_vp, c_ll, c_vp, c_ll, c_ll, c_vp, c_int, c_ll, c_ll, c_int, c_vp]
    lib.pfb_roundtrip_execute.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_vp]
    lib.pfb_roundtrip_execute_vectors.argtypes = [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_ll, c_vp]
    lib.pfb_inverse_plan_create.argtypes = [ctypes.POINTER(c_vp), c_int, c_int, c_int, c_int, c_int, dp, dp, c_int,
                                            c_int, c_int, c_int, c_int]
    lib.pfb_inverse_plan_destroy.argtypes = [c_vp]
    lib.pfb_inverse_nblocks.argtypes = [c_vp, c_ll]
    lib.pfb_inverse_nblocks.restype = c_ll
    lib.pfb_inverse_out_len.argtypes = [c_vp, c_ll]
    lib.pfb_inverse_out_len.restype = c_ll
    lib.pfb_inverse_execute.argtypes = [c_vp, c_vp, c_vp, c_int, c_ll, c_vp]
    lib.pfb_inverse_set_scratch_bytes.argtypes = [c_vp, c_ll]
    lib.pfb_inverse_execute_stage.argtypes = [c_vp, c_int, c_vp, c_vp, c_int, c_ll, c_vp]
    lib.pfb_inverse_set_spectral_taper.argtypes = [c_vp, dp, c_ll]
    c_d = ctypes.c_double
    lib.pfb_generate_test_vector.argtypes = [c_int, c_vp, c_ll, c_int, c_int, c_d, c_d, c_d, c_vp]
    lib.pfb_generate_test_vector_strided.argtypes = [c_int, c_vp, c_ll, c_ll, c_int, c_d, c_d, c_d, c_vp]
    lib.pfb_generate_test_vectors.argtypes = [c_int, c_vp, c_ll, c_int, c_int, c_ll, c_int, c_vp, c_vp]
    lib.pfb_purity_metrics.argtypes = [c_vp, c_int, c_int, c_vp]
    lib.pfb_analysis_execute_batched.argtypes = [c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_vp]
    lib.pfb_inverse_execute_batched.argtypes = [c_vp, c_vp, c_vp, c_int, c_int, c_ll, c_vp]
    lib.pfb_unpack_payload.argtypes = [c_vp, c_vp, c_ll, c_int, c_int, c_vp]
    lib.pfb_pack_payload.argtypes = [c_vp, c_vp, c_ll, c_int, c_int, c_d, c_vp]
    lib.pfb_inverse_execute_vectors.argtypes = [c_vp, c_vp, c_vp, c_int, c_int, c_ll, c_ll, c_vp]
    lib.pfb_purity_metrics_pairs.argtypes = [c_vp, c_int, c_int, c_vp]
    lib.pfb_spectrum_vectors.argtypes = [c_vp, c_ll, c_vp, c_vp, c_int, c_ll, c_vp]
    lib.pfb_inverse_reserve.argtypes = [c_vp, c_int, c_ll, c_int]
    lib.pfb_inverse_set_profiling.argtypes = [c_vp, c_int]
    lib.pfb_inverse_get_profile.argtypes = [c_vp, dp, ctypes.POINTER(c_ll)]
    return lib


lib = _load()


def check(rc: int):
    if rc != 0:
        raise PfbError(f"pfb_b200 error {rc}: {lib.pfb_last_error().decode()}")
