"""ctypes loader for libvbhem_b200.so (the C ABI of include/vbhem_b200.h).

The library is built in-tree by ``csrc/build.sh`` (``__graft_entry__.build()``).  There
is no Python or CPU implementation behind it: if the shared object is missing,
or no CUDA device is visible, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VBHEM_LIB: another build of the same library (e.g. the checked build with device-side assertions,
# scripts/checked_build.sh); never a different implementation
SO_PATH = os.environ.get("VBHEM_LIB") or os.path.join(HERE, "libvbhem_b200.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/vbhem_b200.h one to one
SIGNATURES = {
    "vbhem_version": (C.c_int, []),
    "vbhem_last_error": (C.c_char_p, []),
    "vbhem_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "vbhem_pack_base": (C.c_int, [C.c_int, C.c_int, ip, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp,
                                  C.POINTER(vp)]),
    "vbhem_free_base": (C.c_int, [vp]),
    "vbhem_repack_base": (C.c_int, [vp, ip, dp, dp, dp, dp, dp]),
    "vbhem_set_stream": (C.c_int, [vp, vp]),
    "vbhem_synchronize": (C.c_int, [vp]),
    "vbhem_set_kernel": (C.c_int, [vp, C.c_int]),
    "vbhem_estep_pairs": (C.c_int, [vp, C.c_int, ip, C.c_int, C.c_int, dp, dp, dp, dp, dp,
                                    dp, dp, dp, dp, dp, dp]),
    "vbhem_stats_size": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vbhem_stats_offsets": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "vbhem_estep": (C.c_int, [vp, C.c_int, ip, C.c_int, ip, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp,
                              dp, dp, dp]),
    "vbhem_estep_hyp": (C.c_int, [vp, C.c_int, ip, C.c_int, ip, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp,
                                  dp, dp, dp, dp]),
    "vbhem_estep_stream": (C.c_int, [vp, ip, dp, dp, dp, dp, dp, C.c_int, C.c_int, ip, C.c_int, ip, C.c_int, C.c_int,
                                     dp, dp, dp, dp, dp, dp, dp, dp, dp]),
    "vbhem_estep_hyp_stream": (C.c_int, [vp, ip, dp, dp, dp, dp, dp, C.c_int, C.c_int, ip, C.c_int, ip, C.c_int,
                                         C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]),
    "vbhem_estep_dev": (C.c_int, [vp, C.c_int, ip, C.c_int, ip, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, vp]),
    "vbhem_estep_hyp_dev": (C.c_int, [vp, C.c_int, ip, C.c_int, ip, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp,
                                      vp, dp]),
    "vbhem_fetch": (C.c_int, [vp, dp, dp]),
    "vbhem_em_begin": (C.c_int, [vp, C.c_int, ip, C.c_int, ip, C.c_int, C.c_double, C.c_double, C.c_double,
                                 C.c_double, C.c_double, dp, dp, C.c_int, dp, dp, dp, dp, dp, dp, dp]),
    "vbhem_em_begin_baseem": (C.c_int, [vp, C.c_int, ip, C.c_int, ip, C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_double, dp, dp, C.c_int, ip, ip, dp, C.c_double]),
    "vbhem_em_estep": (C.c_int, [vp, C.c_int, vp]),
    "vbhem_em_lower_bound": (C.c_int, [vp, vp, dp]),
    "vbhem_em_mstep": (C.c_int, [vp, vp]),
    "vbhem_em_mstep_masked": (C.c_int, [vp, vp, C.POINTER(C.c_ubyte)]),
    "vbhem_em_run_groups": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, dp, C.c_int, C.POINTER(C.c_int),
                                      C.POINTER(C.c_int)]),
    "vbhem_em_fetch": (C.c_int, [vp, dp, dp, dp, dp, dp, dp, dp]),
    "vbhem_em_set_group_priors": (C.c_int, [vp, dp, dp, dp, dp, dp, dp, dp, C.c_int]),
    "vbhem_em_lb_deriv": (C.c_int, [vp, dp]),
    "vbhem_em_run_groups_deriv": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, dp, C.c_int, C.POINTER(C.c_int),
                                            C.POINTER(C.c_int), dp]),
    "vbhem_em_run": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vbhem_finalize_stats": (C.c_int, [vp, C.c_int, ip, C.c_int, C.c_int, dp]),
    "vbhem_finalize_stats_dev": (C.c_int, [vp, C.c_int, ip, C.c_int, C.c_int, vp]),
    "vbhmm_pack_data": (C.c_int, [C.c_int, C.c_int, C.c_int, ip, C.POINTER(C.c_longlong), dp, C.POINTER(vp)]),
    "vbhmm_free_data": (C.c_int, [vp]),
    "vbhmm_data_shape": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(C.c_longlong)]),
    "vbhmm_fb": (C.c_int, [vp, C.c_int, dp, dp, dp, dp, dp, C.c_double, dp, dp, dp, dp, dp, dp]),
    "vbhmm_fb_dev": (C.c_int, [vp, C.c_int, dp, dp, dp, dp, dp, C.c_double, dp, dp]),
    "vbhmm_fb_fetch": (C.c_int, [vp, dp, dp, dp, dp]),
    "vbhmm_fb_stats": (C.c_int, [vp, C.c_int, dp, dp, dp, dp, dp, C.c_double, dp, dp, dp, dp, dp, dp, dp, dp]),
    "vbhmm_fb_last_ms": (C.c_int, [vp, dp]),
    "vbhmm_batch_begin": (C.c_int, [vp, C.c_int, ip, ip, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, dp,
                                    dp, C.c_int, dp, dp, dp, dp, dp, dp, C.c_int, C.POINTER(vp)]),
    "vbhmm_batch_run": (C.c_int, [vp, C.c_double, C.POINTER(C.c_int)]),
    "vbhmm_batch_fetch": (C.c_int, [vp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, ip, ip]),
    "vbhmm_batch_last_ms": (C.c_int, [vp, dp]),
    "vbhmm_batch_free": (C.c_int, [vp]),
    "vbhmm_gmm_fit": (C.c_int, [vp, C.c_int, ip, ip, C.c_int, ip, ip, C.c_int, C.c_double, C.c_double, dp, dp, dp, dp,
                                ip]),
    "vbhem_multi_unique_id": (C.c_int, [C.POINTER(C.c_ubyte)]),
    "vbhem_multi_pack_base": (C.c_int, [C.c_int, ip, C.c_int, C.c_int, C.POINTER(C.c_ubyte), C.c_int, ip, C.c_int,
                                        C.c_int, C.c_int, dp, dp, dp, dp, dp, C.POINTER(vp)]),
    "vbhem_multi_free": (C.c_int, [vp]),
    "vbhem_multi_peer_export": (C.c_int, [vp, C.POINTER(C.c_ubyte)]),
    "vbhem_multi_peer_import": (C.c_int, [vp, C.POINTER(C.c_ubyte)]),
    "vbhem_multi_allreduce_mode": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "vbhem_multi_repack_base": (C.c_int, [vp, ip, dp, dp, dp, dp, dp]),
    "vbhem_multi_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vbhem_multi_shard": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(vp)]),
    "vbhem_multi_set_kernel": (C.c_int, [vp, C.c_int]),
    "vbhem_multi_estep": (C.c_int, [vp, C.c_int, ip, C.c_int, ip, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp,
                                    dp, dp, dp]),
    "vbhem_multi_estep_hyp": (C.c_int, [vp, C.c_int, ip, C.c_int, ip, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp,
                                        dp, dp, dp, dp]),
    "vbhem_multi_estep_hyp_stream": (C.c_int, [vp, ip, dp, dp, dp, dp, dp, C.c_int, C.c_int, ip, C.c_int, ip, C.c_int,
                                               C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]),
    "vbhem_multi_estep_hyp_dev": (C.c_int, [vp, C.c_int, ip, C.c_int, ip, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp,
                                            C.c_int]),
    "vbhem_multi_synchronize": (C.c_int, [vp]),
    "vbhem_multi_align": (C.c_int, [vp]),
    "vbhem_multi_stats": (C.c_int, [vp, dp]),
    "vbhem_multi_fetch": (C.c_int, [vp, dp, dp]),
    "vbhem_multi_finalize_stats": (C.c_int, [vp, C.c_int, ip, C.c_int, C.c_int, dp]),
    "vbhem_multi_em_begin": (C.c_int, [vp, C.c_int, ip, C.c_int, ip, C.c_int, C.c_double, C.c_double, C.c_double,
                                       C.c_double, C.c_double, dp, dp, C.c_int, dp, dp, dp, dp, dp, dp, dp]),
    "vbhem_multi_em_run": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vbhem_multi_em_run_groups": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, dp, C.c_int, C.POINTER(C.c_int),
                                            C.POINTER(C.c_int)]),
    "vbhem_multi_em_step": (C.c_int, [vp, C.c_int, dp]),
    "vbhem_multi_em_fetch": (C.c_int, [vp, dp, dp, dp, dp, dp, dp, dp]),
    "vbhem_multi_last_ms": (C.c_int, [vp, dp]),
    "vbhem_multi_counters": (C.c_int, [vp, C.POINTER(C.c_longlong)]),
    "vbhem_set_profile": (C.c_int, [vp, C.c_int]),
    "vbhem_last_timings": (C.c_int, [vp, dp]),
    "vbhem_measure_fp64_peak": (C.c_int, [C.c_int, dp]),
    "vbhem_last_counters": (C.c_int, [vp, C.POINTER(C.c_longlong)]),
}

_lib = None


class VbhemError(RuntimeError):
    """A VBHEM_E* status from the C ABI (code, message)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.msg = msg


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is not built: run __graft_entry__.build() (csrc/build.sh). "
                              "There is no CPU fallback for the E-step.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise VbhemError(rc, lib().vbhem_last_error().decode())
