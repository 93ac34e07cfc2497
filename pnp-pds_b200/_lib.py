"""ctypes binding of libpnp_pds.so (the C ABI declared in include/pnp_pds.h).

There is no CPU fallback: if the shared library cannot be loaded, or no B200 is visible when a
handle is created, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

_lib = None

METHODS = {"A": 0, "B": 1, "C": 2, "FBS": 3, "RED": 4, "ADMM_B2": 5, "ADMM_C": 6, "RED_C": 7, "TV_A": 8, "TV_B3": 9, "TV_FBS": 10}
DEG_OPS = {"Id": 0, "blur": 1, "random_sampling": 2}
CONV_ENGINES = {"tcgen05": 0, "simt": 1}
TRACE_WIDTH = 5
PROF_CATS = ("primal", "dual", "l1ball", "conv_first", "conv_mid", "conv_last")


class PdsConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("batch", "channels", "height", "width", "method", "deg_op", "max_iter",
                                         "reserved", "device", "denoiser_chunk")]


class PdsItemParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("gamma1", "gamma2", "epsilon", "eta", "lam", "alpha")]


class PdsUnetConfig(C.Structure):
    _fields_ = [("batch", C.c_int32), ("in_nc", C.c_int32), ("out_nc", C.c_int32), ("nc", C.c_int32 * 4), ("nb", C.c_int32),
                ("height", C.c_int32), ("width", C.c_int32), ("device", C.c_int32)]


class PdsError(RuntimeError):
    pass


def _declare(lib):
    vp, f, i, u, sz = C.c_void_p, C.c_float, C.c_int, C.c_uint, C.c_size_t
    sig = {
        "pds_last_error": (C.c_char_p, []),
        "pds_abi_version": (i, []),
        "pds_device_count": (i, []),
        "pds_create": (i, [C.POINTER(PdsConfig), C.POINTER(vp)]),
        "pds_destroy": (i, [vp]),
        "pds_set_blur_kernel": (i, [vp, C.POINTER(C.c_double), i]),
        "pds_set_mask": (i, [vp, C.POINTER(C.c_uint8)]),
        "pds_set_item_params": (i, [vp, C.POINTER(PdsItemParams), i]),
        "pds_set_admm": (i, [vp, i, i, f]),
        "pds_set_ssim": (i, [vp, i]),
        "pds_load_dncnn": (i, [vp, vp, sz]),
        "pds_phi": (i, [vp, vp, vp, vp]),
        "pds_phi_adj": (i, [vp, vp, vp, vp]),
        "pds_proj_l2_ball": (i, [vp, vp, vp, f, vp, vp]),
        "pds_proj_l1_ball": (i, [vp, vp, f, vp, vp]),
        "pds_prox_gkl": (i, [vp, vp, vp, f, f, vp, vp]),
        "pds_dncnn_forward": (i, [vp, vp, vp, vp]),
        "pds_set_problem": (i, [vp, vp, vp, vp, vp]),
        "pds_run": (i, [vp, i, vp]),
        "pds_iterations_done": (i, [vp]),
        "pds_get_state": (i, [vp, vp, vp, vp, vp]),
        "pds_get_traces": (i, [vp, vp, sz, vp]),
        "pds_restore_host": (i, [vp, vp, vp, vp, i, vp, vp, vp, sz, vp]),
        "pds_profile_enable": (i, [vp, i]),
        "pds_profile_read": (i, [vp, vp, vp, i, vp]),
        "pds_kernel_launches": (C.c_longlong, [vp]),
        "pds_workspace_bytes": (sz, [vp]),
        "pds_unet_create": (i, [C.POINTER(PdsUnetConfig), C.POINTER(vp)]),
        "pds_unet_destroy": (i, [vp]),
        "pds_unet_blob_bytes": (sz, [C.POINTER(PdsUnetConfig)]),
        "pds_unet_load": (i, [vp, vp, sz]),
        "pds_unet_forward": (i, [vp, vp, vp, vp]),
        "pds_unet_kernel_launches": (C.c_longlong, [vp]),
        "pds_unet_workspace_bytes": (sz, [vp]),
        "pds_debug_set_tc_variant": (i, [vp, i]),
        "pds_debug_set_conv_engine": (i, [vp, i]),
        "pds_debug_chain_trace": (i, [vp, vp]),
        "pds_debug_body_kernel": (i, [vp, i]),
        "pds_debug_roll_band_rows": (i, [i, i, i, i]),
        "pds_debug_umma_probe": (i, [u, u, u, u, vp]),
        "pds_debug_tma_probe": (i, [vp, i, i, i, i, i, i, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTS = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load (building first if the .so is absent or stale and nvcc is present)."""
    global _lib, EXPORTS
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing:
        try:
            if _build.needs_build():
                _build.build()
        except Exception as e:  # no nvcc on this machine: fall through to loading what is there
            if not os.path.exists(path):
                raise PdsError(f"libpnp_pds.so is missing and could not be built: {e}") from e
    if not os.path.exists(path):
        raise PdsError(f"{path} not found — build it with `python __graft_entry__.py` (no CPU fallback exists)")
    lib = C.CDLL(path)
    EXPORTS = _declare(lib)
    if lib.pds_abi_version() != 1:
        raise PdsError("libpnp_pds ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().pds_last_error()
        raise PdsError(msg.decode() if msg else f"libpnp_pds error {rc}")
