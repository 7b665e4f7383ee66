"""ctypes binding of libguacho_gx.so — the same C ABI the Fortran host binds through
ISO_C_BINDING (guacho_b200/fortran/guacho_gpu.f90, include/guacho_gx.h).

There is deliberately no fallback: if the library is missing or fails to load, or no
CUDA device is present, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

from .config import GxConfig

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GUACHO_GX_LIB") or os.path.join(HERE, "libguacho_gx.so")   # override: tuning variants only

GX_OK = 0
ERRORS = {-1: "GX_EINVAL", -2: "GX_ENODEVICE", -3: "GX_ECUDA", -4: "GX_ENOMEM",
          -5: "GX_EUNSUPPORTED", -6: "GX_ESTATE", -7: "GX_ECOMM", -8: "GX_ENUMERIC"}

# every symbol include/guacho_gx.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "gx_create", "gx_destroy", "gx_set_background", "gx_set_state", "gx_set_time", "gx_get_timestep", "gx_tstep", "gx_run",
    "gx_get_state", "gx_get_up", "gx_set_gravity_points", "gx_set_wind_spheres", "gx_register_host_bc", "gx_register_bc_hook",
    "gx_register_host_source", "gx_tc_info",
    "gx_comm_unique_id", "gx_comm_attach", "gx_last_error", "gx_launch_count", "gx_last_elapsed_ms",
    "gx_kernel_time_ms", "gx_set_profiling", "gx_build_info", "gx_riemann_flux",
)

KERNEL_CLASSES = ("flux", "update", "efield", "prim", "bc", "xpose", "visc", "stage1", "stage2", "bupdate", "tcond")


class GxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class WindSphere(C.Structure):
    _fields_ = [("xc", C.c_double), ("yc", C.c_double), ("zc", C.c_double), ("radius", C.c_double),
                ("vwind", C.c_double), ("dens", C.c_double), ("tfac", C.c_double), ("temp", C.c_double),
                ("vbx", C.c_double), ("vby", C.c_double), ("vbz", C.c_double),
                ("bdip", C.c_double), ("pas", C.c_double * 4)]


HOST_BC_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int32, C.c_void_p)
BC_HOOK_FN = C.CFUNCTYPE(None, C.c_int32, C.c_double, C.c_void_p)
HOST_SOURCE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)

_lib = None


def load() -> C.CDLL:
    """Load libguacho_gx.so (built in-tree by guacho_b200.build).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build it with `python -m guacho_b200.build` "
            "(the hydro/MHD step has no CPU or PyTorch fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    L.gx_create.argtypes = [C.POINTER(GxConfig), C.POINTER(vp)]
    L.gx_destroy.argtypes = [vp]
    L.gx_set_background.argtypes = [vp, dp]
    L.gx_set_state.argtypes = [vp, dp]
    L.gx_set_time.argtypes = [vp, C.c_double]
    L.gx_get_timestep.argtypes = [vp, C.c_int32, C.c_int32, C.c_double, C.c_double, dp, C.POINTER(C.c_int32)]
    L.gx_tstep.argtypes = [vp, C.c_double]
    L.gx_run.argtypes = [vp, C.c_int32, C.c_int32, dp, C.POINTER(C.c_int32), dp]
    L.gx_get_state.argtypes = [vp, dp, dp, dp]
    L.gx_get_up.argtypes = [vp, dp]
    L.gx_set_gravity_points.argtypes = [vp, C.c_int32, dp, dp]
    L.gx_set_wind_spheres.argtypes = [vp, C.c_int32, C.POINTER(WindSphere)]
    L.gx_register_host_bc.argtypes = [vp, HOST_BC_FN, vp]
    L.gx_register_bc_hook.argtypes = [vp, BC_HOOK_FN, vp]
    L.gx_register_host_source.argtypes = [vp, HOST_SOURCE_FN, vp]
    L.gx_tc_info.argtypes = [vp, dp, C.POINTER(C.c_int32)]
    L.gx_comm_unique_id.argtypes = [vp, C.c_int32]
    L.gx_comm_attach.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32]
    L.gx_last_error.restype = C.c_char_p
    L.gx_launch_count.restype = C.c_int64
    L.gx_launch_count.argtypes = [vp]
    L.gx_last_elapsed_ms.restype = C.c_double
    L.gx_last_elapsed_ms.argtypes = [vp]
    L.gx_kernel_time_ms.argtypes = [vp, C.c_int32, dp, C.POINTER(C.c_int64)]
    L.gx_set_profiling.argtypes = [vp, C.c_int32]
    L.gx_build_info.restype = C.c_char_p
    L.gx_riemann_flux.argtypes = [C.POINTER(GxConfig), C.c_int32, dp, dp, dp, C.POINTER(C.c_int32)]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int:   # default: int status
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != GX_OK:
        raise GxError(rc, load().gx_last_error().decode(errors="replace"))
