"""The C-ABI library loads and exports every symbol include/guacho_gx.h declares
(no compute calls: this runs without a GPU)."""
import ctypes as C
import os
import re

from guacho_b200 import lib as gxlib
from guacho_b200.config import GxConfig, Params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "guacho_gx.h")).read()
    return sorted(set(re.findall(r"GX_API[^;(]*?\b(gx_\w+)\s*\(", hdr)))


def test_header_symbols_all_exported():
    if not os.path.exists(gxlib.LIB_PATH):
        from guacho_b200.build import build_library
        build_library()
    L = C.CDLL(gxlib.LIB_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in guacho_gx.h but not exported"
    assert set(names) == set(gxlib.EXPORTS)


def test_binding_loads_and_reports_build():
    L = gxlib.load()
    info = L.gx_build_info().decode()
    assert "sm_100a" in info


def test_config_struct_matches_header_size():
    # all int32 first (34 of them), then 9 doubles
    assert C.sizeof(GxConfig) == 34 * 4 + 9 * 8
    c = Params().to_c()
    assert c.struct_bytes == C.sizeof(GxConfig)


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a CUDA device gx_create must fail with GX_ENODEVICE (never compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        return
    L = gxlib.load()
    cfg = Params(nxtot=8, nytot=8, nztot=8).to_c()
    h = C.c_void_p()
    rc = L.gx_create(C.byref(cfg), C.byref(h))
    assert rc == -2, (rc, L.gx_last_error())
    assert b"no CPU fallback" in L.gx_last_error()
