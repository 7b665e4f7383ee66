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
    assert C.sizeof(GxConfig) == 36 * 4 + 14 * 8
    assert GxConfig.pad_.offset == 35 * 4 and GxConfig.dx.offset == 36 * 4 and GxConfig.tsc.offset == 36 * 4 + 8 * 8 and GxConfig.mu.offset == 36 * 4 + 13 * 8   # no implicit padding
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


def test_product_path_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under guacho_b200/ may import, link or name it, and the shared
    library must not depend on it (ldd) or carry its symbols."""
    import subprocess
    pkg = os.path.join(ROOT, "guacho_b200")
    offenders = []
    for dirpath, _dirs, files in os.walk(pkg):
        if "build" in os.path.basename(dirpath):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".f90")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"#include[^\n]*oracle|dlopen[^\n]*oracle|import[^\n]*oracle|from\s+tests|import\s+tests|libguacho_oracle|\borc_[a-z_]+\(", txt):
                    offenders.append(os.path.join(dirpath, f))
    assert offenders == []
    if not os.path.exists(gxlib.LIB_PATH):
        from guacho_b200.build import build_library
        build_library()
    deps = subprocess.run(["ldd", gxlib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps
    syms = subprocess.run(["nm", "-D", "--defined-only", gxlib.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in syms
    exported = sorted(set(re.findall(r"\bT (gx_\w+)", syms)))
    assert exported == sorted(gxlib.EXPORTS)          # nothing but the declared C ABI is exported (-fvisibility=hidden)


def test_every_replaced_interface_cites_the_reference():
    """include/guacho_gx.h: each entry point that replaces a reference interface says which one (file:line)."""
    hdr = open(os.path.join(ROOT, "include", "guacho_gx.h")).read()
    must_cite = ["gx_create", "gx_set_state", "gx_set_time", "gx_get_timestep", "gx_tstep", "gx_run", "gx_get_state",
                 "gx_set_gravity_points", "gx_set_wind_spheres", "gx_register_bc_hook", "gx_comm_unique_id"]
    for name in must_cite:
        pos = hdr.index("GX_API int " + name + "(")
        prev = hdr.rfind("GX_API", 0, pos)              # everything since the previous entry point: comment (+ typedefs)
        comment = hdr[max(prev, 0):pos]
        assert re.search(r"\.f90:\d+", comment), f"{name}: no reference file:line in its comment"
