"""Build libguacho_gx.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m guacho_b200.build [--force]

The kernels are compiled twice: a bit-comparison build (-fmad=false, matches the
reference's no-FMA x86 arithmetic) and the production build (-fmad=true).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "libguacho_gx.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libguacho_gx.so cannot be built (there is no CPU fallback)")


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))] + [
        os.path.join(os.path.dirname(HERE), "include", "guacho_gx.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_library(force: bool = False, verbose: bool = False, extra=(), out: str = LIB) -> str:
    """`extra`/`out` build tuning variants (e.g. -DGX_FLUX_MINBLOCKS=6) next to the default library."""
    global BUILD
    if out == LIB and not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    if out != LIB:
        BUILD = os.path.join(HERE, "csrc", "build_" + os.path.basename(out).replace(".so", ""))
    os.makedirs(BUILD, exist_ok=True)
    jobs = [
        ("gx_kernels_strict.o", "gx_kernels.cu", ["-fmad=false", "-DGX_FLAVOUR_STRICT"]),
        ("gx_kernels_fast.o", "gx_kernels.cu", ["-fmad=true", "-DGX_FLAVOUR_FAST"]),
        ("gx_api.o", "gx_api.cu", ["-fmad=false"]),
    ]
    # fused stage kernels: one translation unit per (flavour, Riemann solver); HLLD first (largest)
    for solver in (4, 3, 2, 1):
        jobs.append((f"gx_stage_fast_{solver}.o", "gx_stage.cu", ["-fmad=true", "-DGX_FLAVOUR_FAST", f"-DGX_STAGE_SOLVER={solver}"]))
        jobs.append((f"gx_stage_strict_{solver}.o", "gx_stage.cu", ["-fmad=false", "-DGX_FLAVOUR_STRICT", f"-DGX_STAGE_SOLVER={solver}"]))
    if os.environ.get("GX_DEV_MINMOD_ONLY"):
        extra = list(extra) + ["-DGX_DEV_MINMOD_ONLY"]
    cmds = [[nvcc] + ARCH + COMMON + flags + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", os.path.join(BUILD, obj)]
            for obj, src, flags in jobs]
    njobs = max(1, min(len(cmds), os.cpu_count() or 4))
    running, failed = [], None
    pending = list(cmds)
    while pending or running:
        while pending and len(running) < njobs:
            cmd = pending.pop(0)
            running.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        cmd, p = running.pop(0)
        log, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(log)
        if p.returncode and failed is None:
            failed = cmd
    if failed:
        raise RuntimeError("nvcc failed: " + " ".join(failed))
    cmd = [nvcc] + ARCH + ["-shared", "-o", out] + [os.path.join(BUILD, j[0]) for j in jobs] + ["-ldl"]
    subprocess.check_call(cmd)
    return out


HOST_SRC = os.path.join(HERE, "host", "guacho_host.cpp")
HOST_BIN = os.path.join(HERE, "host", "guacho_host")


def build_host(force: bool = False) -> str:
    """The compiled host driver (mirror of src/main.f90 over the C ABI): g++ only, links libguacho_gx.so."""
    if not os.path.exists(LIB):
        build_library()
    if not force and os.path.exists(HOST_BIN) and os.path.getmtime(HOST_BIN) >= max(os.path.getmtime(HOST_SRC), os.path.getmtime(LIB)):
        return HOST_BIN
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", HOST_SRC, "-I" + os.path.join(os.path.dirname(HERE), "include"),
                           "-L" + HERE, "-lguacho_gx", "-Wl,-rpath,$ORIGIN/..", "-o", HOST_BIN])
    return HOST_BIN


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, extra=defs, out=(os.path.abspath(outs[0]) if outs else LIB)))
