"""ctypes wrapper around oracle/libguacho_oracle.so (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The oracle is a CPU restatement of the reference's
hydro/MHD step (see oracle/guacho_oracle.cpp); it emulates the MPI block
decomposition inside one process.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from guacho_b200.config import GxConfig, Params

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")

U, UP, PRIMIT, F, G, H, E, TEMP, PRIMIT0 = range(9)


def build_oracle(force: bool = False) -> None:
    need = force or not all(os.path.exists(os.path.join(ORACLE_DIR, n))
                            for n in ("libguacho_oracle.so", "libguacho_oracle_fast.so"))
    if not need:
        src = os.path.getmtime(os.path.join(ORACLE_DIR, "guacho_oracle.cpp"))
        need = any(os.path.getmtime(os.path.join(ORACLE_DIR, n)) < src
                   for n in ("libguacho_oracle.so", "libguacho_oracle_fast.so"))
    if need:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"] + (["-B"] if force else []))


_libs = {}


def load(fast: bool = False):
    name = "libguacho_oracle_fast.so" if fast else "libguacho_oracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(ORACLE_DIR, name)
    if not os.path.exists(path):
        build_oracle()
    L = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.POINTER(GxConfig)]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
    L.orc_num_blocks.argtypes = [C.c_void_p]
    L.orc_block_coords.argtypes = [C.c_void_p, C.c_int, ip]
    L.orc_block_neighbors.argtypes = [C.c_void_p, C.c_int, ip]
    L.orc_set_time.argtypes = [C.c_void_p, C.c_double]
    L.orc_error.argtypes = [C.c_void_p]
    L.orc_block_array_size.restype = C.c_int64
    L.orc_block_array_size.argtypes = [C.c_void_p, C.c_int]
    L.orc_get_block.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    L.orc_set_block.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    L.orc_gather_interior.argtypes = [C.c_void_p, C.c_int, dp]
    L.orc_scatter_u_with_ghosts.argtypes = [C.c_void_p, dp]
    L.orc_impose_ot.argtypes = [C.c_void_p, C.c_double]
    L.orc_init_exo.argtypes = [C.c_void_p] + [C.c_double] * 6
    L.orc_exo_initial_conditions.argtypes = [C.c_void_p]
    L.orc_exo_params.argtypes = [C.c_void_p, dp]
    for fn in ("orc_boundaryI", "orc_boundaryII", "orc_calcprim_u", "orc_calcprim_up", "orc_start", "orc_viscous_copy"):
        getattr(L, fn).argtypes = [C.c_void_p]
    L.orc_get_timestep.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, dp, ip]
    L.orc_tstep.argtypes = [C.c_void_p, C.c_double]
    L.orc_fluxes.argtypes = [C.c_void_p, C.c_int]
    L.orc_step.argtypes = [C.c_void_p, C.c_double]
    L.orc_run.restype = C.c_double
    L.orc_run.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, ip, dp]
    L.orc_u2prim.argtypes = [C.c_void_p, dp, dp, dp]
    L.orc_prim2u.argtypes = [C.c_void_p, dp, dp]
    L.orc_prim2f.argtypes = [C.c_void_p, dp, dp]
    L.orc_riemann.argtypes = [C.c_void_p, dp, dp, dp]
    L.orc_limiter.argtypes = [C.c_void_p, dp, dp, dp, dp]
    L.orc_average.restype = C.c_double
    L.orc_average.argtypes = [C.c_int, C.c_double, C.c_double]
    L.orc_cfast.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, dp]
    L.orc_cfastX.restype = C.c_double
    L.orc_cfastX.argtypes = [C.c_void_p, dp]
    L.orc_csound.restype = C.c_double
    L.orc_csound.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.orc_cool_atomic.argtypes = [C.c_void_p, C.c_double, dp]
    L.orc_cool_rate.restype = C.c_double
    L.orc_cool_rate.argtypes = [C.c_int, C.c_double]
    L.orc_cool_aloss.restype = C.c_double
    L.orc_cool_aloss.argtypes = [C.c_double] * 6
    L.orc_scatter_primit0_with_ghosts.argtypes = [C.c_void_p, dp]
    L.orc_riemann_split_all.argtypes = [C.c_void_p, dp, dp, dp, dp, dp]
    L.orc_tc_info.argtypes = [C.c_void_p, dp, ip]
    L.orc_thermal_conduction.argtypes = [C.c_void_p, C.c_double]
    L.orc_tc_superstep.restype = C.c_double
    L.orc_tc_superstep.argtypes = [C.c_int]
    L.orc_tc_substep.restype = C.c_double
    L.orc_tc_substep.argtypes = [C.c_int, C.c_int]
    L.orc_tc_st_steps.argtypes = [C.c_double, ip, dp]
    L.orc_tc_ksp.restype = C.c_double
    L.orc_tc_ksp.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    _libs[name] = L
    return L


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["F_CONTIGUOUS"] or a.ndim <= 1 and a.flags["C_CONTIGUOUS"], "need contiguous float64"
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    """One oracle instance = the whole MPI job of the reference (all blocks)."""

    def __init__(self, params: Params, fast: bool = False, threads: int = 1):
        params.validate()
        self.p = params
        self.L = load(fast)
        self._cfg = params.to_c()
        self.h = C.c_void_p(self.L.orc_create(C.byref(self._cfg)))
        self.L.orc_set_threads(self.h, threads)
        self.time = 0.0
        self.iter = 1

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- arrays ----
    def block_shape(self, which=U):
        p = self.p
        if which == TEMP:
            return (p.nx + 4, p.ny + 4, p.nz + 4)
        n1 = 3 if which == E else p.neq
        return (n1, p.nx + 4, p.ny + 4, p.nz + 4)

    def get_block(self, b: int, which=U) -> np.ndarray:
        out = np.zeros(self.block_shape(which), dtype=np.float64, order="F")
        self.L.orc_get_block(self.h, b, which, _dp(out))
        return out

    def set_block(self, b: int, arr: np.ndarray, which=U) -> None:
        a = np.asfortranarray(arr, dtype=np.float64)
        assert a.shape == self.block_shape(which)
        self.L.orc_set_block(self.h, b, which, _dp(a))

    def gather(self, which=U) -> np.ndarray:
        """Interior cells of all blocks as a global (n1, nxtot, nytot, nztot) array."""
        p = self.p
        n1 = 3 if which == E else p.neq
        out = np.zeros((n1, p.nxtot, p.nytot, p.nztot), dtype=np.float64, order="F")
        self.L.orc_gather_interior(self.h, which, _dp(out))
        return out

    def scatter_u(self, g: np.ndarray) -> None:
        """Global array with 2 ghost layers (neq, nxtot+4, nytot+4, nztot+4) -> every block's u
        (ghosts included), like initial_conditions() filling nxmin:nxmax."""
        p = self.p
        a = np.asfortranarray(g, dtype=np.float64)
        assert a.shape == (p.neq, p.nxtot + 4, p.nytot + 4, p.nztot + 4)
        self.L.orc_scatter_u_with_ghosts(self.h, _dp(a))

    def scatter_primit0(self, g: np.ndarray) -> None:
        """Background primitives of the split-all solvers (globals primit0, set by the host): global array with ghosts."""
        p = self.p
        a = np.asfortranarray(g, dtype=np.float64)
        assert a.shape == (p.neq, p.nxtot + 4, p.nytot + 4, p.nztot + 4)
        self.L.orc_scatter_primit0_with_ghosts(self.h, _dp(a))

    def riemann_split_all(self, pl, pr, p0l, p0r):
        a, b, c, d, out = self._vec(pl), self._vec(pr), self._vec(p0l), self._vec(p0r), np.zeros(16)
        self.L.orc_riemann_split_all(self.h, _dp(a), _dp(b), _dp(c), _dp(d), _dp(out))
        return out[:self.p.neq].copy()

    def coords(self, b: int):
        c = (C.c_int * 3)()
        self.L.orc_block_coords(self.h, b, c)
        return tuple(c)

    def neighbors(self, b: int):
        n = (C.c_int * 6)()
        self.L.orc_block_neighbors(self.h, b, n)
        return tuple(n)

    @property
    def nblocks(self):
        return self.L.orc_num_blocks(self.h)

    # ---- problems ----
    def impose_ot(self, rsc: float = 1.0):
        self.L.orc_impose_ot(self.h, rsc)

    # ---- the calls main.f90 makes ----
    def start(self):
        """boundaryI + calcprim(u, primit)  (main.f90:76-79)."""
        self.L.orc_start(self.h)

    def get_timestep(self, current_iter=None, n_iter=10, time=None, tprint=1e300):
        dt = C.c_double(0.0)
        dump = C.c_int(0)
        it = self.iter if current_iter is None else current_iter
        t = self.time if time is None else time
        self.L.orc_get_timestep(self.h, it, n_iter, t, tprint, C.byref(dt), C.byref(dump))
        return dt.value, bool(dump.value)

    def tstep(self, dt: float) -> int:
        self.L.orc_set_time(self.h, self.time)
        return self.L.orc_tstep(self.h, dt)

    def thermal_conduction(self, dt_cfl: float) -> None:
        """thermal_conduction() of src/thermal_cond.f90:690-768 alone (primit/Temp must be current: start())."""
        self.L.orc_thermal_conduction(self.h, dt_cfl)

    def tc_info(self):
        """(dt_cond [s], substeps) of the last thermal_conduction call — the reference's log line (:725)."""
        dt, n = C.c_double(0.0), C.c_int(0)
        self.L.orc_tc_info(self.h, C.byref(dt), C.byref(n))
        return dt.value, n.value

    def advance(self, nsteps: int, n_iter: int = 10, tprint: float = 1e300):
        """main.f90:94-125 without output."""
        dts = []
        for _ in range(nsteps):
            dt, _d = self.get_timestep(self.iter, n_iter, self.time, tprint)
            err = self.tstep(dt)
            if err:
                raise FloatingPointError("oracle: Riemann solver fell through all branches (reference would `stop`)")
            self.time += dt
            self.iter += 1
            dts.append(dt)
        return dts

    def run_timed(self, nsteps: int, n_iter: int = 10) -> float:
        t = C.c_double(self.time)
        it = C.c_int(self.iter)
        last = C.c_double(0.0)
        sec = self.L.orc_run(self.h, nsteps, n_iter, C.byref(t), C.byref(it), C.byref(last))
        self.time, self.iter = t.value, it.value
        return sec

    # ---- single-cell KAT entry points ----
    def _vec(self, v):
        a = np.zeros(16, dtype=np.float64)
        a[:len(v)] = v
        return a

    def u2prim(self, uu):
        a, out, T = self._vec(uu), np.zeros(16), C.c_double(0)
        self.L.orc_u2prim(self.h, _dp(a), _dp(out), C.byref(T))
        return out[:self.p.neq].copy(), T.value

    def prim2u(self, prim):
        a, out = self._vec(prim), np.zeros(16)
        self.L.orc_prim2u(self.h, _dp(a), _dp(out))
        return out[:self.p.neq].copy()

    def prim2f(self, prim):
        a, out = self._vec(prim), np.zeros(16)
        self.L.orc_prim2f(self.h, _dp(a), _dp(out))
        return out[:self.p.neq].copy()

    def riemann(self, pl, pr):
        a, b, out = self._vec(pl), self._vec(pr), np.zeros(16)
        err = self.L.orc_riemann(self.h, _dp(a), _dp(b), _dp(out))
        return out[:self.p.neq].copy(), err

    def limiter(self, pll, pl, pr, prr):
        a, b, c, d = self._vec(pll), self._vec(pl), self._vec(pr), self._vec(prr)
        self.L.orc_limiter(self.h, _dp(a), _dp(b), _dp(c), _dp(d))
        return b[:self.p.neq].copy(), c[:self.p.neq].copy()

    def cfast(self, p, d, bx, by, bz):
        out = np.zeros(3)
        self.L.orc_cfast(self.h, p, d, bx, by, bz, _dp(out))
        return out

    def cfastX(self, prim):
        return self.L.orc_cfastX(self.h, _dp(self._vec(prim)))

    def cool_atomic(self, dt_seconds, uu):
        """atomic(dt, uu, 1., 1.) of src/cooling_h.f90:259-371 on one cell -> new uu."""
        a = self._vec(uu)
        self.L.orc_cool_atomic(self.h, dt_seconds, _dp(a))
        return a[:self.p.neq].copy()

    def csound(self, p, d):
        return self.L.orc_csound(self.h, p, d)
