"""Host-side mirror of the reference's driver for the hot path.

:class:`Block` wraps one ``gx_solver`` (= one MPI rank / one GPU of the reference's
Cartesian decomposition) and exposes the calls ``src/main.f90`` makes into the step:
``initflow -> boundaryI -> calcprim`` (:meth:`set_state`), ``get_timestep``, ``tstep``,
and "state on the host before write_output" (:meth:`get_state`).  :class:`Simulation`
mirrors the time loop of ``src/main.f90:94-125``.

Everything numerical happens inside libguacho_gx.so; this module only marshals
numpy arrays in the reference layout ``(neq, nx+4, ny+4, nz+4)`` (Fortran order).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Sequence

import numpy as np

from . import lib as _lib
from .config import Params
from .lib import GxError, WindSphere, check, KERNEL_CLASSES


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(C.c_double))


def riemann_flux(params: Params, wl: np.ndarray, wr: np.ndarray):
    """prim2fhll / prim2fhllc / prim2fhlle / prim2fhlld (src/hll.f90:47, hllc.f90:44, hlle.f90:48, hlld.f90:48) for n interfaces
    on the device: wl, wr of shape (n, neqdyn), rotated primitive states -> (flux (n, neqdyn), err (n,))."""
    L = _lib.load()
    wl = np.ascontiguousarray(wl, dtype=np.float64)
    wr = np.ascontiguousarray(wr, dtype=np.float64)
    if wl.shape != wr.shape or wl.ndim != 2 or wl.shape[1] != params.neqdyn:
        raise ValueError(f"wl, wr must both have shape (n, {params.neqdyn})")
    n = wl.shape[0]
    ff = np.zeros_like(wl)
    err = np.zeros(n, dtype=np.int32)
    cfg = params.to_c((0, 0, 0))
    check(L.gx_riemann_flux(C.byref(cfg), n, _dp(wl), _dp(wr), _dp(ff), err.ctypes.data_as(C.POINTER(C.c_int32))))
    return ff, err


class Block:
    """One block of the domain on one GPU (reference: one MPI rank)."""

    def __init__(self, params: Params, coords: Sequence[int] = (0, 0, 0)):
        params.validate()
        self.p = params
        self.coords = tuple(int(c) for c in coords)
        self.L = _lib.load()
        self._cfg = params.to_c(self.coords)
        h = C.c_void_p()
        check(self.L.gx_create(C.byref(self._cfg), C.byref(h)))
        self.h = h
        self._host_bc_ref = None
        self._bc_hook_ref = None
        self._host_src_ref = None
        self._cb_error = None        # exception raised inside a user callback (ctypes would print and swallow it)

    # -- lifetime --
    def close(self) -> None:
        if getattr(self, "h", None):
            self.L.gx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- shapes --
    @property
    def shape(self):
        return self.p.block_shape()

    @property
    def rank(self) -> int:
        """Row-major rank <-> coords map of the reference (py/guacho_utils.py:104-118)."""
        cx, cy, cz = self.coords
        return (cx * self.p.MPI_NBY + cy) * self.p.MPI_NBZ + cz

    def empty_state(self) -> np.ndarray:
        return np.zeros(self.shape, dtype=np.float64, order="F")

    def set_background(self, primit0: np.ndarray) -> None:
        """primit0 of the split-all solver (src/globals.f90:42): background primitives, same shape as u; before set_state."""
        if primit0.shape != self.shape:
            raise ValueError(f"primit0 has shape {primit0.shape}, expected {self.shape}")
        a = np.asfortranarray(primit0, dtype=np.float64)
        self._check(self.L.gx_set_background(self.h, _dp(a)))

    # -- calls main.f90 makes --
    def set_state(self, u: np.ndarray) -> None:
        """initflow -> boundaryI -> calcprim (main.f90:73-79)."""
        if u.shape != self.shape:
            raise ValueError(f"u has shape {u.shape}, expected {self.shape}")
        a = np.asfortranarray(u, dtype=np.float64)
        self._check(self.L.gx_set_state(self.h, _dp(a)))

    def set_time(self, time: float) -> None:
        check(self.L.gx_set_time(self.h, float(time)))

    def get_timestep(self, current_iter: int, n_iter: int, current_time: float, tprint: float):
        """get_timestep (hydro_core.f90:623-697) -> (dt, dump_flag)."""
        dt = C.c_double(0.0)
        dump = C.c_int32(0)
        check(self.L.gx_get_timestep(self.h, current_iter, n_iter, current_time, tprint, C.byref(dt), C.byref(dump)))
        return dt.value, bool(dump.value)

    def tstep(self, dt_cfl: float) -> None:
        """tstep (hydro_solver.f90:134-229)."""
        self._check(self.L.gx_tstep(self.h, float(dt_cfl)))

    def run(self, n_steps: int, time: float, it: int, n_iter_ramp: int = 10):
        """n_steps iterations of main.f90's loop body, without output -> (time, iter, last_dt)."""
        t = C.c_double(time)
        i = C.c_int32(it)
        last = C.c_double(0.0)
        self._check(self.L.gx_run(self.h, n_steps, n_iter_ramp, C.byref(t), C.byref(i), C.byref(last)))
        return t.value, i.value, last.value

    def get_state(self, u: bool = True, primit: bool = False, temp: bool = False):
        """State on the host in reference layout (before write_output, main.f90:85,112)."""
        p = self.p
        ua = self.empty_state() if u else None
        pa = self.empty_state() if primit else None
        ta = np.zeros((p.nx + 4, p.ny + 4, p.nz + 4), dtype=np.float64, order="F") if temp else None
        check(self.L.gx_get_state(self.h, _dp(ua), _dp(pa), _dp(ta)))
        out = [a for a in (ua, pa, ta) if a is not None]
        return out[0] if len(out) == 1 else tuple(out)

    def get_state_into(self, u: np.ndarray) -> None:
        check(self.L.gx_get_state(self.h, _dp(u), None, None))

    def get_up(self) -> np.ndarray:
        a = self.empty_state()
        check(self.L.gx_get_up(self.h, _dp(a)))
        return a

    # -- user_mod plugin surface --
    def set_gravity_points(self, gm: Sequence[float], pos: Sequence[Sequence[float]]) -> None:
        gm_a = np.ascontiguousarray(gm, dtype=np.float64)
        pos_a = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1)
        check(self.L.gx_set_gravity_points(self.h, len(gm_a), _dp(gm_a), _dp(pos_a)))

    def set_wind_spheres(self, spheres: Sequence[WindSphere]) -> None:
        arr = (WindSphere * len(spheres))(*spheres)
        check(self.L.gx_set_wind_spheres(self.h, len(spheres), arr))

    def register_host_bc(self, fn: Optional[Callable[[np.ndarray, int], None]]) -> None:
        """Slow path mirroring impose_user_bc(u, order): `fn(u, order)` edits u in place."""
        if fn is None:
            self._host_bc_ref = None
            check(self.L.gx_register_host_bc(self.h, _lib.HOST_BC_FN(0), None))
            return
        shape = self.shape

        def tramp(ptr, order, _user):
            try:
                n = int(np.prod(shape))
                arr = np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape, order="F")
                fn(arr, int(order))
            except BaseException as e:       # re-raised by _check after the C call returns
                self._cb_error = self._cb_error or e

        self._host_bc_ref = _lib.HOST_BC_FN(tramp)
        check(self.L.gx_register_host_bc(self.h, self._host_bc_ref, None))

    def register_bc_hook(self, fn: Optional[Callable[[int, float], None]]) -> None:
        """`fn(order, time)` runs at the top of every impose_user_bc application; it may re-position the
        device functors (set_wind_spheres / set_gravity_points), like exoplanet.f90:137-144 moves the planet."""
        if fn is None:
            self._bc_hook_ref = None
            check(self.L.gx_register_bc_hook(self.h, _lib.BC_HOOK_FN(0), None))
            return
        def tramp(order, time, _user):
            try:
                fn(int(order), float(time))
            except BaseException as e:
                self._cb_error = self._cb_error or e

        self._bc_hook_ref = _lib.BC_HOOK_FN(tramp)
        check(self.L.gx_register_bc_hook(self.h, self._bc_hook_ref, None))

    def register_host_source(self, fn: Optional[Callable[[np.ndarray, np.ndarray], None]]) -> None:
        """Slow path mirroring get_user_source_terms (src/sources.f90:205): once per stage `fn(primit, s)` gets the block's
        primitives and a zero-filled source array, both (neq, nx+4, ny+4, nz+4) in reference layout, and adds to `s`."""
        if fn is None:
            self._host_src_ref = None
            check(self.L.gx_register_host_source(self.h, _lib.HOST_SOURCE_FN(0), None))
            return
        shape = self.shape

        def tramp(pw, ps, _user):
            try:
                n = int(np.prod(shape))
                w = np.ctypeslib.as_array(pw, shape=(n,)).reshape(shape, order="F")
                sarr = np.ctypeslib.as_array(ps, shape=(n,)).reshape(shape, order="F")
                fn(w, sarr)
            except BaseException as e:
                self._cb_error = self._cb_error or e

        self._host_src_ref = _lib.HOST_SOURCE_FN(tramp)
        check(self.L.gx_register_host_source(self.h, self._host_src_ref, None))

    def _check(self, rc: int) -> None:
        """check() for the calls that may run user callbacks: an exception raised inside one is re-raised here."""
        err, self._cb_error = self._cb_error, None
        if err is not None:
            raise err
        check(rc)

    # -- multi-GPU --
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.load().gx_comm_unique_id(buf, 128))
        return buf.raw

    def comm_attach(self, unique_id: bytes, rank: int, nranks: int) -> None:
        buf = C.create_string_buffer(unique_id, 128)
        check(self.L.gx_comm_attach(self.h, buf, 128, rank, nranks))

    # -- diagnostics --
    def tc_info(self):
        """(dt_cond [s], substeps) of the last step's thermal conduction — the reference's thermal_conduction.log line."""
        dt, n = C.c_double(0.0), C.c_int32(0)
        self._check(self.L.gx_tc_info(self.h, C.byref(dt), C.byref(n)))
        return dt.value, n.value

    @property
    def launch_count(self) -> int:
        return int(self.L.gx_launch_count(self.h))

    @property
    def last_elapsed_ms(self) -> float:
        return float(self.L.gx_last_elapsed_ms(self.h))

    def set_profiling(self, on: bool) -> None:
        check(self.L.gx_set_profiling(self.h, int(on)))

    def kernel_times(self) -> dict:
        out = {}
        for i, name in enumerate(KERNEL_CLASSES):
            ms = C.c_double(0)
            n = C.c_int64(0)
            check(self.L.gx_kernel_time_ms(self.h, i, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def interior(self, a: np.ndarray) -> np.ndarray:
        """Physical cells (1:nx, 1:ny, 1:nz) of a reference-layout array."""
        return a[..., 2:-2, 2:-2, 2:-2]


class Simulation:
    """The time loop of src/main.f90:94-125 for one block (single GPU) or one rank of many."""

    def __init__(self, block: Block, n_iter_ramp: int = 10):
        self.b = block
        self.p = block.p
        self.n_iter_ramp = n_iter_ramp           # main.f90:97 passes 10
        self.time = 0.0
        self.tprint = self.p.dtprint              # init.f90:131
        self.itprint = 0
        self.iteration = 1                        # currentIteration, init.f90:125
        self.on_output: Optional[Callable[["Simulation"], None]] = None

    def initflow(self, u: np.ndarray) -> None:
        self.b.set_state(u)

    def warm_start(self, outputpath: str, itprint0: int) -> None:
        """`iwarm = .true.`: restart from this block's BIN dump number `itprint0` (src/init.f90:134-142 sets
        itprint = itprint0, time = itprint*dtprint, tprint = time + dtprint; :436-471 reads u with ghosts and moves
        itprint on).  currentIteration restarts at 1, so the 10-step CFL ramp runs again, as in the reference."""
        from .bin_io import bin_name, read_bin
        u, hdr = read_bin(bin_name(outputpath, getattr(self.b, "rank", 0), itprint0))
        if tuple(u.shape) != tuple(self.p.block_shape()):
            raise ValueError(f"dump holds {u.shape}, this block is {self.p.block_shape()}")
        self.itprint = itprint0
        self.time = float(itprint0) * self.p.dtprint
        self.tprint = self.time + self.p.dtprint
        self.iteration = 1
        self.b.set_state(u)
        self.itprint += 1

    def step(self):
        dt, dump = self.b.get_timestep(self.iteration, self.n_iter_ramp, self.time, self.tprint)
        self.b.set_time(self.time)
        self.b.tstep(dt)
        self.time += dt
        if dump:
            if self.on_output:
                self.on_output(self)
            self.tprint += self.p.dtprint
            self.itprint += 1
        self.iteration += 1
        return dt

    def run(self, tmax: Optional[float] = None, max_steps: Optional[int] = None):
        tmax = self.p.tmax if tmax is None else tmax
        n = 0
        while self.time <= tmax and (max_steps is None or n < max_steps):
            self.step()
            n += 1
        return n


__all__ = ["Block", "Simulation", "GxError", "WindSphere"]
