"""Pins the CPU oracle (oracle/guacho_oracle.cpp) — runs without a GPU.

The reference ships no golden vectors and cannot be built here, so the
oracle is pinned against (a) `published_*` fixtures computed from the PUBLISHED form of each
algorithm (tests/golden/make_golden.py: Miyoshi & Kusano 2005 jump-condition HLLD, HLL with
Davis speeds, Toro HLLC, the exact Sod solution, the analytic first Orszag-Tang time step)
and (b) `oracle_*` fixtures that freeze its own end-to-end output bitwise.
"""
import os

import numpy as np
import pytest

from guacho_b200.config import (Params, ot_shipped, SOLVER_HLL, SOLVER_HLLC, SOLVER_HLLE, SOLVER_HLLD,
                                LIMITER_MINMOD, BC_OUTFLOW)
from tests.oracle_lib import Oracle, U, PRIMIT
from tests.util import global_ic, oracle_from_ic

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def _oracle_fluxes(p: Params, WL, WR):
    o = Oracle(p)
    n = WL.shape[1]
    out = np.zeros((p.neq, n))
    for m in range(n):
        f, err = o.riemann(WL[:p.neq, m], WR[:p.neq, m])
        assert err == 0
        out[:, m] = f
    return out


@pytest.mark.parametrize("key,solver,mhd", [("hlld", SOLVER_HLLD, True), ("hlle", SOLVER_HLLE, True),
                                            ("hll", SOLVER_HLL, False), ("hllc", SOLVER_HLLC, False)])
def test_riemann_flux_matches_published_algorithm(key, solver, mhd):
    """prim2fhll* (src/hll.f90:47-82, hllc.f90:44-140, hlle.f90:48-83, hlld.f90:48-319) vs the
    published flux formulae in conserved/jump-condition form, 1024 random interfaces covering
    every wave region.  Tolerance 1e-11 of the flux scale: the two forms are algebraically
    equal but evaluate in a different order."""
    g = _load("published_riemann.npz")
    p = Params(nxtot=8, nytot=8, nztot=8, mhd=mhd, riemann_solver=solver, enable_flux_cd=False)
    assert abs(p.gamma - float(g["gamma"])) < 1e-15 and abs(p.cv - float(g["cv"])) < 1e-15
    got = _oracle_fluxes(p, g["WL"], g["WR"])
    ref = g[key]
    scale = np.abs(ref).max(axis=0) + 1.0
    err = (np.abs(got - ref) / scale).max()
    assert err <= 1e-11, err
    if key == "hlld":      # every region of the five-wave fan is exercised
        assert (np.bincount(g["hlld_region"], minlength=6) >= 30).all()


def test_first_timestep_of_shipped_orszag_tang():
    """get_timestep (src/hydro_core.f90:623-697) on the shipped OT set-up (512x512x2, cfl 0.2,
    10-step ramp) vs the value derived analytically from the initial condition."""
    g = _load("published_ot_dt.npz")
    p = ot_shipped()
    o = oracle_from_ic(p, global_ic(p, "ot"), threads=4)
    dt, dump = o.get_timestep(1, 10, 0.0, p.dtprint)
    assert not dump
    assert abs(dt - float(g["dt_first"])) <= 1e-13 * dt
    assert abs(dt - 1.7222826491e-7) <= 1e-17          # SURVEY 8(c) KAT 6
    # after the ramp the step is cfl * dtp
    dt11, _ = o.get_timestep(11, 10, 0.0, p.dtprint)
    assert abs(dt11 - 0.2 * float(g["dtp"])) <= 1e-13 * dt11


def test_sod_tube_converges_to_exact_solution():
    """Full driver (tstep: both stages, boundaries, CFL) on the Sod problem with HLLC + minmod
    vs the exact Riemann solution; L1(rho) at N=400 must be at the level second-order
    Godunov codes reach (< 4e-3), and halve or better from N=200."""
    g = _load("published_sod.npz")

    def run(n):
        p = Params(nxtot=n, nytot=2, nztot=2, xmax=1.0, ymax=2.0 / n, zmax=2.0 / n, mhd=False, cv=2.5,
                   riemann_solver=SOLVER_HLLC, enable_flux_cd=False, slope_limiter=LIMITER_MINMOD, cfl=0.4,
                   bc_left=BC_OUTFLOW, bc_right=BC_OUTFLOW)
        x = (np.arange(-1, n + 3) - 0.5) / n
        rho = np.where(x < 0.5, 1.0, 0.125)
        pr = np.where(x < 0.5, 1.0, 0.1)
        u0 = np.zeros(p.block_shape(), order="F")
        u0[0] = rho[:, None, None]
        u0[4] = (p.cv * pr)[:, None, None]
        o = oracle_from_ic(p, u0, threads=2)
        while o.time < 0.2:
            dt, _ = o.get_timestep(o.iter, 10, o.time, 0.2)
            assert o.tstep(dt) == 0
            o.time += dt
            o.iter += 1
        return o.get_block(0, PRIMIT)[0, 2:-2, 2, 2]
    from tests.golden.make_golden import sod_exact
    errs = {}
    for n in (200, 400):
        xs = (np.arange(n) + 0.5) / n
        exact = sod_exact(xs, 0.2)[0]
        errs[n] = np.abs(run(n) - exact).mean()
    assert np.allclose(sod_exact(g["x"], 0.2)[0], g["rho"], rtol=0, atol=1e-14)      # fixture == generator
    assert abs(float(g["pstar"]) - 0.30313) < 1e-5 and abs(float(g["ustar"]) - 0.92745) < 1e-5   # Toro table 4.2
    assert errs[400] < 4e-3, errs
    assert errs[400] < 0.62 * errs[200], errs


@pytest.mark.parametrize("name", ["oracle_ot_hlld_cd_24x20x4", "oracle_random_hlld_cd_16x12x10", "oracle_random_hllc_16x12x10"])
def test_oracle_end_to_end_frozen(name):
    """The oracle reproduces its committed end-to-end fixtures bitwise (guards silent edits)."""
    g = _load(name + ".npz")
    nx, ny, nz = (int(v) for v in g["params"])
    kw = dict(nxtot=nx, nytot=ny, nztot=nz, zmax=float(g["zmax"]))
    if "hllc" in name:
        kw.update(mhd=False, riemann_solver=SOLVER_HLLC, enable_flux_cd=False)
    p = Params(**kw)
    o = oracle_from_ic(p, g["u0"], threads=1)
    dts = o.advance(int(g["nsteps"]))
    assert dts == list(g["dts"])
    assert np.array_equal(o.get_block(0, U)[..., 2:-2, 2:-2, 2:-2], g["u"])
