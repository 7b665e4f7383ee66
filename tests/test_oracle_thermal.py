"""Known answers for the oracle's restatement of src/thermal_cond.f90: the reference ships no
vectors for it, so it is pinned by what is published about the method and by what follows from the code.

* super-time-stepping (Alexiades, Amiez & Gremaud 1996): the N substeps tau_j = dt / ((nu-1) cos(pi (2j-1)/(2N)) + 1 + nu)
  add up to the superstep  dt N/(2 sqrt nu) [(1+sqrt nu)^2N - (1-sqrt nu)^2N] / [(1+sqrt nu)^2N + (1-sqrt nu)^2N];
* Spitzer conduction of a small cosine perturbation decays like exp(-D lambda t) with D = K(T) Tempsc / (cv rho Psc rsc^2);
* uniform temperature is a fixed point, zero-gradient walls conserve the total energy, the block decomposition does not
  change the result.
"""
import numpy as np
import pytest

from guacho_b200.config import (Params, SOLVER_HLLC, SOLVER_HLLD, BC_OUTFLOW, BC_PERIODIC, TC_ISOTROPIC, TC_ANISOTROPIC)
from tests.oracle_lib import Oracle, U, TEMP, load
from tests.util import global_ic, oracle_from_ic, interior, tc_scalings

MU, RG, GAMMA, T0 = 0.6, 8.3145e7, 5.0 / 3.0, 1.0e6


def cgs(**kw):
    return tc_scalings(**kw)


OUTFLOW = dict(bc_left=BC_OUTFLOW, bc_right=BC_OUTFLOW, bc_bottom=BC_OUTFLOW, bc_top=BC_OUTFLOW, bc_out=BC_OUTFLOW, bc_in=BC_OUTFLOW)


def test_sts_substeps_add_up_to_the_superstep():
    L = load()
    for n in range(1, 80):
        total = sum(L.orc_tc_substep(j, n) for j in range(1, n + 1))
        assert abs(total - L.orc_tc_superstep(n)) <= 1e-11 * total, n
    # closed form, evaluated independently
    snu = np.sqrt(0.01)
    for n in (1, 2, 7, 33):
        a, b = (1 + snu) ** (2 * n), (1 - snu) ** (2 * n)
        assert abs(L.orc_tc_superstep(n) - n / (2 * snu) * (a - b) / (a + b)) <= 1e-13 * n


def test_st_steps_picks_the_smallest_sufficient_schedule():
    import ctypes as C
    L = load()
    for fs in (1.0001, 1.9, 3.0, 8.2, 40.0, 150.0):
        ns, fstep = C.c_int(0), C.c_double(0.0)
        L.orc_tc_st_steps(fs, C.byref(ns), C.byref(fstep))
        assert L.orc_tc_superstep(ns.value) > fs
        assert ns.value == 1 or L.orc_tc_superstep(ns.value - 1) <= fs
        assert abs(fstep.value * L.orc_tc_superstep(ns.value) - fs) <= 1e-14 * fs and 0 < fstep.value <= 1


def test_spitzer_coefficients():
    L = load()
    for T in (1e4, 3.3e5, 2e7):
        assert abs(L.orc_tc_ksp(0, T, 0, 0) - 6e-7 * T ** 2.5) <= 1e-15 * 6e-7 * T ** 2.5
        assert abs(L.orc_tc_ksp(1, T, 0, 0) - 9.2181e-7 * T ** 2.5) <= 1e-15 * 9.2181e-7 * T ** 2.5
        n, b2 = 3e8, 7.0
        assert abs(L.orc_tc_ksp(2, T, n, b2) - 0.30089e33 * n * n / (b2 * np.sqrt(T))) <= 1e-14 * 0.30089e33 * n * n / (b2 * np.sqrt(T))


def _static_state(p, temp_of_x):
    """rho = 1, v = 0 (B = 0), p/rho = temp_of_x(x)/Tempsc: u with ghosts for one block."""
    g = np.zeros((p.neq, p.nxtot + 4, p.nytot + 4, p.nztot + 4), order="F")
    x = (np.arange(p.nxtot + 4) - 2 + 0.5) * p.dx
    g[0] = 1.0
    g[4] = p.cv * (temp_of_x(x) / p.Tempsc)[:, None, None]
    return g


@pytest.mark.parametrize("m", [1, 4])
@pytest.mark.parametrize("nsub_expected", ["single", "super"])
def test_cosine_mode_decays_at_the_spitzer_rate(nsub_expected, m):
    n = 32
    p = Params(nxtot=n, nytot=4, nztot=4, ymax=4.0 / n, zmax=4.0 / n, mhd=False, riemann_solver=SOLVER_HLLC, enable_flux_cd=False,
               th_cond=TC_ISOTROPIC, **cgs(), **OUTFLOW)
    k = m * np.pi / 1.0                                       # cos(m pi x / L): zero gradient at both walls
    amp = 1e-5
    g = _static_state(p, lambda x: T0 * (1.0 + amp * np.cos(k * x)))
    o = oracle_from_ic(p, g)
    D = 6e-7 * T0 ** 2.5 * p.Tempsc / (p.cv * 1.0 * (p.rhosc * p.vsc2) * p.rsc ** 2)        # code length^2 per second
    lam = (2 - 2 * np.cos(k * p.dx)) / p.dx ** 2              # eigenvalue of the discrete Neumann Laplacian for this mode
    # choose the hydro step so that conduction needs one step / a super-time-stepping schedule
    o.thermal_conduction(1e-30)
    dt_cond, _ = o.tc_info()
    dt_hydro = 0.5 * dt_cond if nsub_expected == "single" else 12.0 * dt_cond
    o = oracle_from_ic(p, g)
    o.thermal_conduction(dt_hydro / p.tsc)
    _, nsteps = o.tc_info()
    assert (nsteps == 1) if nsub_expected == "single" else (nsteps > 2)
    T = interior(o.get_block(0, TEMP)[None])[0][:, 0, 0]
    x = (np.arange(n) + 0.5) * p.dx
    a1 = 2.0 * np.mean((T / T0 - 1.0) * np.cos(k * x))        # projection on the mode
    expect = amp * np.exp(-D * lam * dt_hydro)
    decay = amp - expect
    print(f"mode {m} {nsub_expected}: substeps {nsteps}, decayed by {(amp - a1) / amp:.4%}, analytic {decay / amp:.4%}")
    # explicit Euler / first-order super-time-stepping against the exponential: a few per cent of the decay
    assert abs(a1 - expect) <= (0.03 if nsub_expected == "single" else 0.10) * decay, (a1, expect, nsteps)
    assert a1 < amp - 0.9 * decay                              # it did decay


def test_uniform_temperature_is_a_fixed_point():
    p = Params(nxtot=12, nytot=10, nztot=8, zmax=1.0, th_cond=TC_ANISOTROPIC, tc_saturation=True, **cgs(), **OUTFLOW)
    g = global_ic(p, "random")
    # make p/rho uniform: keep rho, v, B and re-set the energy
    rho = g[0]
    ke = 0.5 * (g[1] ** 2 + g[2] ** 2 + g[3] ** 2) / rho
    g[4] = p.cv * (0.7 * rho) + ke + 0.5 * (g[5] ** 2 + g[6] ** 2 + g[7] ** 2)
    o = oracle_from_ic(p, g)
    T = interior(o.get_block(0, TEMP)[None])[0]
    u0 = o.get_block(0, U).copy()
    o.thermal_conduction(1e-3)
    d = interior(o.get_block(0, U)) - interior(u0)
    # round-off differences of T between cells are conducted; nothing of order one may happen
    assert np.abs(d[4]).max() <= 1e-9 * np.abs(interior(u0)[4]).max(), np.abs(d[4]).max()
    assert np.ptp(T) <= 1e-9 * T.mean()
    assert np.abs(d[[0, 1, 2, 3, 5, 6, 7]]).max() == 0.0


@pytest.mark.parametrize("mode,sat", [(TC_ISOTROPIC, False), (TC_ISOTROPIC, True), (TC_ANISOTROPIC, False), (TC_ANISOTROPIC, True)])
def test_zero_gradient_walls_conserve_energy_and_only_energy_changes(mode, sat):
    p = Params(nxtot=16, nytot=12, nztot=10, zmax=1.0, th_cond=mode, tc_saturation=sat, **cgs(), **OUTFLOW)
    g = global_ic(p, "random")
    o = oracle_from_ic(p, g)
    u0 = interior(o.get_block(0, U)).copy()
    o.thermal_conduction(3e-3)                                 # several substeps
    _, nsteps = o.tc_info()
    assert nsteps > 1
    u1 = interior(o.get_block(0, U))
    assert np.abs(u1[[0, 1, 2, 3, 5, 6, 7]] - u0[[0, 1, 2, 3, 5, 6, 7]]).max() == 0.0
    assert np.abs(u1[4] - u0[4]).max() > 1e-6 * np.abs(u0[4]).max()
    if mode == TC_ISOTROPIC:                                   # face fluxes telescope; the ghost cells copy their neighbours: no wall flux
        assert abs(u1[4].sum() - u0[4].sum()) <= 1e-12 * np.abs(u0[4]).sum()


@pytest.mark.parametrize("blocks", [(2, 1, 1), (1, 1, 2), (1, 2, 2)])
def test_isotropic_conduction_does_not_depend_on_the_block_decomposition(blocks):
    kw = dict(nxtot=16, nytot=12, nztot=12, zmax=1.0, th_cond=TC_ISOTROPIC, tc_saturation=True, **cgs(), **OUTFLOW)
    p1 = Params(**kw)
    g = global_ic(p1, "random")
    o1 = oracle_from_ic(p1, g)
    pb = Params(MPI_NBX=blocks[0], MPI_NBY=blocks[1], MPI_NBZ=blocks[2], **kw)
    ob = oracle_from_ic(pb, g)
    for o in (o1, ob):
        dt, _ = o.get_timestep(11, 10, 0.0, 1e300)
        assert o.tstep(dt) == 0
    assert o1.tc_info() == ob.tc_info() and o1.tc_info()[1] > 1
    assert np.array_equal(o1.gather(U), ob.gather(U))


def test_periodic_boundaries_get_zero_gradient_energy_ghosts():
    """thermal_bounds overwrites every face of the domain with a zero-gradient copy of u(5), whatever the boundary type
    (src/thermal_cond.f90:589-614): with periodic boundaries the energy ghost layer is NOT the periodic image."""
    p = Params(nxtot=12, nytot=10, nztot=8, zmax=1.0, th_cond=TC_ISOTROPIC, **cgs())
    assert p.bc_left == BC_PERIODIC
    g = global_ic(p, "random")
    o = oracle_from_ic(p, g)
    o.thermal_conduction(3e-3)
    u = o.get_block(0, U)
    assert np.array_equal(u[4, 1, 2:-2, 2:-2], u[4, 2, 2:-2, 2:-2]) and np.array_equal(u[4, -2, 2:-2, 2:-2], u[4, -3, 2:-2, 2:-2])
    assert not np.array_equal(u[4, 1, 2:-2, 2:-2], u[4, -3, 2:-2, 2:-2])
    assert np.array_equal(u[0, 1, 2:-2, 2:-2], u[0, -3, 2:-2, 2:-2])          # the other variables keep boundaryI's periodic image
