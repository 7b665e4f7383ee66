"""GPU parity of the thermal-conduction operator (src/thermal_cond.f90): whole steps with th_cond != 0
through the C ABI against the oracle on the same inputs, <= 1e-12 relative per conserved variable (pow() on the device and
in libm differ by an ulp, so not bitwise), the conduction time scale and the number of substeps compared every step."""
import numpy as np
import pytest

from guacho_b200.config import (Params, SOLVER_HLLC, SOLVER_HLLD, BC_OUTFLOW, TC_ISOTROPIC, TC_ANISOTROPIC)
from tests.oracle_lib import U
from tests.util import global_ic, oracle_from_ic, rel_err_per_var, interior, tc_scalings

pytestmark = pytest.mark.gpu
TOL = 1e-12
OUTFLOW = dict(bc_left=BC_OUTFLOW, bc_right=BC_OUTFLOW, bc_bottom=BC_OUTFLOW, bc_top=BC_OUTFLOW, bc_out=BC_OUTFLOW, bc_in=BC_OUTFLOW)


def run_tc(p: Params, nsteps=2, first_iter=11, problem="random"):
    """first_iter = 11: past the 10-step CFL ramp (hydro_core.f90:677-682), so that the hydro step is long against the
    conduction time scale and the super-time-stepping schedule has several substeps."""
    from guacho_b200.solver import Block
    g = global_ic(p, problem)
    o = oracle_from_ic(p, g)
    info = []
    with Block(p) as b:
        b.set_state(g)
        t, it = 0.0, first_iter
        for _ in range(nsteps):
            dt_o, _ = o.get_timestep(it, 10, t, 1e300)
            dt_g, _ = b.get_timestep(it, 10, t, 1e300)
            assert abs(dt_g - dt_o) <= 1e-13 * abs(dt_o), (dt_g, dt_o)
            assert o.tstep(dt_o) == 0
            b.tstep(dt_o)
            (dc_o, n_o), (dc_g, n_g) = o.tc_info(), b.tc_info()
            assert n_o == n_g and abs(dc_g - dc_o) <= 1e-13 * dc_o, ((dc_o, n_o), (dc_g, n_g))
            info.append(n_g)
            t += dt_o
            it += 1
        ug = b.get_state()
        kt = b.launch_count
    assert kt > 0
    return interior(ug), interior(o.get_block(0, U)), info


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("mode,sat", [(TC_ISOTROPIC, False), (TC_ISOTROPIC, True), (TC_ANISOTROPIC, False), (TC_ANISOTROPIC, True)])
def test_thermal_conduction_mhd_outflow(mode, sat, strict):
    p = Params(nxtot=32, nytot=24, nztot=20, zmax=1.0, th_cond=mode, tc_saturation=sat, strict_fp=strict, **tc_scalings(), **OUTFLOW)
    ug, uo, nsub = run_tc(p)
    assert min(nsub) > 1, nsub                       # super-time-stepping ran
    assert rel_err_per_var(ug, uo).max() <= TOL, rel_err_per_var(ug, uo)


@pytest.mark.parametrize("sat", [False, True])
def test_marching_kernel_and_the_two_kernel_substep_agree(sat, monkeypatch):
    """One block, isotropic: k_tc_march (one launch per substep, every face flux once) against k_tc_update + k_tc_prim (what blocks
    with neighbours and the anisotropic operator use) — the strict flavours evaluate the same expressions on the same values."""
    from guacho_b200.solver import Block
    out = {}
    for strict in (True, False):
        p = Params(nxtot=40, nytot=20, nztot=24, zmax=1.0, th_cond=TC_ISOTROPIC, tc_saturation=sat, strict_fp=strict, **tc_scalings(), **OUTFLOW)
        g = global_ic(p, "random")
        for march in (True, False):
            if march:
                monkeypatch.delenv("GX_NO_TC_MARCH", raising=False)
            else:
                monkeypatch.setenv("GX_NO_TC_MARCH", "1")
            with Block(p) as b:
                b.set_state(g)
                t, it = 0.0, 11
                for _ in range(2):
                    dt, _ = b.get_timestep(it, 10, t, 1e300)
                    b.tstep(dt); t += dt; it += 1
                    assert b.tc_info()[1] > 1
                out[(strict, march)] = interior(b.get_state())
    assert np.array_equal(out[(True, True)], out[(True, False)])
    assert rel_err_per_var(out[(False, True)], out[(False, False)]).max() <= TOL


def test_thermal_conduction_single_substep_during_the_cfl_ramp():
    """dt_cond >= dt_hydro: SuperStep = .false., one substep of the hydro step (thermal_cond.f90:714-720)."""
    p = Params(nxtot=24, nytot=20, nztot=16, zmax=1.0, th_cond=TC_ISOTROPIC, **tc_scalings(), **OUTFLOW)
    ug, uo, nsub = run_tc(p, nsteps=3, first_iter=1)
    assert nsub == [1, 1, 1]
    assert rel_err_per_var(ug, uo).max() <= TOL


def test_thermal_conduction_periodic_box_keeps_the_zero_gradient_energy_ghosts():
    """Periodic boundaries: thermal_bounds still overwrites the energy ghost layer of every domain face with a zero-gradient
    copy (thermal_cond.f90:589-614); isotropic fluxes read face ghosts only, so the periodic box compares too."""
    p = Params(nxtot=32, nytot=24, nztot=20, zmax=1.0, th_cond=TC_ISOTROPIC, tc_saturation=True, **tc_scalings(rhosc=1e-16))
    ug, uo, nsub = run_tc(p)
    assert min(nsub) > 3
    assert rel_err_per_var(ug, uo).max() <= TOL


def test_thermal_conduction_hydro():
    p = Params(nxtot=24, nytot=20, nztot=16, zmax=1.0, mhd=False, riemann_solver=SOLVER_HLLC, enable_flux_cd=False,
               th_cond=TC_ISOTROPIC, tc_saturation=True, **tc_scalings(), **OUTFLOW)
    ug, uo, nsub = run_tc(p)
    assert min(nsub) > 1
    assert rel_err_per_var(ug, uo).max() <= TOL


def test_thermal_conduction_on_the_pass_per_routine_path():
    """8-wave source => the unfused kernels (primitives in HBM): thermal conduction runs after finish_u and the primitives
    and CFL candidates are refreshed after it."""
    p = Params(nxtot=24, nytot=20, nztot=16, zmax=1.0, enable_flux_cd=False, eight_wave=True,
               th_cond=TC_ANISOTROPIC, **tc_scalings(), **OUTFLOW)
    ug, uo, nsub = run_tc(p, nsteps=3)
    assert min(nsub) > 1
    assert rel_err_per_var(ug, uo).max() <= TOL


def test_thermal_conduction_needs_its_scalings():
    from guacho_b200.lib import GxError
    from guacho_b200.solver import Block
    p = Params(nxtot=16, nytot=16, nztot=16, zmax=1.0, th_cond=TC_ISOTROPIC, **tc_scalings(mu=0.0))
    with pytest.raises(GxError):
        Block(p)
    p = Params(nxtot=16, nytot=16, nztot=16, zmax=1.0, mhd=False, riemann_solver=SOLVER_HLLC, enable_flux_cd=False,
               th_cond=TC_ANISOTROPIC, **tc_scalings())
    with pytest.raises(GxError):
        Block(p)
