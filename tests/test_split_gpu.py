"""GPU parity of SOLVER_HLLE_SPLIT_ALL (src/hlle_split_all.f90): the device sweep with background states
through the C ABI against the oracle on the same inputs — whole steps, both build flavours, flux-CD on and off."""
import numpy as np
import pytest

from guacho_b200.config import Params, SOLVER_HLLE, SOLVER_HLLE_SPLIT_ALL, LIMITER_MINMOD, LIMITER_VAN_LEER, LIMITER_SUPERBEE
from tests.oracle_lib import Oracle, U, PRIMIT
from tests.util import global_ic, rel_err_per_var, interior

pytestmark = pytest.mark.gpu
TOL = 1e-12


def background(p: Params, uniform: bool) -> np.ndarray:
    """primit0 with ghosts: uniform, or a smooth static field (the reference never fills it; any array is legal input)."""
    g0 = np.zeros((p.neq, p.nxtot + 4, p.nytot + 4, p.nztot + 4), order="F")
    x = (np.arange(p.nxtot + 4) - 1.5) * p.dx
    y = (np.arange(p.nytot + 4) - 1.5) * p.dy
    z = (np.arange(p.nztot + 4) - 1.5) * p.dz
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    s = 0.0 if uniform else 0.1
    g0[0] = 0.7 + s * np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y)
    g0[4] = 0.4 + s * np.cos(2 * np.pi * Z)
    g0[5] = 0.3 + s * np.sin(2 * np.pi * Y)
    g0[6] = -0.2 + s * np.sin(2 * np.pi * Z)
    g0[7] = 0.5 + s * np.sin(2 * np.pi * X)
    return g0


def run_split(p: Params, uniform=False, nsteps=3):
    from guacho_b200.solver import Block
    g = global_ic(p, "random")
    g0 = background(p, uniform)
    fl = g.copy()                                      # any fluctuation field will do: keep rho + rho0, p + p0 positive
    fl[0] -= 0.5
    o = Oracle(p, threads=4)
    o.scatter_u(fl); o.scatter_primit0(g0); o.start()
    with Block(p) as b:
        b.set_background(g0)
        b.set_state(fl)
        t, it = 0.0, 1
        for _ in range(nsteps):
            dt_o, _ = o.get_timestep(it, 10, t, 1e300)
            dt_g, _ = b.get_timestep(it, 10, t, 1e300)
            assert abs(dt_g - dt_o) <= 1e-13 * abs(dt_o), (dt_g, dt_o)
            assert o.tstep(dt_o) == 0
            b.tstep(dt_o)
            t += dt_o; it += 1
        ug, wg = b.get_state(u=True, primit=True)
        assert b.launch_count > 0
    return interior(ug), interior(o.get_block(0, U)), interior(wg), interior(o.get_block(0, PRIMIT))


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("cd", [True, False])
@pytest.mark.parametrize("lim", [LIMITER_MINMOD, LIMITER_VAN_LEER])
def test_split_all_hlle_matches_the_oracle(lim, cd, strict):
    p = Params(nxtot=32, nytot=24, nztot=20, zmax=1.0, riemann_solver=SOLVER_HLLE_SPLIT_ALL, slope_limiter=lim, enable_flux_cd=cd, strict_fp=strict)
    ug, uo, wg, wo = run_split(p)
    assert rel_err_per_var(ug, uo).max() <= (1e-14 if strict else TOL), rel_err_per_var(ug, uo)
    assert rel_err_per_var(wg, wo).max() <= TOL


def test_split_all_with_uniform_background_tracks_plain_hlle_on_the_gpu():
    """The identity that pins the oracle (tests/test_oracle_split.py), on the device: fluctuation + uniform background
    evolves like the total state under plain HLLE."""
    from guacho_b200.solver import Block
    p = Params(nxtot=32, nytot=24, nztot=20, zmax=1.0, riemann_solver=SOLVER_HLLE_SPLIT_ALL, slope_limiter=LIMITER_SUPERBEE)
    g = global_ic(p, "random")
    g0 = background(p, True)
    bg = g0[:, 2, 2, 2]
    fl = g.copy()
    fl[0] -= bg[0]; fl[4] -= p.cv * bg[4] + 0.5 * (bg[5] ** 2 + bg[6] ** 2 + bg[7] ** 2); fl[5] -= bg[5]; fl[6] -= bg[6]; fl[7] -= bg[7]
    with Block(p) as bs, Block(p.replace(riemann_solver=SOLVER_HLLE)) as bp:
        bs.set_background(g0); bs.set_state(fl); bp.set_state(g)
        t, it = 0.0, 1
        for _ in range(3):
            dt, _ = bp.get_timestep(it, 10, t, 1e300)
            dts, _ = bs.get_timestep(it, 10, t, 1e300)
            assert abs(dts - dt) <= 1e-12 * dt
            bs.tstep(dt); bp.tstep(dt); t += dt; it += 1
        us, up = interior(bs.get_state()), interior(bp.get_state())
    us[0] += bg[0]; us[4] += p.cv * bg[4] + 0.5 * (bg[5] ** 2 + bg[6] ** 2 + bg[7] ** 2); us[5] += bg[5]; us[6] += bg[6]; us[7] += bg[7]
    assert rel_err_per_var(us, up).max() <= 1e-11, rel_err_per_var(us, up)


def test_split_all_needs_its_background():
    from guacho_b200.lib import GxError
    from guacho_b200.solver import Block
    p = Params(nxtot=16, nytot=16, nztot=16, zmax=1.0, riemann_solver=SOLVER_HLLE_SPLIT_ALL)
    with Block(p) as b:
        with pytest.raises(GxError):
            b.set_state(b.empty_state())
    with Block(p.replace(riemann_solver=SOLVER_HLLE)) as b:
        with pytest.raises(GxError):
            b.set_background(b.empty_state())
