"""Known answers for the oracle's restatement of the split-all HLLE solver (src/hlle_split_all.f90, the split branches of
prim2u / prim2f / u2primSplitAll / get_timestep in src/hydro_core.f90).  The reference holds no vectors for it
(and never fills primit0 itself); what follows from the code:
* with a zero background the split expressions add exact zeros to the plain ones: HLLE bit for bit;
* a uniform, static background is an exact steady state, and the split flux differs from the flux of the total state by a
  constant: fluctuation + background evolves like the total state under the (pinned) plain HLLE, to round-off.
"""
import numpy as np
import pytest

from guacho_b200.config import Params, SOLVER_HLLE, SOLVER_HLLE_SPLIT_ALL, LIMITER_MINMOD, LIMITER_VAN_LEER
from tests.oracle_lib import Oracle, U, PRIMIT
from tests.util import global_ic, oracle_from_ic, rel_err_per_var

BG = np.array([0.7, 0.0, 0.0, 0.0, 0.4, 0.3, -0.2, 0.5])          # rho0, v0 = 0, p0, B0


def split_pair(p_split: Params, g_total: np.ndarray, bg, nsteps=3):
    """(oracle with the split solver on fluctuation = total - background, oracle with plain HLLE on the total state)."""
    p_plain = p_split.replace(riemann_solver=SOLVER_HLLE)
    o_plain = oracle_from_ic(p_plain, g_total)
    g0 = np.zeros_like(g_total)
    for q in range(8):
        g0[q] = bg[q]
    fl = g_total.copy()
    fl[0] -= bg[0]
    fl[4] -= p_split.cv * bg[4] + 0.5 * (bg[5] ** 2 + bg[6] ** 2 + bg[7] ** 2)     # Out_BIN_Module.f90:152-153 read backwards
    fl[5] -= bg[5]; fl[6] -= bg[6]; fl[7] -= bg[7]
    o_split = Oracle(p_split, threads=4)
    o_split.scatter_u(fl)
    o_split.scatter_primit0(g0)
    o_split.start()
    t, it = 0.0, 1
    for _ in range(nsteps):
        dt_s, _ = o_split.get_timestep(it, 10, t, 1e300)
        dt_p, _ = o_plain.get_timestep(it, 10, t, 1e300)
        assert abs(dt_s - dt_p) <= 1e-12 * dt_p, (dt_s, dt_p)
        assert o_split.tstep(dt_p) == 0 and o_plain.tstep(dt_p) == 0
        t += dt_p; it += 1
    return o_split, o_plain


@pytest.mark.parametrize("cd", [True, False])
def test_zero_background_is_plain_hlle_bit_for_bit(cd):
    p = Params(nxtot=16, nytot=12, nztot=10, zmax=1.0, riemann_solver=SOLVER_HLLE_SPLIT_ALL, enable_flux_cd=cd)
    g = global_ic(p, "random")
    o_split, o_plain = split_pair(p, g, np.zeros(8))
    assert np.array_equal(o_split.gather(U), o_plain.gather(U))
    assert np.array_equal(o_split.gather(PRIMIT), o_plain.gather(PRIMIT))


@pytest.mark.parametrize("lim", [LIMITER_MINMOD, LIMITER_VAN_LEER])
def test_uniform_background_plus_fluctuation_evolves_like_the_total_state(lim):
    p = Params(nxtot=16, nytot=12, nztot=10, zmax=1.0, riemann_solver=SOLVER_HLLE_SPLIT_ALL, slope_limiter=lim)
    g = global_ic(p, "random")                       # total state; rho ~ 1 +- 0.2, p ~ 1 +- 0.2: above the background
    o_split, o_plain = split_pair(p, g, BG)
    us, up = o_split.gather(U), o_plain.gather(U)
    tot = us.copy()
    tot[0] += BG[0]
    tot[4] += p.cv * BG[4] + 0.5 * (BG[5] ** 2 + BG[6] ** 2 + BG[7] ** 2)
    tot[5] += BG[5]; tot[6] += BG[6]; tot[7] += BG[7]
    assert rel_err_per_var(tot, up).max() <= 1e-12, rel_err_per_var(tot, up)
    assert us[0].mean() < 0.5 < up[0].mean()          # it really ran on the fluctuation


def test_per_interface_split_flux_with_zero_background_equals_hlle():
    d = np.load("tests/golden/published_riemann.npz")
    p = Params(nxtot=8, nytot=8, nztot=8, zmax=1.0, riemann_solver=SOLVER_HLLE_SPLIT_ALL)
    o_s, o_p = Oracle(p), Oracle(p.replace(riemann_solver=SOLVER_HLLE))
    z = np.zeros(8)
    WL, WR = d["WL"].T.copy(), d["WR"].T.copy()
    for n in range(0, len(WL), 7):
        f_p, err = o_p.riemann(WL[n], WR[n])
        assert err == 0 and np.array_equal(o_s.riemann_split_all(WL[n], WR[n], z, z), f_p)


def test_per_interface_split_flux_is_the_total_flux_minus_a_constant():
    d = np.load("tests/golden/published_riemann.npz")
    p = Params(nxtot=8, nytot=8, nztot=8, zmax=1.0, riemann_solver=SOLVER_HLLE_SPLIT_ALL)
    o_s, o_p = Oracle(p), Oracle(p.replace(riemann_solver=SOLVER_HLLE))
    bg = BG * 0.05
    # flux of the background alone (v0 = 0): only the momentum flux carries its total pressure
    const = np.zeros(8)
    const[1] = bg[4] + 0.5 * (bg[6] ** 2 + bg[7] ** 2 - bg[5] ** 2)
    const[2] = -bg[5] * bg[6]
    const[3] = -bg[5] * bg[7]
    WL, WR = d["WL"].T.copy(), d["WR"].T.copy()
    worst = 0.0
    for n in range(0, len(WL), 5):
        f_tot, err = o_p.riemann(WL[n], WR[n])
        fl_l, fl_r = WL[n] - bg, WR[n] - bg
        fl_l[1:4], fl_r[1:4] = WL[n][1:4], WR[n][1:4]            # velocities are not split (u2primSplitAll: v = m / (rho + rho0))
        f_s = o_s.riemann_split_all(fl_l, fl_r, bg, bg)
        worst = max(worst, np.abs(f_s + const - f_tot).max() / max(1.0, np.abs(f_tot).max()))
    assert worst <= 1e-13, worst
