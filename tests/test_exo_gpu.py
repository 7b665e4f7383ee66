"""BASELINE.json configs[3]: the EXO exoplanet-wind set-up (EXO/parameters.f90, EXO/user_mod.f90,
EXO/exoplanet.f90) on a reduced grid — HLLD + full MHD + flux-CD, 2 passive scalars, EOS_H_RATE,
outflow walls, user boundary (two wind spheres, the planet orbiting), point-mass gravity of star and
planet as the user source, eta = 0.01, cfl 0.4.  The hydro/MHD step only (the cell-local COOL_H operator
that follows it in tstep is a SURVEY 8(f) 'next' row and is skipped on both sides).

The CUDA path gets the user plugins as device functors through the C ABI (gx_set_wind_spheres,
gx_set_gravity_points), re-positioned from the gx_register_bc_hook host hook at every boundary call, which is
where exoplanet.f90:137-144 moves the planet."""
import ctypes as C

import numpy as np
import pytest

from guacho_b200.config import Params, EOS_H_RATE, BC_OUTFLOW
from tests.oracle_lib import Oracle, U
from tests.util import rel_err_per_var, interior

pytestmark = pytest.mark.gpu

AU, AMH, RG, DAY = 1.496e13, 1.66e-24, 8.3145e7, 86400.0


def exo_params(nx, ny, nz, **kw):
    o = BC_OUTFLOW
    return Params(nxtot=nx, nytot=ny, nztot=nz, xmax=1.0, ymax=0.25, zmax=1.0, mhd=True, npas=2, eq_of_state=EOS_H_RATE,
                  enable_flux_cd=True, user_source_terms=True, bc_left=o, bc_right=o, bc_bottom=o, bc_top=o, bc_out=o, bc_in=o,
                  bc_user=True, cv=1.5, Tempsc=1.0e4 * (2.5 / 1.5), cfl=0.4, eta=0.01, **kw)


def scalings(p):
    gamma = p.gamma
    rsc = 0.3 * AU / 1.0
    rhosc = AMH * 1.0
    vsc2 = gamma * RG * 1.0e4 / 1.0
    tsc = rsc / np.sqrt(vsc2)
    bsc = np.sqrt(4.0 * np.arccos(-1.0) * rhosc * vsc2)
    return rsc, rhosc, p.Tempsc, vsc2, tsc, bsc


def exo_oracle(p):
    o = Oracle(p, threads=4)
    o.L.orc_init_exo(o.h, *[C.c_double(v) for v in scalings(p)])
    o.L.orc_exo_initial_conditions(o.h)
    return o


def exo_state(o):
    v = np.zeros(19)
    o.L.orc_exo_params(o.h, v.ctypes.data_as(C.POINTER(C.c_double)))
    keys = "RSW TSW VSW dsw RsS bsw bpw RPW TPW VPW dpw torb rorb omegap MassS MassP xp yp zp".split()
    return dict(zip(keys, v))


def set_functors(b, p, e, time):
    """What EXO/exoplanet.f90:137-144 and EXO/user_mod.f90:174-187 compute on the host each call."""
    from guacho_b200.lib import WindSphere
    rsc, rhosc, Tempsc, vsc2, tsc, bsc = scalings(p)
    pi = np.arccos(-1.0)
    phi = -25.0 * pi / 180.0
    xp = e["rorb"] * np.cos(e["omegap"] * time + phi)
    zp = e["rorb"] * np.sin(e["omegap"] * time + phi)
    vx = -e["omegap"] * e["rorb"] * np.sin(e["omegap"] * time + phi)
    vz = e["omegap"] * e["rorb"] * np.cos(e["omegap"] * time + phi)
    star = WindSphere(xc=0, yc=0, zc=0, radius=e["RSW"], vwind=e["VSW"], dens=e["dsw"], tfac=1.0, temp=e["TSW"], vbx=0, vby=0, vbz=0,
                      bdip=e["bsw"], pas=(C.c_double * 4)(0.0001, 1.0, 0, 0))
    planet = WindSphere(xc=xp, yc=0, zc=zp, radius=e["RPW"], vwind=e["VPW"], dens=e["dpw"], tfac=1.8, temp=e["TPW"], vbx=vx, vby=0.0, vbz=vz,
                        bdip=e["bpw"], pas=(C.c_double * 4)(0.2, -1.0, 0, 0))
    b.set_wind_spheres([star, planet])
    G = 6.67259e-8
    b.set_gravity_points([0.3 * G * e["MassS"] / rsc / vsc2, G * e["MassP"] / rsc / vsc2], [[0, 0, 0], [xp, 0, zp]])


@pytest.mark.parametrize("cool", [False, True])
@pytest.mark.parametrize("strict", [True, False])
def test_exo_reduced_grid_three_steps(strict, cool):
    """cool=True is EXO exactly as shipped: the COOL_H operator (src/cooling_h.f90) runs on the
    device after viscous_copy.  Its exp/log/pow are CUDA's FP64 routines vs glibc's in the oracle (<= 2 ulp
    apart), so the same 1e-12 gate applies."""
    from guacho_b200.solver import Block
    from guacho_b200.config import COOL_H
    p = exo_params(48, 12, 48, strict_fp=strict)
    if cool:
        p = p.replace(cooling=COOL_H, tsc=scalings(p)[4])
    o = exo_oracle(p)
    u0 = o.get_block(0, U)
    e = exo_state(o)
    o.start()
    with Block(p) as b:
        # the reference moves the planet inside impose_user_bc (exoplanet.f90:137-144) and the gravity source
        # then sees that position: same hook point here
        b.register_bc_hook(lambda order, time: set_functors(b, p, e, time))
        set_functors(b, p, e, 0.0)              # init_exo / initial_conditions (EXO/user_mod.f90:43-121)
        b.set_time(0.0)
        b.set_state(u0)
        t, it = 0.0, 1
        for _ in range(3):
            dt_o, _ = o.get_timestep(it, 10, t, 1e300)
            dt_g, _ = b.get_timestep(it, 10, t, 1e300)
            assert abs(dt_g - dt_o) <= 1e-12 * dt_o, (dt_g, dt_o)
            o.time = t
            assert o.tstep(dt_o) == 0
            b.set_time(t)
            b.tstep(dt_o)
            t += dt_o
            it += 1
        ug = interior(b.get_state())
    uo = interior(o.get_block(0, U))
    err = rel_err_per_var(ug, uo)
    assert err.max() <= 1e-12, err


def _run_exo_plugin_vs_oracle(nx, ny, nz, nsteps, strict):
    """The product-side EXO plugin (guacho_b200/exo.py: its own initial conditions and functor placement) against the oracle's
    restatement of EXO/user_mod.f90 + EXO/exoplanet.f90, COOL_H included (EXO exactly as shipped)."""
    from guacho_b200.solver import Block
    from guacho_b200.exo import Exo, exo_params as plugin_params
    p = plugin_params(nx, ny, nz, strict_fp=strict)
    # ONE oracle block: with eta != 0 the reference depends on the block decomposition (viscous_copy reads stale ghosts),
    # so the one-GPU run is compared with the one-rank reference
    o = Oracle(p, threads=1)
    o.L.orc_init_exo(o.h, *[C.c_double(v) for v in scalings(p)])
    o.L.orc_exo_initial_conditions(o.h)
    o.start()
    ex = Exo(p)
    with Block(p) as b:
        ex.attach(b, 0.0)
        b.set_state(ex.initial_conditions())
        t, it = 0.0, 1
        for _ in range(nsteps):
            dt_o, _ = o.get_timestep(it, 10, t, 1e300)
            dt_g, _ = b.get_timestep(it, 10, t, 1e300)
            assert abs(dt_g - dt_o) <= 1e-12 * dt_o, (dt_g, dt_o)
            o.time = t
            assert o.tstep(dt_o) == 0
            b.set_time(t)
            b.tstep(dt_o)
            t += dt_o
            it += 1
        ug = interior(b.get_state())
    uo = o.gather(U)
    return rel_err_per_var(ug, uo)


@pytest.mark.parametrize("strict", [True, False])
def test_exo_plugin_reduced_grid(strict):
    err = _run_exo_plugin_vs_oracle(96, 24, 96, 3, strict)
    assert err.max() <= 1e-12, err


def test_exo_as_shipped_400x100x400_one_step():
    """BASELINE configs[3] at its shipped size (EXO/parameters.f90:137-145: 400 x 100 x 400), one step of the production kernels
    against the oracle (one block, like the one-GPU run; ~30 s of CPU)."""
    err = _run_exo_plugin_vs_oracle(400, 100, 400, 1, False)
    assert err.max() <= 1e-12, err
