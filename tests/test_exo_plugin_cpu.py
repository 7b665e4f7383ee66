"""The host-side EXO plugin (guacho_b200/exo.py: EXO/parameters.f90 scalings, init_exo, initial_conditions) against the
oracle's restatement of the same Fortran (oracle/guacho_oracle.cpp: orc_init_exo / orc_exo_initial_conditions) — no GPU."""
import ctypes as C

import numpy as np

from guacho_b200.exo import Exo, Scalings, exo_params
from tests.oracle_lib import Oracle, U


def _oracle(p):
    s = Scalings.of(p)
    o = Oracle(p, threads=2)
    o.L.orc_init_exo(o.h, *[C.c_double(v) for v in (s.rsc, s.rhosc, s.Tempsc, s.vsc2, s.tsc, s.bsc)])
    o.L.orc_exo_initial_conditions(o.h)
    return o


def test_exo_parameters_and_initial_conditions_match_the_oracle():
    p = exo_params(200, 50, 200, cooling=False)       # fine enough for cell centres inside the planet's wind sphere
    o = _oracle(p)
    e = Exo(p)
    v = np.zeros(19)
    o.L.orc_exo_params(o.h, v.ctypes.data_as(C.POINTER(C.c_double)))
    keys = "RSW TSW VSW dsw RsS bsw bpw RPW TPW VPW dpw torb rorb omegap MassS MassP xp yp zp".split()
    d = dict(zip(keys, v))
    for k in ("RSW", "TSW", "VSW", "dsw", "bsw", "bpw", "RPW", "TPW", "VPW", "dpw", "torb", "rorb", "omegap", "MassS", "MassP"):
        assert abs(d[k] - getattr(e, k)) <= 1e-15 * abs(d[k]), k
    (xp, yp, zp), _ = e.planet(0.0)
    assert abs(xp - d["xp"]) <= 1e-15 and abs(zp - d["zp"]) <= 1e-15 and yp == d["yp"] == 0.0
    u, uo = e.initial_conditions(), o.get_block(0, U)
    for q in range(p.neq):
        assert np.abs(u[q] - uo[q]).max() <= 1e-14 * np.abs(uo[q]).max(), q
    # both spheres are on this grid
    assert (uo[9] < 0).any() and (uo[8] == 0.0001 * uo[0]).any()


def test_shipped_exo_parameters():
    p = exo_params()
    assert (p.nxtot, p.nytot, p.nztot, p.neq, p.npas) == (400, 100, 400, 10, 2)
    assert p.mhd and p.enable_flux_cd and p.user_source_terms and p.bc_user and p.eta == 0.01 and p.cfl == 0.4
    s = Scalings.of(p)
    assert abs(s.tsc - 0.3 * 1.496e13 / np.sqrt(p.gamma * 8.3145e7 * 1.0e4)) <= 1e-9 * s.tsc and p.tsc == s.tsc
