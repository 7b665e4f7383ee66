#!/usr/bin/env python
"""Generates the committed fixtures in tests/golden/ (run from the repo root:
``python tests/golden/make_golden.py``).

The reference (Fortran + MPI) cannot be compiled or run in this image and ships no test
vectors, so these fixtures are NOT outputs of the reference.  They are
of two kinds, and the file names say which:

* ``published_*.npz`` — answers computed HERE, in numpy, from the PUBLISHED form of each
  algorithm, written independently of oracle/guacho_oracle.cpp and of the reference's
  algebra:
    - HLLD: Miyoshi & Kusano (2005), J. Comput. Phys. 208, 315, eqs. (38)-(41) for S_M and
      p_T*, (43)-(48) for U*, (51) for S*, (59)-(63) for U**, and the jump-condition flux
      form (64)-(66)  F* = F_K + S_K (U*_K - U_K),
      F** = F_K + S*_K U**_K - (S*_K - S_K) U*_K - S_K U_K.
      The reference (src/hlld.f90:48-319) instead evaluates the flux function directly on
      the star state, so agreement is an algebraic identity holding to round-off when
      B_x is continuous across the interface (the fixture uses B_xL == B_xR).
    - HLL / HLLE: Harten, Lax & van Leer (1983) two-wave flux with Davis (1988) speeds.
    - HLLC: Toro, Spruce & Speares (1994) with the Batten et al. (1997) contact speed,
      F*_K = F_K + S_K (U*_K - U_K).
    - Sod shock tube: exact Riemann solution (Toro 2009, ch. 4) sampled on the cell centres.
    - first CFL step of the shipped Orszag-Tang set-up, from the analytic initial
      condition (OT/orzag_tang.f90:14-70) and the fast-speed formula.
  The CPU tests check the oracle against these; tolerance 1e-11 relative (different but
  algebraically equivalent operation order).

* ``oracle_*.npz`` — small end-to-end states produced by the C++ oracle (after it passed
  the published_* and the known-answer tests).  They pin the oracle against silent edits
  (CPU test, bitwise) and give the GPU tests an answer that does not need the oracle
  library at run time.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

GAMMA = 5.0 / 3.0
CV = 1.0 / (GAMMA - 1.0)          # parameters.f90: cv = 1.5, gamma = (cv+1)/cv


# ----------------------------------------------------------------------------------------
# published algorithms, conserved-variable form.  W = (rho, u, v, w, p[, Bx, By, Bz])
def mhd_cons(W):
    r, u, v, w, p, bx, by, bz = W
    e = p / (GAMMA - 1.0) + 0.5 * r * (u * u + v * v + w * w) + 0.5 * (bx * bx + by * by + bz * bz)
    return np.array([r, r * u, r * v, r * w, e, bx, by, bz])


def mhd_flux(W):
    r, u, v, w, p, bx, by, bz = W
    U = mhd_cons(W)
    pt = p + 0.5 * (bx * bx + by * by + bz * bz)
    vb = u * bx + v * by + w * bz
    return np.array([r * u, r * u * u + pt - bx * bx, r * u * v - bx * by, r * u * w - bx * bz,
                     (U[4] + pt) * u - bx * vb, 0.0 * u, by * u - bx * v, bz * u - bx * w])


def fast_speed_x(W):
    r, u, v, w, p, bx, by, bz = W
    b2 = bx * bx + by * by + bz * bz
    a2 = GAMMA * p / r
    s = a2 + b2 / r
    return np.sqrt(0.5 * (s + np.sqrt(s * s - 4.0 * a2 * bx * bx / r)))


def davis(uL, cL, uR, cR):
    return np.minimum(uL - cL, uR - cR), np.maximum(uL + cL, uR + cR)


def hll_generic(FL, FR, UL, UR, sl, sr):
    F = (sr * FL - sl * FR + sl * sr * (UR - UL)) / (sr - sl)
    F = np.where(sl > 0, FL, F)
    return np.where(sr < 0, FR, F)


def hlle_mhd(WL, WR):
    sl, sr = davis(WL[1], fast_speed_x(WL), WR[1], fast_speed_x(WR))
    return hll_generic(mhd_flux(WL), mhd_flux(WR), mhd_cons(WL), mhd_cons(WR), sl, sr)


def hydro_cons(W):
    r, u, v, w, p = W
    return np.array([r, r * u, r * v, r * w, p / (GAMMA - 1.0) + 0.5 * r * (u * u + v * v + w * w)])


def hydro_flux(W):
    r, u, v, w, p = W
    U = hydro_cons(W)
    return np.array([r * u, r * u * u + p, r * u * v, r * u * w, (U[4] + p) * u])


def hll_hydro(WL, WR):
    cL, cR = np.sqrt(GAMMA * WL[4] / WL[0]), np.sqrt(GAMMA * WR[4] / WR[0])
    sl, sr = davis(WL[1], cL, WR[1], cR)
    return hll_generic(hydro_flux(WL), hydro_flux(WR), hydro_cons(WL), hydro_cons(WR), sl, sr)


def hllc_hydro(WL, WR):
    cL, cR = np.sqrt(GAMMA * WL[4] / WL[0]), np.sqrt(GAMMA * WR[4] / WR[0])
    sl, sr = davis(WL[1], cL, WR[1], cR)
    rL, uL, pL = WL[0], WL[1], WL[4]
    rR, uR, pR = WR[0], WR[1], WR[4]
    sm = (pR - pL + rL * uL * (sl - uL) - rR * uR * (sr - uR)) / (rL * (sl - uL) - rR * (sr - uR))

    def star(W, s):
        r, u, v, w, p = W
        U = hydro_cons(W)
        f = r * (s - u) / (s - sm)
        return np.array([f, f * sm, f * v, f * w, f * (U[4] / r + (sm - u) * (sm + p / (r * (s - u))))])
    FL, FR, UL, UR = hydro_flux(WL), hydro_flux(WR), hydro_cons(WL), hydro_cons(WR)
    FsL = FL + sl * (star(WL, sl) - UL)
    FsR = FR + sr * (star(WR, sr) - UR)
    F = np.where(sm >= 0, FsL, FsR)
    F = np.where(sl > 0, FL, F)
    return np.where(sr < 0, FR, F)


def hlld_mk2005(WL, WR):
    """Miyoshi & Kusano (2005); Bx := (BxL+BxR)/2 as in the reference (equal in the fixture)."""
    bx = 0.5 * (WL[5] + WR[5])
    sgn = np.copysign(1.0, bx)
    sl, sr = davis(WL[1], fast_speed_x(WL), WR[1], fast_speed_x(WR))

    def side(W, s):
        r, u, v, w, p, _bx, by, bz = W
        pt = p + 0.5 * (bx * bx + by * by + bz * bz)
        e = p / (GAMMA - 1.0) + 0.5 * r * (u * u + v * v + w * w) + 0.5 * (bx * bx + by * by + bz * bz)
        return dict(r=r, u=u, v=v, w=w, by=by, bz=bz, pt=pt, e=e, s=s,
                    U=np.array([r, r * u, r * v, r * w, e, bx + 0 * r, by, bz]))
    L, R = side(WL, sl), side(WR, sr)
    for K in (L, R):   # physical flux with the interface Bx
        r, u, v, w, by, bz, pt, e = (K[k] for k in "r u v w by bz pt e".split())
        K["F"] = np.array([r * u, r * u * u + pt - bx * bx, r * u * v - bx * by, r * u * w - bx * bz,
                           (e + pt) * u - bx * (u * bx + v * by + w * bz), 0 * r, by * u - bx * v, bz * u - bx * w])
    den = (sr - R["u"]) * R["r"] - (sl - L["u"]) * L["r"]
    sm = ((sr - R["u"]) * R["r"] * R["u"] - (sl - L["u"]) * L["r"] * L["u"] - R["pt"] + L["pt"]) / den          # (38)
    pts = ((sr - R["u"]) * R["r"] * L["pt"] - (sl - L["u"]) * L["r"] * R["pt"]
           + L["r"] * R["r"] * (sr - R["u"]) * (sl - L["u"]) * (R["u"] - L["u"])) / den                         # (41)
    for K in (L, R):
        r, u, v, w, by, bz, s = (K[k] for k in "r u v w by bz s".split())
        rs = r * (s - u) / (s - sm)                                                                          # (43)
        d = r * (s - u) * (s - sm) - bx * bx
        vs = v - bx * by * (sm - u) / d                                                                      # (44)
        ws = w - bx * bz * (sm - u) / d                                                                      # (46)
        bys = by * (r * (s - u) ** 2 - bx * bx) / d                                                          # (45)
        bzs = bz * (r * (s - u) ** 2 - bx * bx) / d                                                          # (47)
        es = ((s - u) * K["e"] - K["pt"] * u + pts * sm + bx * ((u * bx + v * by + w * bz) - (sm * bx + vs * bys + ws * bzs))) / (s - sm)   # (48)
        K.update(rs=rs, vs=vs, ws=ws, bys=bys, bzs=bzs, es=es,
                 Us=np.array([rs, rs * sm, rs * vs, rs * ws, es, bx + 0 * r, bys, bzs]))
    sqL, sqR = np.sqrt(L["rs"]), np.sqrt(R["rs"])
    ssl, ssr = sm - np.abs(bx) / sqL, sm + np.abs(bx) / sqR                                                  # (51)
    dd = sqL + sqR
    vss = (sqL * L["vs"] + sqR * R["vs"] + (R["bys"] - L["bys"]) * sgn) / dd                                 # (59)
    wss = (sqL * L["ws"] + sqR * R["ws"] + (R["bzs"] - L["bzs"]) * sgn) / dd                                 # (60)
    byss = (sqL * R["bys"] + sqR * L["bys"] + sqL * sqR * (R["vs"] - L["vs"]) * sgn) / dd                    # (61)
    bzss = (sqL * R["bzs"] + sqR * L["bzs"] + sqL * sqR * (R["ws"] - L["ws"]) * sgn) / dd                    # (62)
    vbss = sm * bx + vss * byss + wss * bzss
    for K, sq, pm in ((L, sqL, -1.0), (R, sqR, +1.0)):
        vbs = sm * bx + K["vs"] * K["bys"] + K["ws"] * K["bzs"]
        ess = K["es"] + pm * sq * (vbs - vbss) * sgn                                                         # (63)
        K["Uss"] = np.array([K["rs"], K["rs"] * sm, K["rs"] * vss, K["rs"] * wss, ess, bx + 0 * sm, byss, bzss])
    FsL = L["F"] + sl * (L["Us"] - L["U"])                                                                   # (64)
    FsR = R["F"] + sr * (R["Us"] - R["U"])
    FssL = L["F"] + ssl * L["Uss"] - (ssl - sl) * L["Us"] - sl * L["U"]                                      # (65)
    FssR = R["F"] + ssr * R["Uss"] - (ssr - sr) * R["Us"] - sr * R["U"]
    # region order of the reference: SL>0, SR<0, SL*>=0, SR*<=0, SM>=0, else
    F = np.where(sm >= 0, FssL, FssR)
    F = np.where(ssr <= 0, FsR, F)
    F = np.where(ssl >= 0, FsL, F)
    F = np.where(sr < 0, mhd_flux(WR), F)
    F = np.where(sl > 0, mhd_flux(WL), F)
    F[5] = 0.0
    region = np.where(sl > 0, 0, np.where(sr < 0, 5, np.where(ssl >= 0, 1, np.where(ssr <= 0, 4, np.where(sm >= 0, 2, 3)))))
    return F, region


# ----------------------------------------------------------------------------------------
def sod_exact(x, t, x0=0.5, left=(1.0, 0.0, 1.0), right=(0.125, 0.0, 0.1), gamma=1.4):
    """Exact solution of the Sod problem (Toro 2009, sect. 4.2-4.5): left rarefaction, contact, right shock."""
    rl, ul, pl = left
    rr, ur, pr = right
    g = gamma
    al, ar = np.sqrt(g * pl / rl), np.sqrt(g * pr / rr)

    def f(p):
        fl = 2 * al / (g - 1) * ((p / pl) ** ((g - 1) / (2 * g)) - 1)
        A, B = 2 / ((g + 1) * rr), (g - 1) / (g + 1) * pr
        fr = (p - pr) * np.sqrt(A / (p + B))
        return fl + fr + ur - ul
    lo, hi = 1e-8, 10.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if f(lo) * f(mid) <= 0:
            hi = mid
        else:
            lo = mid
    ps = 0.5 * (lo + hi)
    us = ul - 2 * al / (g - 1) * ((ps / pl) ** ((g - 1) / (2 * g)) - 1)
    rsl = rl * (ps / pl) ** (1 / g)
    rsr = rr * ((ps / pr + (g - 1) / (g + 1)) / ((g - 1) / (g + 1) * ps / pr + 1))
    asl = al * (ps / pl) ** ((g - 1) / (2 * g))
    sh = ur + ar * np.sqrt((g + 1) / (2 * g) * ps / pr + (g - 1) / (2 * g))
    xi = (x - x0) / t
    rho, u, p = np.empty_like(x), np.empty_like(x), np.empty_like(x)
    for n, s in enumerate(xi):
        if s < ul - al:
            rho[n], u[n], p[n] = rl, ul, pl
        elif s < us - asl:
            c = 2 / (g + 1) + (g - 1) / ((g + 1) * al) * (ul - s)
            rho[n], u[n], p[n] = rl * c ** (2 / (g - 1)), 2 / (g + 1) * (al + (g - 1) / 2 * ul + s), pl * c ** (2 * g / (g - 1))
        elif s < us:
            rho[n], u[n], p[n] = rsl, us, ps
        elif s < sh:
            rho[n], u[n], p[n] = rsr, us, ps
        else:
            rho[n], u[n], p[n] = rr, ur, pr
    return rho, u, p, dict(pstar=ps, ustar=us, rho_star_l=rsl, rho_star_r=rsr, shock_speed=sh)


def ot_first_dt(nx=512, ny=512, nz=2, zmax=2.0 / 512, cfl=0.2, n_iter=10):
    """dt of iteration 1 of the shipped Orszag-Tang run (OT/parameters.f90), from the analytic
    ICs (OT/orzag_tang.f90:14-70) and get_timestep's rule (src/hydro_core.f90:644-682):
    min over cells and axes of d/(|v|+c_fast,axis), times cfl * 2^-(n_iter+1-1)."""
    pi = np.arccos(-1.0)
    dx, dy, dz = 1.0 / nx, 1.0 / ny, zmax / nz
    x = ((np.arange(1, nx + 1) + 0.5) * dx)[:, None]
    y = ((np.arange(1, ny + 1) + 0.5) * dy)[None, :]
    rho, p = 25.0 / (36.0 * pi), 5.0 / (12.0 * pi)
    vx, vy = -np.sin(2 * pi * y) + 0 * x, np.sin(2 * pi * x) + 0 * y
    bx, by = -np.sin(2 * pi * y) / np.sqrt(4 * pi) + 0 * x, np.sin(4 * pi * x) / np.sqrt(4 * pi) + 0 * y
    b2 = bx * bx + by * by
    s = GAMMA * p + b2

    def cf(bn):
        return np.sqrt(0.5 * (s + np.sqrt(s * s - 4 * GAMMA * p * bn * bn)) / rho)
    dtp = min((dx / (np.abs(vx) + cf(bx))).min(), (dy / (np.abs(vy) + cf(by))).min(), (dz / (0.0 + cf(0 * bx))).min())
    return cfl * 2.0 ** (-(n_iter + 1 - 1)) * dtp, dtp


def main():
    rng = np.random.default_rng(20261017)
    n = 1024
    # MHD states with continuous Bx; velocities up to Mach ~2.5 so that every HLLD region occurs
    def mhd_states():
        W = np.empty((8, n))
        W[0] = rng.uniform(0.2, 3.0, n); W[1:4] = rng.normal(0, 1.2, (3, n)); W[4] = rng.uniform(0.1, 3.0, n)
        W[5:8] = rng.normal(0, 0.8, (3, n))
        return W
    WL, WR = mhd_states(), mhd_states()
    WR[5] = WL[5]
    WL[1, :32] += 6.0; WR[1, :32] += 6.0          # supersonic to the right (UL)
    WL[1, 32:64] -= 6.0; WR[1, 32:64] -= 6.0    # supersonic to the left (UR)
    Fd, region = hlld_mk2005(WL, WR)
    Fe = hlle_mhd(WL, WR)
    HL, HR = WL[:5].copy(), WR[:5].copy()
    np.savez_compressed(os.path.join(HERE, "published_riemann.npz"), gamma=GAMMA, cv=CV, WL=WL, WR=WR,
                        hlld=Fd, hlld_region=region, hlle=Fe, hll=hll_hydro(HL, HR), hllc=hllc_hydro(HL, HR))
    print("published_riemann.npz: HLLD region counts (UL, UL*, UL**, UR**, UR*, UR):", np.bincount(region, minlength=6))

    xs = (np.arange(400) + 0.5) / 400
    rho, u, p, info = sod_exact(xs, 0.2)
    np.savez_compressed(os.path.join(HERE, "published_sod.npz"), x=xs, t=0.2, rho=rho, u=u, p=p, gamma=1.4, **info)

    dt1, dtp = ot_first_dt()
    np.savez(os.path.join(HERE, "published_ot_dt.npz"), dt_first=dt1, dtp=dtp)
    print(f"published_ot_dt.npz: first dt {dt1:.13e}  (SURVEY 8(c) KAT 6 quotes 1.7222826491e-7)")

    # ---- oracle_* fixtures: produced by the C++ oracle ----
    from guacho_b200.config import Params, SOLVER_HLLD, SOLVER_HLLC
    from tests.oracle_lib import U
    from tests.util import global_ic, oracle_from_ic
    for name, p, problem, nsteps in (
            ("oracle_ot_hlld_cd_24x20x4", Params(nxtot=24, nytot=20, nztot=4, zmax=4.0 / 24), "ot", 3),
            ("oracle_random_hlld_cd_16x12x10", Params(nxtot=16, nytot=12, nztot=10, zmax=1.0), "random", 3),
            ("oracle_random_hllc_16x12x10", Params(nxtot=16, nytot=12, nztot=10, zmax=1.0, mhd=False, riemann_solver=SOLVER_HLLC, enable_flux_cd=False), "random", 3)):
        g = global_ic(p, problem)
        o = oracle_from_ic(p, g, threads=1)
        dts = o.advance(nsteps)
        u = o.get_block(0, U)[..., 2:-2, 2:-2, 2:-2]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), u0=g, u=u, dts=np.array(dts), nsteps=nsteps, problem=problem,
                            params=np.array([p.nxtot, p.nytot, p.nztot]), zmax=p.zmax)
        print(name, "dts", dts)


if __name__ == "__main__":
    main()
