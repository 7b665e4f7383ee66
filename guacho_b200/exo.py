"""Host-side plugin of the EXO problem — the exoplanet-wind set-up the reference ships in ``EXO/`` (BASELINE.json configs[3]).

This is the ``user_mod`` of that problem on the host side of the C ABI: what ``EXO/parameters.f90`` (scalings, switches),
``EXO/exoplanet.f90:59-112`` (``init_exo``), ``EXO/user_mod.f90:58-121`` (``initial_conditions``) and the per-call state of
``impose_exo`` / ``get_user_source_terms`` (``exoplanet.f90:137-144``, ``user_mod.f90:174-187``) compute on the host.  The
per-cell work of those two routines runs on the device as functors (``gx_set_wind_spheres`` / ``gx_set_gravity_points``),
re-positioned from the ``gx_register_bc_hook`` hook at every boundary call, where the reference moves the planet.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .config import Params, EOS_H_RATE, BC_OUTFLOW, COOL_H, NGHOST
from .lib import WindSphere

# src/constants.f90:30-50
PI = float(np.arccos(-1.0))
AMH, RG, GGRAV = 1.66e-24, 8.3145e7, 6.67259e-8
MSUN, RSUN, MJUP, RJUP = 1.99e33, 6.955e10, 1.898e30, 7.1492e9
AU, DAY, YR = 1.496e13, 86400.0, 3.1536e7


def exo_params(nx: int = 400, ny: int = 100, nz: int = 400, cooling: bool = True, **kw) -> Params:
    """EXO/parameters.f90 as shipped: 400 x 100 x 400, HLLD + full MHD + flux-CD, 2 passives, EOS_H_RATE, COOL_H, outflow walls
    + user boundary, user (gravity) source, eta = 0.01, cfl = 0.4."""
    cv = 1.5
    gamma = (cv + 1.0) / cv
    o = BC_OUTFLOW
    p = Params(nxtot=nx, nytot=ny, nztot=nz, xmax=1.0, ymax=0.25, zmax=1.0, mhd=True, npas=2, eq_of_state=EOS_H_RATE,
               enable_flux_cd=True, user_source_terms=True, bc_left=o, bc_right=o, bc_bottom=o, bc_top=o, bc_out=o, bc_in=o,
               bc_user=True, cv=cv, Tempsc=1.0e4 * gamma, cfl=0.4, eta=0.01, **kw)
    if cooling:
        p = p.replace(cooling=COOL_H, tsc=Scalings.of(p).tsc)
    return p


@dataclass
class Scalings:
    """EXO/parameters.f90:154-170"""
    rsc: float
    rhosc: float
    Tempsc: float
    vsc2: float
    tsc: float
    bsc: float

    @staticmethod
    def of(p: Params) -> "Scalings":
        T0, mu = 1.0e4, 1.0
        rsc = 0.3 * AU / 1.0
        rhosc = AMH * mu
        vsc2 = p.gamma * RG * T0 / mu
        return Scalings(rsc, rhosc, T0 * p.gamma, vsc2, rsc / np.sqrt(vsc2), float(np.sqrt(4.0 * PI * rhosc * vsc2)))


class Exo:
    """init_exo (EXO/exoplanet.f90:59-112): star / planet wind parameters in code units and the orbit."""

    def __init__(self, p: Params):
        self.p = p
        s = self.s = Scalings.of(p)
        self.MassS, self.MassP = 1.1 * MSUN, 0.67 * MJUP
        amdot = 2.0e-14 * MSUN / YR
        self.TSW = 1.56e6 / s.Tempsc
        RSW, VSW = 1.2 * RSUN, 1.0e5
        self.dsw = ((amdot / RSW) / (4 * PI * RSW * VSW)) / s.rhosc
        self.RSW, self.VSW = RSW / s.rsc, VSW / np.sqrt(s.vsc2)
        self.bsw = 1.0 / s.bsc
        ampdot = 1.0e10
        self.TPW = 1.0e4 / s.Tempsc
        RPW, VPW = 3.0 * 1.38 * RJUP, 10.0e5
        self.dpw = ((ampdot / RPW) / (4 * PI * RPW * VPW)) / s.rhosc
        self.RPW, self.VPW = RPW / s.rsc, VPW / np.sqrt(s.vsc2)
        self.bpw = 0.04 / s.bsc
        self.rorb = 0.047 * AU / s.rsc
        self.torb = 3.52 * DAY / s.tsc
        self.omegap = 2.0 * PI / self.torb
        self.phi = -25.0 * PI / 180.0

    # -- the planet at `time` (exoplanet.f90:137-144) --
    def planet(self, time: float):
        a = self.omegap * time + self.phi
        return (self.rorb * np.cos(a), 0.0, self.rorb * np.sin(a)), (-self.omegap * self.rorb * np.sin(a), 0.0, self.omegap * self.rorb * np.cos(a))

    def spheres(self, time: float):
        (xp, _yp, zp), (vx, _vy, vz) = self.planet(time)
        star = WindSphere(xc=0, yc=0, zc=0, radius=self.RSW, vwind=self.VSW, dens=self.dsw, tfac=1.0, temp=self.TSW, vbx=0, vby=0, vbz=0,
                          bdip=self.bsw, pas=(C.c_double * 4)(0.0001, 1.0, 0, 0))
        planet = WindSphere(xc=xp, yc=0, zc=zp, radius=self.RPW, vwind=self.VPW, dens=self.dpw, tfac=1.8, temp=self.TPW, vbx=vx, vby=0.0, vbz=vz,
                            bdip=self.bpw, pas=(C.c_double * 4)(0.2, -1.0, 0, 0))
        return [star, planet]

    def gravity(self, time: float):
        """GM and positions of get_user_source_terms (user_mod.f90:174-187)."""
        (xp, _yp, zp), _v = self.planet(time)
        s = self.s
        return [0.3 * GGRAV * self.MassS / s.rsc / s.vsc2, GGRAV * self.MassP / s.rsc / s.vsc2], [[0.0, 0.0, 0.0], [xp, 0.0, zp]]

    def attach(self, block, time: float = 0.0) -> None:
        """Device functors + the hook that moves the planet at every impose_user_bc application."""
        def place(_order, t):
            block.set_wind_spheres(self.spheres(t))
            block.set_gravity_points(*self.gravity(t))
        block.register_bc_hook(place)
        place(0, time)
        block.set_time(time)

    # -- initial_conditions (user_mod.f90:58-121): the stellar wind everywhere, then impose_exo(u, 0) --
    def initial_conditions(self, coords=(0, 0, 0)) -> np.ndarray:
        p = self.p
        i = np.arange(1 - NGHOST, p.nx + NGHOST + 1, dtype=np.float64) + coords[0] * p.nx
        j = np.arange(1 - NGHOST, p.ny + NGHOST + 1, dtype=np.float64) + coords[1] * p.ny
        k = np.arange(1 - NGHOST, p.nz + NGHOST + 1, dtype=np.float64) + coords[2] * p.nz
        x = ((i - p.nxtot // 2 + 0.5) * p.dx)[:, None, None]
        y = ((j - p.nytot // 2 + 0.5) * p.dy)[None, :, None]
        z = ((k - p.nztot // 2 + 0.5) * p.dz)[None, None, :]
        u = np.zeros(p.block_shape(), dtype=np.float64, order="F")
        rads = np.sqrt(x ** 2 + y ** 2 + z ** 2)
        dens = self.dsw * self.RSW ** 2 / rads ** 2
        u[0] = dens
        u[1], u[2], u[3] = dens * (self.VSW * x / rads), dens * (self.VSW * y / rads), dens * (self.VSW * z / rads)
        cpi = self.bsw * (self.RSW / rads) ** 3 / (2.0 * rads ** 2)
        u[5], u[6], u[7] = 3.0 * y * x * cpi, (3.0 * y ** 2 - rads ** 2) * cpi, 3.0 * y * z * cpi
        u[4] = 0.5 * dens * self.VSW ** 2 + p.cv * dens * self.TSW + 0.5 * (u[5] ** 2 + u[6] ** 2 + u[7] ** 2)
        u[8], u[9] = 0.0001 * dens, dens
        # impose_exo(u, 0.): exoplanet.f90:125-266 — star first, `else if` planet
        (xp, _yp, zp), (vxo, vyo, vzo) = self.planet(0.0)
        for (xc, zc, R, V, D, tfac, T, b0, vb, pas, only) in (
                (0.0, 0.0, self.RSW, self.VSW, self.dsw, 1.0, self.TSW, self.bsw, (0.0, 0.0, 0.0), (0.0001, 1.0), None),
                (xp, zp, self.RPW, self.VPW, self.dpw, 1.8, self.TPW, self.bpw, (vxo, vyo, vzo), (0.2, -1.0), "not star")):
            xl, yl, zl = x - xc, y + 0.0 * x, z - zc
            rad = np.sqrt(xl ** 2 + yl ** 2 + zl ** 2)
            m = rad <= R
            if only:
                m &= ~(rads <= self.RSW)
            rad = np.where(rad == 0.0, p.dx * 0.10, rad)
            velx, vely, velz = vb[0] + V * xl / rad, vb[1] + V * yl / rad, vb[2] + V * zl / rad
            c = b0 * (R / rad) ** 3 / (2.0 * rad ** 2)
            bx, by, bz = 3.0 * yl * xl * c, (3.0 * yl ** 2 - rad ** 2) * c, 3.0 * yl * zl * c
            e = 0.5 * D * (velx ** 2 + vely ** 2 + velz ** 2) + p.cv * D * tfac * T + 0.5 * (bx ** 2 + by ** 2 + bz ** 2)
            for q, val in ((0, D + 0 * rad), (1, D * velx), (2, D * vely), (3, D * velz), (4, e), (5, bx), (6, by), (7, bz),
                           (8, pas[0] * D + 0 * rad), (9, pas[1] * D + 0 * rad)):
                u[q] = np.where(m, val, u[q])
        return u
