"""Initial conditions — the host-side `user_mod.initial_conditions(u)` plugins.

Each function fills one block's conserved array in the reference layout
``u(neq, nxmin:nxmax, nymin:nymax, nzmin:nzmax)`` (Fortran order, ghosts included),
exactly what the reference's ``initial_conditions`` is asked to do
(``OT/user_mod.f90:58-66``).  They run on the host, once, outside the hot path.
"""
from __future__ import annotations

import numpy as np

from .config import Params, NGHOST


def _block_index_grids(p: Params, coords):
    """Fortran local indices i=nxmin..nxmax etc. shifted to global cell numbers."""
    i = np.arange(1 - NGHOST, p.nx + NGHOST + 1, dtype=np.float64) + coords[0] * p.nx
    j = np.arange(1 - NGHOST, p.ny + NGHOST + 1, dtype=np.float64) + coords[1] * p.ny
    k = np.arange(1 - NGHOST, p.nz + NGHOST + 1, dtype=np.float64) + coords[2] * p.nz
    return i, j, k


def prim_to_cons(p: Params, rho, vx, vy, vz, pres, bx=None, by=None, bz=None, passives=()):
    """prim2u (src/hydro_core.f90:331-383) on whole arrays; returns Fortran-ordered u."""
    shape = np.broadcast(rho, vx, vy, vz, pres).shape
    u = np.zeros((p.neq,) + shape, dtype=np.float64, order="F")
    u[0] = rho
    u[1] = rho * vx
    u[2] = rho * vy
    u[3] = rho * vz
    u[4] = 0.5 * rho * (vx * vx + vy * vy + vz * vz) + p.cv * pres
    if p.bfield:
        if p.mhd:
            u[4] = u[4] + 0.5 * (bx * bx + by * by + bz * bz)
        u[5], u[6], u[7] = bx, by, bz
    for q, s in enumerate(passives):
        u[p.neqdyn + q] = s
    return u


def orszag_tang(p: Params, coords=(0, 0, 0), rsc: float = 1.0) -> np.ndarray:
    """Orszag-Tang vortex, restating OT/orzag_tang.f90:14-70 (rho=25/36pi, p=5/12pi)."""
    pi = np.arccos(-1.0)
    twopi = 2.0 * pi
    rho = 25.0 / (36.0 * pi)
    pres = 5.0 / (12.0 * pi)
    i, j, k = _block_index_grids(p, coords)
    x = ((i + 0.5) * p.dx * rsc)[:, None, None]
    y = ((j + 0.5) * p.dy * rsc)[None, :, None]
    # the field does not depend on z: evaluate one (x, y) plane and broadcast it along z on assignment
    one = np.ones((i.size, j.size, 1))
    vx = -np.sin(y * twopi) * one
    vy = np.sin(x * twopi) * one
    vz = 0.0 * one
    u = np.zeros(p.block_shape(), dtype=np.float64, order="F")
    u[0] = rho
    u[1] = rho * vx
    u[2] = rho * vy
    u[3] = rho * vz
    if p.bfield:
        bx = -np.sin(y * twopi) / np.sqrt(4 * pi) * one
        by = np.sin(2.0 * x * twopi) / np.sqrt(4 * pi) * one
        bz = 0.0 * one
        u[4] = 0.5 * rho * (vx ** 2 + vy ** 2 + vz ** 2) + p.cv * pres + 0.5 * (bx ** 2 + by ** 2 + bz ** 2)
        u[5], u[6], u[7] = bx, by, bz
    else:
        u[4] = 0.5 * rho * (vx ** 2 + vy ** 2 + vz ** 2) + p.cv * pres
    for q in range(p.npas):
        u[p.neqdyn + q] = rho * (0.5 + 0.25 * q)
    return u


def mhd_blast(p: Params, coords=(0, 0, 0), r0: float = 0.1, p_in: float = 10.0, p_out: float = 0.1) -> np.ndarray:
    """3-D MHD blast wave (builder-defined): rho=1, p=0.1 (10 inside
    r<r0 of the box centre), B=(1/sqrt2, 1/sqrt2, 0), v=0."""
    i, j, k = _block_index_grids(p, coords)
    x = ((i - 0.5) * p.dx - 0.5 * p.dx * p.nxtot)[:, None, None]
    y = ((j - 0.5) * p.dy - 0.5 * p.dy * p.nytot)[None, :, None]
    z = ((k - 0.5) * p.dz - 0.5 * p.dz * p.nztot)[None, None, :]
    r = np.sqrt(x * x + y * y + z * z)
    one = np.ones(r.shape)
    pres = np.where(r < r0, p_in, p_out)
    b = 1.0 / np.sqrt(2.0)
    return prim_to_cons(p, one, 0 * one, 0 * one, 0 * one, pres, b * one, b * one, 0 * one,
                        passives=[one * (0.1 + 0.2 * q) for q in range(p.npas)])


def smooth_random(p: Params, coords=(0, 0, 0), seed: int = 12345, kmax: int = 8, amp: float = 0.2, vamp: float = 0.5) -> np.ndarray:
    """Branch-coverage field (builder-defined): periodic low-pass
    (|k|<=kmax) random perturbations, rho=1+amp*xi1, p=1+amp*xi2, v=0.5*xi3..5,
    B=0.5*xi6..8.  Built from a fixed set of Fourier modes so that any block of any
    decomposition evaluates the same global function."""
    rng = np.random.default_rng(seed)
    nmodes = 24
    nfields = 8 + p.npas
    kvec = rng.integers(-kmax, kmax + 1, size=(nfields, nmodes, 3))
    phase = rng.uniform(0, 2 * np.pi, size=(nfields, nmodes))
    ampl = rng.normal(size=(nfields, nmodes)) / np.sqrt(nmodes)
    i, j, k = _block_index_grids(p, coords)
    X = ((i - 0.5) / p.nxtot)[:, None, None]
    Y = ((j - 0.5) / p.nytot)[None, :, None]
    Z = ((k - 0.5) / p.nztot)[None, None, :]

    def xi(f):
        out = np.zeros((i.size, j.size, k.size))
        for m in range(nmodes):
            kx, ky, kz = kvec[f, m]
            out += ampl[f, m] * np.cos(2 * np.pi * (kx * X + ky * Y + kz * Z) + phase[f, m])
        return np.clip(out, -2.0, 2.0)

    rho = 1.0 + amp * xi(0)
    pres = 1.0 + amp * xi(1)
    vx, vy, vz = vamp * xi(2), vamp * xi(3), vamp * xi(4)      # vamp >~ 3: supersonic interfaces (sl > 0 / sr < 0 branches of the solvers)
    bx, by, bz = 0.5 * xi(5), 0.5 * xi(6), 0.5 * xi(7)
    pas = [rho * (0.5 + 0.2 * xi(8 + q)) for q in range(p.npas)]
    return prim_to_cons(p, rho, vx, vy, vz, pres, bx, by, bz, passives=pas)


PROBLEMS = {"ot": orszag_tang, "blast": mhd_blast, "random": smooth_random}
