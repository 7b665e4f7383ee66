"""Shared helpers for the parity tests (oracle vs CUDA path through the C ABI)."""
from __future__ import annotations

import numpy as np

from guacho_b200 import problems
from guacho_b200.config import Params
from tests.oracle_lib import Oracle, U


def global_ic(p: Params, problem: str = "ot", **kw) -> np.ndarray:
    """IC for the whole domain as ONE block with ghosts: (neq, nxtot+4, nytot+4, nztot+4)."""
    p1 = p.replace(MPI_NBX=1, MPI_NBY=1, MPI_NBZ=1)
    return problems.PROBLEMS[problem](p1, (0, 0, 0), **kw)


def block_ic(p: Params, g: np.ndarray, coords) -> np.ndarray:
    """Cut one block (with its own ghosts) out of a global-with-ghosts IC array."""
    i0, j0, k0 = coords[0] * p.nx, coords[1] * p.ny, coords[2] * p.nz
    return np.asfortranarray(g[:, i0:i0 + p.nx + 4, j0:j0 + p.ny + 4, k0:k0 + p.nz + 4])


def oracle_from_ic(p: Params, g: np.ndarray, threads: int = 4) -> Oracle:
    o = Oracle(p, threads=threads)
    o.scatter_u(g)
    o.start()
    return o


def rel_err_per_var(a: np.ndarray, ref: np.ndarray) -> np.ndarray:
    """max|a-ref| / max|ref| per conserved variable (the parity gate of SURVEY §8(c))."""
    out = np.zeros(a.shape[0])
    for q in range(a.shape[0]):
        den = np.abs(ref[q]).max()
        num = np.abs(a[q] - ref[q]).max()
        out[q] = num / den if den > 0 else num
    return out


def interior(a: np.ndarray) -> np.ndarray:
    return a[..., 2:-2, 2:-2, 2:-2]


def tc_scalings(rhosc: float = 1e-15, **kw) -> dict:
    """cgs scalings for the thermal-conduction tests (pattern of parameters.f90:152-170): a hot (1e6 K), thin plasma in a box
    of 1e10 cm, for which the Spitzer time scale of a ~16-cell grid is comparable to the hydro step (rhosc = 1e-15) or well
    below it (smaller rhosc: super-time-stepping with several substeps)."""
    mu, Rg, gamma, T0, rsc = 0.6, 8.3145e7, 5.0 / 3.0, 1.0e6, 1e10
    vsc2 = gamma * Rg * T0 / mu
    d = dict(rsc=rsc, rhosc=rhosc, vsc2=vsc2, tsc=rsc / np.sqrt(vsc2), bsc=float(np.sqrt(4 * np.pi * rhosc * vsc2)), mu=mu, Tempsc=T0 * gamma)
    d.update(kw)
    return d
