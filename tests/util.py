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
