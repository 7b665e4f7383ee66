"""Cartesian block decomposition — the replacement for the reference's compile-time
``MPI_NBX/NBY/NBZ`` + ``mpi_cart_create/coords/shift`` (src/init.f90:61-110).

Pure host logic (no GPU): rank <-> coords map, neighbour ranks, slab/pencil choice.
The rank order is MPI's row-major one, ``rank = (cx*NBY + cy)*NBZ + cz``, which is what
the reference's Python reader assumes (py/guacho_utils.py:104-118).
"""
from __future__ import annotations

from typing import Tuple

from .config import Params, BC_PERIODIC


def rank_of(coords, nb) -> int:
    return (coords[0] * nb[1] + coords[1]) * nb[2] + coords[2]


def coords_of(rank: int, nb) -> Tuple[int, int, int]:
    cz = rank % nb[2]
    cy = (rank // nb[2]) % nb[1]
    cx = rank // (nb[1] * nb[2])
    return (cx, cy, cz)


def periodic_dims(p: Params):
    return (p.bc_left == BC_PERIODIC and p.bc_right == BC_PERIODIC,
            p.bc_bottom == BC_PERIODIC and p.bc_top == BC_PERIODIC,
            p.bc_out == BC_PERIODIC and p.bc_in == BC_PERIODIC)


def neighbors(p: Params, coords):
    """(left, right, bottom, top, out, in) ranks; -1 = MPI_PROC_NULL (mpi_cart_shift,
    src/init.f90:108-110: left/bottom/out are the -x/-y/-z sources, right/top/in the +dests)."""
    nb = (p.MPI_NBX, p.MPI_NBY, p.MPI_NBZ)
    per = periodic_dims(p)
    out = []
    for d in range(3):
        for step in (-1, +1):
            c = list(coords)
            c[d] += step
            if c[d] < 0 or c[d] >= nb[d]:
                if not per[d]:
                    out.append(-1)
                    continue
                c[d] %= nb[d]
            out.append(rank_of(c, nb))
    return tuple(out)


def slab_decomposition(nranks: int, axis: int = 2) -> Tuple[int, int, int]:
    """Slabs along `axis` (default z: the slowest index of the device SoA layout, so ghost
    planes are contiguous)."""
    nb = [1, 1, 1]
    nb[axis] = nranks
    return tuple(nb)


def pencil_decomposition(nranks: int) -> Tuple[int, int, int]:
    """y-z pencils: the most square factorisation nranks = nby*nbz with nby <= nbz."""
    best = (1, nranks)
    f = 1
    while f * f <= nranks:
        if nranks % f == 0:
            best = (f, nranks // f)
        f += 1
    return (1, best[0], best[1])


def choose_decomposition(p: Params, nranks: int, min_thickness: int = 32) -> Tuple[int, int, int]:
    """z-slabs unless they would be thinner than `min_thickness` planes, then y-z pencils."""
    if p.nztot % nranks == 0 and p.nztot // nranks >= min_thickness:
        return slab_decomposition(nranks, 2)
    nb = pencil_decomposition(nranks)
    if p.nytot % nb[1] == 0 and p.nztot % nb[2] == 0:
        return nb
    if p.nztot % nranks == 0:
        return slab_decomposition(nranks, 2)
    raise ValueError(f"cannot split {p.nxtot}x{p.nytot}x{p.nztot} over {nranks} ranks")


def halo_bytes_per_step(p: Params) -> int:
    """Bytes one block sends per tstep: boundaryII (2 layers of up), boundaryI (1 layer of u) and,
    with flux-CD, two 1-layer E exchanges, over faces that have a real neighbour."""
    nb = (p.MPI_NBX, p.MPI_NBY, p.MPI_NBZ)
    n = (p.nx, p.ny, p.nz)
    total = 0
    for d in range(3):
        if nb[d] == 1:
            continue
        t = [n[a] for a in range(3) if a != d]
        total += 2 * p.neq * 2 * (t[0] + 4) * (t[1] + 4)          # up, 2 layers, both faces
        total += 2 * p.neq * 1 * (t[0] + 2) * (t[1] + 2)          # u, 1 layer
        if p.enable_flux_cd:
            total += 2 * 2 * 3 * (t[0] + 2) * (t[1] + 2)          # E, 1 layer, twice per step
    return total * 8
