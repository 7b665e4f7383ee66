"""BIN dumps in the reference's format — output, restart and parity artefact.

``write_bin`` reproduces ``write_header`` + ``write_BIN`` of ``src/Out_BIN_Module.f90:40-102,129-165``
byte for byte (DOUBLEP build: ASCII banner, 0xFF marker, 'd', int32/float64 binary header, then
``u(neq, nx+2g, ny+2g, nz+2g)`` column-major WITH ghosts), one file per block named
``points<rank:03d>.<itprint:03d>.bin`` so that the reference's own readers (``py/guacho_utils.py:8-119``)
and its warm start (``src/init.f90:436-471``) work unchanged on GPU output.  ``read_bin`` is the warm-start
reader.  ``write_divb`` restates the div B dump (``Out_BIN_Module.f90:205-237``).
Host-side I/O only: nothing here is on the timed path.
"""
from __future__ import annotations

import os
import struct
from typing import Tuple

import numpy as np

from .config import Params, NGHOST

LF = b"\n"


def _es(x: float) -> str:
    return f"{x:10.3E}"            # Fortran es10.3


def header_bytes(p: Params, coords, neq_out: int, nghost_out: int, rsc: float = 1.0, vsc: float = 1.0, rhosc: float = 1.0) -> bytes:
    """write_header (src/Out_BIN_Module.f90:40-102)."""
    nx, ny, nz = p.nx, p.ny, p.nz
    x0, y0, z0 = coords[0] * nx, coords[1] * ny, coords[2] * nz
    lines = [
        "**************** Output for Guacho v1.3****************",
        f"Dimensions    : {nx} {ny} {nz}",
        "Spacings      : " + _es(p.dx) + _es(p.dy) + _es(p.dz),
        f"Block Origin, cells    : {x0} {y0} {z0}",
        f"MPI blocks (X, Y, Z)   : {p.MPI_NBX} {p.MPI_NBY} {p.MPI_NBZ}",
        f"Number of Equations/dynamical ones  {neq_out}/{p.neqdyn}",
        f"Number of Ghost Cells  {nghost_out}",
        "Scalings",
        "r_sc: " + _es(rsc) + " v_sc: " + _es(vsc) + " rho_sc: " + _es(rhosc),
        f"Specfic heat at constant volume Cv: {p.cv:7.2f}",
        "Double precision 8 byte floats",
        "*******************************************************",
    ]
    out = b"".join(s.rstrip().encode("ascii") + LF for s in lines)       # trim(cbuffer), lf
    out += b"\xff" + LF + b"d"
    out += struct.pack("<3i", nx, ny, nz) + struct.pack("<3d", p.dx, p.dy, p.dz) + struct.pack("<3i", x0, y0, z0)
    out += struct.pack("<3i", p.MPI_NBX, p.MPI_NBY, p.MPI_NBZ) + struct.pack("<2i", neq_out, p.neqdyn) + struct.pack("<i", nghost_out)
    out += struct.pack("<3d", rsc, vsc, rhosc) + struct.pack("<d", p.cv)
    return out


def bin_name(outputpath: str, rank: int, itprint: int, base: str = "points") -> str:
    """MPI build naming (Out_BIN_Module.f90:129-130); the readers look for <path>BIN/<base><rank>.<it>.bin."""
    return os.path.join(outputpath, "BIN", f"{base}{rank:03d}.{itprint:03d}.bin")


def write_bin(outputpath: str, u: np.ndarray, p: Params, coords, rank: int, itprint: int, **scal) -> str:
    """write_BIN: header + u(:,:,:,:) with ghosts (Out_BIN_Module.f90:152-165)."""
    if u.shape != p.block_shape():
        raise ValueError(f"u has shape {u.shape}, expected {p.block_shape()}")
    path = bin_name(outputpath, rank, itprint)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(header_bytes(p, coords, p.neq, NGHOST, **scal))
        f.write(np.asfortranarray(u, dtype="<f8").tobytes(order="F"))
    return path


def divergence_b(u: np.ndarray, p: Params) -> np.ndarray:
    """Central-difference div B over the physical cells (Out_BIN_Module.f90:212-222)."""
    g = NGHOST
    c = (slice(g, -g),) * 3
    bx, by, bz = u[5], u[6], u[7]
    return ((bx[g + 1:bx.shape[0] - g + 1, g:-g, g:-g] - bx[g - 1:-g - 1, g:-g, g:-g]) / (2.0 * p.dx)
            + (by[g:-g, g + 1:by.shape[1] - g + 1, g:-g] - by[g:-g, g - 1:-g - 1, g:-g]) / (2.0 * p.dy)
            + (bz[g:-g, g:-g, g + 1:bz.shape[2] - g + 1] - bz[g:-g, g:-g, g - 1:-g - 1]) / (2.0 * p.dz))


def write_divb(outputpath: str, u: np.ndarray, p: Params, coords, rank: int, itprint: int, **scal) -> str:
    """divB-<rank>.<it>.bin: same header with neq = 1, nghost = 0, then div B (Out_BIN_Module.f90:205-237)."""
    path = os.path.join(outputpath, "BIN", f"divB-{rank:03d}.{itprint:03d}.bin")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(header_bytes(p, coords, 1, 0, **scal))
        f.write(np.asfortranarray(divergence_b(u, p), dtype="<f8").tobytes(order="F"))
    return path


def read_bin(path: str) -> Tuple[np.ndarray, dict]:
    """Warm-start reader (src/init.f90:448-468): skip the ASCII banner up to the 0xFF byte, read the binary
    header and u(:,:,:,:) including ghosts.  Returns (u, header)."""
    with open(path, "rb") as f:
        raw = f.read()
    k = raw.index(b"\xff")                      # init.f90:452-455
    off = k + 2                                 # 0xFF, LF
    kind = raw[off:off + 1]
    off += 1
    if kind != b"d":
        raise ValueError("only DOUBLEP dumps are supported (the hot path is FP64)")
    nx, ny, nz = struct.unpack_from("<3i", raw, off); off += 12
    dx, dy, dz = struct.unpack_from("<3d", raw, off); off += 24
    x0, y0, z0 = struct.unpack_from("<3i", raw, off); off += 12
    mx, my, mz = struct.unpack_from("<3i", raw, off); off += 12
    neq, neqdyn = struct.unpack_from("<2i", raw, off); off += 8
    (nghost,) = struct.unpack_from("<i", raw, off); off += 4
    rsc, vsc, rhosc = struct.unpack_from("<3d", raw, off); off += 24
    (cv,) = struct.unpack_from("<d", raw, off); off += 8
    shape = (neq, nx + 2 * nghost, ny + 2 * nghost, nz + 2 * nghost)
    u = np.frombuffer(raw, dtype="<f8", count=int(np.prod(shape)), offset=off).reshape(shape, order="F").copy(order="F")
    hdr = dict(n=(nx, ny, nz), d=(dx, dy, dz), origin=(x0, y0, z0), mpi=(mx, my, mz), neq=neq, neqdyn=neqdyn, nghost=nghost,
               scal=(rsc, vsc, rhosc), cv=cv)
    return u, hdr


def write_all_blocks(outputpath: str, blocks, itprint: int, **scal):
    """One file per block, like the reference's take-turns loop (Out_BIN_Module.f90:149-171).
    `blocks` is an iterable of guacho_b200.solver.Block (state is downloaded with get_state)."""
    paths = []
    for b in blocks:
        paths.append(write_bin(outputpath, b.get_state(), b.p, b.coords, b.rank, itprint, **scal))
    return paths
