"""BIN dump format (src/Out_BIN_Module.f90) — no GPU needed.

tests/golden/reference_reader_bin.npz was produced by reading our files with the REFERENCE's reader
(py/guacho_utils.py; see tests/golden/make_bin_golden.py), so these tests pin the writer to bytes the reference
accepts and our reader to the arrays the reference returns."""
import os

import numpy as np

from guacho_b200.bin_io import write_bin, read_bin, write_divb, divergence_b, header_bytes
from guacho_b200.config import Params
from guacho_b200.decomp import coords_of
from tests.util import block_ic

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_reader_bin.npz")
SCAL = dict(rsc=3.0e12, vsc=1.2e6, rhosc=1.66e-24)


def _setup():
    g = np.load(GOLD)
    p = Params(nxtot=8, nytot=6, nztot=4, zmax=1.0, MPI_NBX=2, MPI_NBY=1, MPI_NBZ=2)
    return g, p, np.asfortranarray(g["u_global"])


def test_writer_reproduces_the_bytes_the_reference_reader_accepted(tmp_path):
    g, p, u = _setup()
    c = coords_of(1, (2, 1, 2))
    path = write_bin(str(tmp_path) + "/", block_ic(p, u, c), p, c, 1, 7, **SCAL)
    assert path.endswith("BIN/points001.007.bin")                       # Out_BIN_Module.f90:129-130
    assert open(path, "rb").read() == g["file1"].tobytes()
    head = header_bytes(p, c, p.neq, 2, **SCAL)
    assert head.startswith(b"**************** Output for Guacho v1.3****************\n")
    assert b"Spacings      :  1.250E-01 1.667E-01 2.500E-01\n" in head   # es10.3
    assert b"\xff\nd" in head


def test_reader_returns_what_the_reference_reader_returned(tmp_path):
    g, p, u = _setup()
    paths = []
    for r in range(4):
        c = coords_of(r, (2, 1, 2))
        paths.append(write_bin(str(tmp_path) + "/", block_ic(p, u, c), p, c, r, 7, **SCAL))
    ub, h = read_bin(paths[1])
    assert h["n"] == tuple(g["hdr_n"]) and h["origin"] == tuple(g["hdr_origin"]) and h["mpi"] == tuple(g["hdr_mpi"])
    assert np.array_equal(h["d"], g["hdr_d"]) and np.array_equal(h["scal"], g["hdr_scal"]) and h["cv"] == g["hdr_cv"][0]
    assert (h["neq"], h["neqdyn"], h["nghost"]) == (int(g["hdr_neq"]), int(g["hdr_neqdyn"]), int(g["hdr_nghost"]))
    assert np.array_equal(ub, block_ic(p, u, coords_of(1, (2, 1, 2))))              # warm start: u with ghosts, bitwise
    assert np.array_equal(ub[6, 2:-2, 2:-2, 2:-2], g["block1_by"])                   # == readbin3d_block(conserved=True)
    # whole domain assembled from the per-block headers (x0,y0,z0), as readbin3d_all does (it returns map3d.T)
    rho = np.zeros((8, 6, 4))
    pres = np.zeros((8, 6, 4))
    for path in paths:
        a, hh = read_bin(path)
        x0, y0, z0 = hh["origin"]; nx, ny, nz = hh["n"]
        w = a[:, 2:-2, 2:-2, 2:-2]
        rho[x0:x0 + nx, y0:y0 + ny, z0:z0 + nz] = w[0]
        v = w[1:4] / w[0]
        pth = (w[4] - 0.5 * w[0] * (v[0] ** 2 + v[1] ** 2 + v[2] ** 2)) / hh["cv"]
        pres[x0:x0 + nx, y0:y0 + ny, z0:z0 + nz] = pth - 0.5 * (w[5] ** 2 + w[6] ** 2 + w[7] ** 2) / hh["cv"]
    assert np.array_equal(rho.T, g["rho_all"])
    assert np.allclose(pres.T, g["pres_all"], rtol=1e-14, atol=0)                    # guacho_utils.u2prim, equation 4


def test_divb_dump(tmp_path):
    g, p, u = _setup()
    c = coords_of(0, (2, 1, 2))
    ub = block_ic(p, u, c)
    path = write_divb(str(tmp_path) + "/", ub, p, c, 0, 3, **SCAL)
    raw = open(path, "rb").read()
    d = np.frombuffer(raw[-8 * 4 * 6 * 2:], dtype="<f8").reshape((4, 6, 2), order="F")
    ref = np.zeros((4, 6, 2))
    for i in range(4):
        for j in range(6):
            for k in range(2):     # Out_BIN_Module.f90:217-219 with Fortran indices i+1.. shifted by the 2 ghosts
                I, J, K = i + 2, j + 2, k + 2
                ref[i, j, k] = ((ub[5, I + 1, J, K] - ub[5, I - 1, J, K]) / (2. * p.dx) + (ub[6, I, J + 1, K] - ub[6, I, J - 1, K]) / (2. * p.dy)
                                + (ub[7, I, J, K + 1] - ub[7, I, J, K - 1]) / (2. * p.dz))
    assert np.array_equal(d, ref) and np.array_equal(divergence_b(ub, p), ref)
