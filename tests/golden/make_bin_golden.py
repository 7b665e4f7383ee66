#!/usr/bin/env python
"""Fixture for the BIN dump format, generated WITH THE REFERENCE'S OWN READER (py/guacho_utils.py, imported
from /root/reference in the build container; it cannot travel to the GPU box, hence the fixture).

A 2x1x2-block random MHD state is written by guacho_b200.bin_io.write_bin, read back block by block and as a
whole domain through the reference's read_header / readbin3d_block / readbin3d_all, and the arrays the
reference returned are committed together with one of the files (bytes).  tests/test_bin_io.py then checks
that (a) the writer still produces those bytes and (b) our reader returns what the reference's reader returned.
    python tests/golden/make_bin_golden.py
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/py")

import guacho_utils as ref           # noqa: E402  (the reference's reader)
from guacho_b200.bin_io import write_bin     # noqa: E402
from guacho_b200.config import Params        # noqa: E402
from guacho_b200.decomp import coords_of     # noqa: E402
from tests.util import global_ic, block_ic   # noqa: E402


def main():
    p = Params(nxtot=8, nytot=6, nztot=4, zmax=1.0, MPI_NBX=2, MPI_NBY=1, MPI_NBZ=2)
    g = global_ic(p, "random")
    nb = (2, 1, 2)
    with tempfile.TemporaryDirectory() as d:
        out = d + "/"
        files = []
        for r in range(4):
            c = coords_of(r, nb)
            files.append(write_bin(out, block_ic(p, g, c), p, c, r, 7, rsc=3.0e12, vsc=1.2e6, rhosc=1.66e-24))
        with contextlib.redirect_stdout(io.StringIO()):
            hdr = ref.read_header(files[1], verbose=False)
            hdr[0].close()
            blk = ref.readbin3d_block(files[1], 6, conserved=True)                   # By of block 1
            rho_all = ref.readbin3d_all(7, 0, path=out + "BIN/", conserved=True)      # whole-domain density (transposed by the reader)
            pres_all = ref.readbin3d_all(7, 4, path=out + "BIN/", mhd=True, scale=False)   # the reader's own u2prim -> thermal pressure
        raw = open(files[1], "rb").read()
    np.savez_compressed(os.path.join(HERE, "reference_reader_bin.npz"), u_global=g, file1=np.frombuffer(raw, dtype=np.uint8),
                        hdr_n=np.array(hdr[2]), hdr_d=np.array(hdr[3]), hdr_origin=np.array(hdr[4]), hdr_mpi=np.array(hdr[5]),
                        hdr_neq=hdr[6], hdr_neqdyn=hdr[7], hdr_nghost=hdr[8], hdr_scal=np.array(hdr[9]), hdr_cv=np.array(hdr[10]),
                        block1_by=blk, rho_all=rho_all, pres_all=pres_all)
    print("reference reader: header", hdr[2:9], "rho_all", rho_all.shape)


if __name__ == "__main__":
    main()
