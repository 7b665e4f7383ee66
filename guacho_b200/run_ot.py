"""Run the shipped Orszag-Tang problem like the reference does (`cd OT && make && mpirun -np N ./guacho`), with the
step on the GPU(s):

    python -m guacho_b200.run_ot [--grid NX NY NZ] [--tmax T] [--dtprint DT] [--out DIR] [--strict]
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 -m guacho_b200.run_ot --blocks 1 1 4 ...

One process per GPU (block); every rank writes its own `BIN/points<rank>.<it>.bin` in the reference's format
(src/Out_BIN_Module.f90), so `py/guacho_utils.py` and `OT/plots.py` of the reference read the output unchanged.
This is the Python twin of guacho_b200/host/guacho_host.cpp; the loop is src/main.f90:94-125.
"""
from __future__ import annotations

import argparse
import sys

from . import problems
from .bin_io import write_bin, write_divb
from .config import ot_shipped
from .decomp import coords_of
from .distributed import env_rank, init_process_group, make_rank_block
from .solver import Simulation


def parse(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--grid", type=int, nargs=3, default=[512, 512, 2], metavar=("NX", "NY", "NZ"))
    ap.add_argument("--blocks", type=int, nargs=3, default=None, metavar=("NBX", "NBY", "NBZ"),
                    help="block decomposition (MPI_NBX/NBY/NBZ); default: one block, or z slabs over the ranks")
    ap.add_argument("--tmax", type=float, default=0.5)
    ap.add_argument("--dtprint", type=float, default=0.1)
    ap.add_argument("--out", default="./")
    ap.add_argument("--strict", action="store_true", help="bit-comparison kernels (-fmad=false)")
    ap.add_argument("--divb", action="store_true", help="also dump div B (dump_divb in parameters.f90)")
    ap.add_argument("--warm", type=int, default=None, metavar="ITPRINT0", help="iwarm: restart from BIN dump number ITPRINT0 under --out")
    ap.add_argument("--quiet", action="store_true")
    return ap.parse_args(argv)


def main(argv=None) -> int:
    a = parse(argv)
    rank, local_rank, world = env_rank()
    nx, ny, nz = a.grid
    nb = tuple(a.blocks) if a.blocks else (1, 1, world)
    if nb[0] * nb[1] * nb[2] != world:
        raise SystemExit(f"--blocks {nb} needs {nb[0] * nb[1] * nb[2]} ranks, launched with {world}")
    p = ot_shipped(nxtot=nx, nytot=ny, nztot=nz, zmax=1.0 * nz / nx, MPI_NBX=1, tmax=a.tmax, dtprint=a.dtprint, strict_fp=a.strict)
    if world > 1:
        init_process_group()
    blk = make_rank_block(p, rank, world, local_rank, nb=nb)
    coords = coords_of(rank, nb)
    sim = Simulation(blk)

    def dump(s: Simulation) -> None:                       # write_output (src/output.f90:40-55)
        u = blk.get_state()
        path = write_bin(a.out, u, blk.p, coords, rank, s.itprint if s.time > 0 else 0)
        if a.divb:
            write_divb(a.out, u, blk.p, coords, rank, s.itprint if s.time > 0 else 0)
        if rank == 0 and not a.quiet:
            print(f"****************** wrote output *************** : {path}", flush=True)

    if a.warm is not None:                                 # iwarm (init.f90:134-142, 436-471): no output of the restart state
        sim.warm_start(a.out, a.warm)
    else:
        sim.initflow(problems.orszag_tang(blk.p, coords))  # initflow -> boundaryI -> calcprim (main.f90:73-79)
        dump(sim)                                          # main.f90:84-87: the initial condition is output 0 ...
        sim.itprint = 1                                    # ... and itprint moves on
    sim.on_output = dump
    while sim.time <= p.tmax:                              # main.f90:94
        dt = sim.step()
        if rank == 0 and not a.quiet:
            print(f"Iteration {sim.iteration - 1} | time:{sim.time - dt:12.3E} | dt:{dt:12.3E} | tprint:{sim.tprint:12.3E}", flush=True)
    if rank == 0:
        print("--- My work here is done, have a nice day ---", flush=True)
    blk.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
