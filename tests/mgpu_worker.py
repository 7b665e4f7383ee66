"""Worker for the multi-GPU parity test (run under torch.distributed.run, one rank per GPU).

Every rank advances its block of a decomposed domain through the C ABI (NCCL halo exchange and
CFL all-reduce inside libguacho_gx.so); rank 0 also advances the same problem as ONE block and
checks that the gathered interiors are bitwise equal (G-GPU == 1-GPU).
usage: mgpu_worker.py NBX NBY NBZ NX NY NZ NSTEPS PROBLEM [strict] [outflowz] [eta] [oracle] [tcond]
  eta    : eta = 0.01 (viscous_copy reads up's stale half-step ghosts, SURVEY Q5: the reference itself depends on the decomposition)
  tcond  : isotropic saturated thermal conduction (src/thermal_cond.f90) with several super-time-stepping substeps per step, outflow walls
  oracle : compare with the CPU oracle run on the SAME block grid (1e-12 relative per variable; bitwise with `strict`) instead of one GPU block
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("MGPU_WATCHDOG_S", "90")), exit=True)   # a hang shows where each rank is stuck
    import torch
    import torch.distributed as dist
    from guacho_b200.config import Params
    from guacho_b200.decomp import coords_of
    from guacho_b200.distributed import init_process_group, make_rank_block
    from guacho_b200.solver import Block
    from guacho_b200 import problems

    nbx, nby, nbz, nx, ny, nz, nsteps = (int(a) for a in sys.argv[1:8])
    problem = sys.argv[8]
    strict = "strict" in sys.argv[9:]
    outflowz = "outflowz" in sys.argv[9:]           # physical (mirror) boundaries at the ends of the z decomposition
    with_eta = "eta" in sys.argv[9:]
    vs_oracle = "oracle" in sys.argv[9:]
    rank, local_rank, world = init_process_group("nccl")
    assert world == nbx * nby * nbz
    torch.cuda.set_device(local_rank)
    from guacho_b200.config import BC_OUTFLOW
    kw = dict(bc_out=BC_OUTFLOW, bc_in=BC_OUTFLOW) if outflowz else {}
    if with_eta:
        kw["eta"] = 0.01
    if "tcond" in sys.argv[9:]:
        from guacho_b200.config import TC_ISOTROPIC
        from tests.util import tc_scalings
        kw.update(th_cond=TC_ISOTROPIC, tc_saturation=True, bc_left=BC_OUTFLOW, bc_right=BC_OUTFLOW, bc_bottom=BC_OUTFLOW, bc_top=BC_OUTFLOW,
                  bc_out=BC_OUTFLOW, bc_in=BC_OUTFLOW, **tc_scalings(rhosc=1e-18))
    p = Params(nxtot=nx, nytot=ny, nztot=nz, zmax=1.0, strict_fp=strict, **kw)
    nb = (nbx, nby, nbz)
    blk = make_rank_block(p, rank, world, local_rank, nb=nb)
    coords = coords_of(rank, nb)
    # every rank cuts its block out of the SAME global array (evaluating the IC per block can differ
    # from the global evaluation by an ulp: numpy's SIMD and scalar-tail cos() are not bit-identical)
    from tests.util import global_ic, block_ic
    gic = global_ic(p, problem)
    blk.set_state(block_ic(blk.p, gic, coords))
    t, it = 0.0, 1
    dts = []
    for _ in range(nsteps):
        dt, _d = blk.get_timestep(it, 10, t, 1e300)
        blk.tstep(dt)
        dts.append(dt)
        t += dt
        it += 1
    if p.th_cond:
        assert blk.tc_info()[1] > 1, blk.tc_info()          # the super-time-stepping schedule ran
    mine = np.ascontiguousarray(blk.interior(blk.get_state()))
    gathered = [None] * world
    dist.all_gather_object(gathered, (coords, mine, dts))
    ok = True
    if rank == 0:
        full = np.zeros((blk.p.neq, nx, ny, nz))
        for c, a, d in gathered:
            assert d == dts, "ranks disagree on dt (the CFL all-reduce is an exact min)"
            bx, by, bz = blk.p.nx, blk.p.ny, blk.p.nz
            full[:, c[0] * bx:(c[0] + 1) * bx, c[1] * by:(c[1] + 1) * by, c[2] * bz:(c[2] + 1) * bz] = a
        if vs_oracle:
            from tests.oracle_lib import Oracle, U
            o = Oracle(p.replace(MPI_NBX=nbx, MPI_NBY=nby, MPI_NBZ=nbz), threads=min(world, 8))
            o.scatter_u(gic)
            o.start()
            t1, it1 = 0.0, 1
            for n in range(nsteps):
                dt, _d = o.get_timestep(it1, 10, t1, 1e300)
                assert abs(dt - dts[n]) <= 1e-13 * dt, (dt, dts[n])
                assert o.tstep(dts[n]) == 0
                t1 += dts[n]; it1 += 1
            ref = o.gather(U)
            err = max(np.abs(full[q] - ref[q]).max() / max(np.abs(ref[q]).max(), 1e-300) for q in range(full.shape[0]))
            ok = bool(err <= (0.0 if strict else 1e-12))
            print(f"MGPU nb={nb} grid={nx}x{ny}x{nz} steps={nsteps} problem={problem} strict={strict} eta={p.eta}: max rel |multi - oracle(same blocks)| = {err:.3e} -> {'OK' if ok else 'FAIL'}", flush=True)
        else:
            p1 = p.replace(device=local_rank)
            with Block(p1) as one:
                one.set_state(gic)
                t1, it1 = 0.0, 1
                for n in range(nsteps):
                    dt, _d = one.get_timestep(it1, 10, t1, 1e300)
                    if dt != dts[n]:
                        print(f"MGPU dt mismatch at step {n}: single {dt!r} multi {dts[n]!r}", flush=True)
                    one.tstep(dts[n])
                    t1 += dt
                    it1 += 1
                ref = one.interior(one.get_state())
            diff = np.abs(full - ref).max()
            ok = bool(diff == 0.0)
            print(f"MGPU nb={nb} grid={nx}x{ny}x{nz} steps={nsteps} problem={problem} strict={strict}: max|multi - single| = {diff:.3e} -> {'OK' if ok else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
