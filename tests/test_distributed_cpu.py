"""Host-side logic of the N>1 path on CPU: two `gloo` ranks bootstrap like the GPU run does
(torch.distributed as the MPI_Init/MPI_Bcast stand-in), agree on the decomposition, cut their
blocks out of the global initial condition and reproduce the oracle's block layout.  The
device-side halo exchange itself (NCCL inside libguacho_gx.so) is covered by tests/test_multigpu.py."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["GX_ROOT"])
import numpy as np
import torch.distributed as dist
from guacho_b200.config import Params
from guacho_b200.decomp import choose_decomposition, coords_of, neighbors, halo_bytes_per_step
from guacho_b200.distributed import init_process_group, broadcast_bytes
from tests.util import global_ic, block_ic
from tests.oracle_lib import Oracle, U

rank, local_rank, world = init_process_group("gloo")
assert world == 2 and dist.get_backend() == "gloo"
payload = bytes(range(128)) if rank == 0 else None
got = broadcast_bytes(payload, 128, 0)                       # the NCCL unique id travels this way
assert got == bytes(range(128))
p = Params(nxtot=16, nytot=12, nztot=64, zmax=1.0)
nb = choose_decomposition(p, world)
assert nb == (1, 1, 2)                                       # 32-plane z slabs
pb = p.replace(MPI_NBX=nb[0], MPI_NBY=nb[1], MPI_NBZ=nb[2])
c = coords_of(rank, nb)
assert c == (0, 0, rank)
nbrs = neighbors(pb, c)
assert nbrs[4] == 1 - rank and nbrs[5] == 1 - rank and nbrs[0] == rank    # periodic: the other slab on both z sides, self in x/y
g = global_ic(p, "random")
mine = block_ic(pb, g, c)
assert mine.shape == pb.block_shape()
# the oracle scatters the same global array the same way (block `rank` of the emulated MPI job)
o = Oracle(pb); o.scatter_u(g)
assert np.array_equal(o.get_block(rank, U), mine)
gathered = [None, None]
dist.all_gather_object(gathered, (c, float(np.abs(mine).sum())))
t = __import__("torch").tensor([float(rank + 1)], dtype=__import__("torch").float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)                     # bench.py's max-over-ranks timing reduction
assert float(t.item()) == 2.0
assert halo_bytes_per_step(pb) == 8 * (2 * 8 * 2 * 20 * 16 + 2 * 8 * 18 * 14 + 2 * 2 * 3 * 18 * 14)
if rank == 0:
    assert [x[0] for x in gathered] == [(0, 0, 0), (0, 0, 1)]
    print("GLOO-OK")
dist.destroy_process_group()
'''


def test_two_rank_gloo_bootstrap_and_decomposition(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, GX_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "GLOO-OK" in r.stdout


def test_decomposition_choices():
    from guacho_b200.config import Params
    from guacho_b200.decomp import choose_decomposition, pencil_decomposition, slab_decomposition, coords_of, rank_of
    p = Params(nxtot=512, nytot=512, nztot=512, zmax=1.0)
    assert choose_decomposition(p, 8) == (1, 1, 8)                       # 64-plane slabs
    assert choose_decomposition(p, 8, min_thickness=128) == (1, 2, 4)    # pencils when slabs get thin
    assert pencil_decomposition(4) == (1, 2, 2) and slab_decomposition(4) == (1, 1, 4)
    for nb in ((1, 1, 8), (1, 2, 4), (2, 2, 2)):
        for r in range(8):
            assert rank_of(coords_of(r, nb), nb) == r
