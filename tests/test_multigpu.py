"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the decomposed run through
NCCL halo exchange must reproduce the single-block run bitwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(nb, grid, nsteps, problem, strict=False, port=29541, extra=()):
    world = nb[0] * nb[1] * nb[2]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), *map(str, nb), *map(str, grid), str(nsteps), problem]
    if strict:
        cmd.append("strict")
    cmd.extend(extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout
    for ln in r.stdout.splitlines():
        if ln.startswith("MGPU"):
            print(ln)                    # kept in the log of `pytest -s` (profiles/r2_multigpu_tests_8gpu.log)


@pytest.mark.parametrize("nb", [(1, 1, 2), (2, 1, 1), (1, 2, 1)])
def test_two_gpu_slabs_bitwise(nb):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(nb, (64, 48, 40), 3, "random")


def test_four_gpu_pencils_bitwise():
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    _run((1, 2, 2), (64, 48, 40), 3, "random", port=29543)


def test_two_gpu_z_slabs_with_physical_ends_bitwise():
    """Outflow (mirror) boundaries at the two ends of the z decomposition: each rank has one real neighbour and one
    physical face; the peer-memory push runs serialized (no overlap) next to the ghost-fill kernels."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run((1, 1, 2), (64, 48, 40), 3, "blast", port=29545, extra=("outflowz",))


def test_four_gpu_z_slabs_bitwise():
    """Four z slabs, periodic: distinct low/high neighbours for every rank — the decomposition bench.py uses for
    weak scaling (peer-memory push overlapped with the interior launches)."""
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    _run((1, 1, 4), (64, 48, 80), 3, "random", port=29546)


def test_eight_gpu_z_slabs_bitwise():
    """Eight z slabs, periodic — the decomposition of the driver's 1 -> 8 GPU scaling run (bench.py --gpus 8): peer-memory
    push over NVLink overlapped with the interior launches; the gathered state equals the single-block run bitwise."""
    if _ngpu() < 8:
        pytest.skip("needs 8 GPUs")
    _run((1, 1, 8), (64, 48, 192), 3, "random", port=29547)


@pytest.mark.parametrize("strict", [True, False])
def test_two_gpu_viscosity_matches_the_oracle_on_the_same_blocks(strict):
    """eta != 0 on two GPUs.  viscous_copy (src/hydro_solver.f90:54-63) reads up's ghost cells, which still hold the
    HALF-step halo exchanged by boundaryII (:169) while the interior holds full-step values (:188), so the
    reference's result depends on the block decomposition next to every internal face.  The contract is therefore
    equality with the reference run on the SAME block grid: the oracle with MPI_NBZ = 2."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run((1, 1, 2), (48, 40, 32), 3, "random", strict=strict, port=29548, extra=("eta", "oracle"))


def test_two_gpu_y_blocks_viscosity_matches_the_oracle_on_the_same_blocks():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run((1, 2, 1), (48, 40, 32), 3, "random", strict=True, port=29549, extra=("eta", "oracle"))


@pytest.mark.parametrize("nb,port", [((1, 1, 2), 29550), ((1, 2, 1), 29551)])
def test_two_gpu_thermal_conduction_matches_the_oracle_on_the_same_blocks(nb, port):
    """Thermal conduction (src/thermal_cond.f90) on two GPUs: the conduction time scale is a global minimum (mpi_allreduce,
    :104), every substep exchanges one layer of u(5) alone (thermal_bounds, :496-616) — z slabs through the peer-memory push
    of a single variable, y blocks through pack / NCCL / unpack."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(nb, (32, 24, 24), 3, "random", port=port, extra=("tcond", "oracle"))
