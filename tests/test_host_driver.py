"""The compiled host driver (guacho_b200/host/guacho_host.cpp), the mirror of src/main.f90 over the C ABI."""
import os
import subprocess

import numpy as np
import pytest

from guacho_b200.build import build_host
from guacho_b200.bin_io import read_bin
from guacho_b200.config import ot_shipped


def test_host_driver_builds_and_refuses_to_run_without_a_gpu(tmp_path):
    exe = build_host()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test below")
    r = subprocess.run([exe, "-n", "8", "8", "2", "-tmax", "0.01", "-o", str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_host_driver_reproduces_the_python_host_bitwise(tmp_path):
    """Same problem through the two hosts (C++ loop of main.f90 vs guacho_b200.solver.Simulation): same iteration
    count, same dumps bit for bit, and the dumps are in the reference's BIN format."""
    from guacho_b200.solver import Block, Simulation
    from tests.util import global_ic
    exe = build_host()
    n = 64
    r = subprocess.run([exe, "-n", str(n), str(n), "2", "-tmax", "0.004", "-dtprint", "0.002", "-o", str(tmp_path)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    iters = int(r.stdout.strip().splitlines()[-1].split("(")[1].split()[0])
    p = ot_shipped(nxtot=n, nytot=n, nztot=2, zmax=2.0 / n, MPI_NBX=1, tmax=0.004, dtprint=0.002)
    files = sorted(os.listdir(os.path.join(tmp_path, "BIN")))
    # dump 0 is the initial condition as the C++ host evaluated it (std::sin; numpy's SIMD sin may differ in the last
    # bit): check it against the Python IC to round-off, then start the Python host from exactly that state
    u0, h0 = read_bin(os.path.join(tmp_path, "BIN", files[0]))
    ic = global_ic(p, "ot")
    assert np.abs(u0[..., 2:-2, 2:-2, 2:-2] - ic[..., 2:-2, 2:-2, 2:-2]).max() <= 1e-15
    dumps = []
    with Block(p) as b:
        sim = Simulation(b)
        sim.on_output = lambda s: dumps.append(b.get_state())
        sim.initflow(u0)
        dumps.append(b.get_state())
        assert sim.run() == iters
    assert files == [f"points000.{k:03d}.bin" for k in range(len(dumps))] and len(dumps) >= 3
    for k, ref in enumerate(dumps):
        u, h = read_bin(os.path.join(tmp_path, "BIN", files[k]))
        assert h["n"] == (n, n, 2) and h["neq"] == 8 and h["nghost"] == 2 and h["mpi"] == (1, 1, 1)
        assert np.array_equal(u[..., 2:-2, 2:-2, 2:-2], ref[..., 2:-2, 2:-2, 2:-2]), k


def test_python_runner_parses_and_refuses_without_a_gpu(tmp_path):
    """guacho_b200.run_ot is the Python twin of the compiled host; without a CUDA device it must fail loudly."""
    import torch
    from guacho_b200 import run_ot
    from guacho_b200.lib import GxError
    a = run_ot.parse(["--grid", "64", "64", "2", "--tmax", "0.01", "--out", str(tmp_path)])
    assert a.grid == [64, 64, 2] and a.tmax == 0.01 and a.blocks is None
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(GxError) as e:
        run_ot.main(["--grid", "16", "16", "2", "--tmax", "0.001", "--out", str(tmp_path) + "/", "--quiet"])
    assert e.value.code == -2
