"""bench.py pieces that run without a GPU: the reference arm (the CPU restatement timed on the host cores) and
the workload/block-count logic."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--gpus", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "zone-updates/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_sample_block_count_divides_the_grid(monkeypatch):
    import bench
    seen = {}

    class FakeOracle:
        def __init__(self, p, fast=False, threads=1):
            p.validate()
            seen["blocks"], seen["threads"] = p.MPI_NBX, threads
            self.iter = 1
        def scatter_u(self, g): pass
        def start(self): pass
        def run_timed(self, n): return 1.0
    import tests.oracle_lib as ol
    monkeypatch.setattr(ol, "Oracle", FakeOracle)
    for cores, want in ((16, 16), (24, 16), (7, 4), (1, 1), (96, 64)):
        bench.cpu_reference(bench.workload(256, 1), 1, 1, cores)
        assert seen["blocks"] == want and seen["threads"] == want


def test_workloads():
    import bench
    p = bench.workload(256, 4)
    assert (p.nxtot, p.nytot, p.nztot, p.zmax) == (256, 256, 1024, 4.0) and p.enable_flux_cd and p.neq == 8
    p = bench.workload(512, 4, strong=True)
    assert (p.nxtot, p.nytot, p.nztot) == (512, 512, 512)
    p = bench.workload(384, 1, solver="hllc")
    assert p.neq == 5 and not p.mhd and not p.enable_flux_cd
    p = bench.workload(384, 1, solver="hlle")
    assert p.neq == 8 and p.enable_flux_cd
