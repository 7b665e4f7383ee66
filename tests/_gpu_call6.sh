#!/bin/bash
# 4-GPU call: pencil/slab parity tests, weak-scaling bench at N=4, strong-scaling 512^3 at N=4
mkdir -p gpurun_out
( time python -m pytest tests/test_multigpu.py -x -q ) 2>&1 | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_n4.json 2>gpurun_out/bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 4 --grid 512 --strong --steps 10 --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_n4_strong512.json 2>gpurun_out/bench_n4_strong512.err
python - <<'PY'
import json
for f in ("bench_n4", "bench_n4_strong512"):
    try:
        for ln in open(f"gpurun_out/{f}.json"):
            if ln.startswith("{"):
                d = json.loads(ln)
                print(f, ": value %.3f Gz/s  ms/step %.3f  scaling %s  launches %d  kernels %s" % (d["value"]/1e9, d["ms_per_step"], d["scaling"], d["gpu_launches"], {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()}))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/bench_n4.err gpurun_out/bench_n4_strong512.err
