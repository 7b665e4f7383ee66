#!/bin/bash
# 2-GPU call: multi-GPU parity tests, EXO parity, weak-scaling bench at N=1 and N=2
mkdir -p gpurun_out
nvidia-smi -L
( time python -m pytest tests/test_multigpu.py tests/test_exo_gpu.py -x -q ) 2>&1 | tail -15
python bench.py --steps 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_n1.json 2>gpurun_out/bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_n2.json 2>gpurun_out/bench_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        for ln in open(f"gpurun_out/bench_n{n}.json"):
            if ln.startswith("{"):
                d = json.loads(ln)
                print(n, "GPUs: value %.3f Gz/s  ms/step %.3f  e2e %.3f  launches %d  kernels %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["gpu_launches"], {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()}))
    except Exception as e:
        print(n, "failed", e)
PY
tail -5 gpurun_out/bench_n2.err
