#!/bin/bash
mkdir -p gpurun_out
export MGPU_WATCHDOG_S=60
( time timeout 400 python -m pytest tests/test_multigpu.py -x -q ) 2>&1 | tail -15
for p2p in 0 1; do
  if [ $p2p = 0 ]; then export GX_NO_P2P=1; else unset GX_NO_P2P; fi
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$p2p bench.py --gpus 2 --steps 20 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_n2_p2p$p2p.json 2>gpurun_out/bench_n2_p2p$p2p.err
  python - <<PY
import json
for ln in open("gpurun_out/bench_n2_p2p$p2p.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("p2p=$p2p: value %.3f Gz/s  ms/step %.3f  e2e %.3f launches %d  kernels %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["gpu_launches"], {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()}))
PY
done
tail -n 5 gpurun_out/bench_n2_p2p1.err
