#!/bin/bash
# GPU call: tests, baseline bench, tuning variants, 512^3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
python bench.py --steps 30 > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err; tail -c 3000 gpurun_out/bench_base.json
fmt='
import sys, json
for ln in sys.stdin:
    if ln.startswith("{"):
        d = json.loads(ln); r = d["roofline"]
        print("value %.3f Gz/s  ms/step %.3f  kernels %s" % (d["value"]/1e9, d["ms_per_step"], {k: round(v,3) for k,v in r["kernel_ms_per_step"].items() if v}))
    else: print(ln.rstrip())
'
for lib in guacho_b200/libguacho_gx.so guacho_b200/variants/*.so; do
  echo "=== $lib"
  GUACHO_GX_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | python -c "$fmt"
done
echo "=== 512^3 baseline"
python bench.py --grid 512 --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_512.json 2>&1; python -c "$fmt" < gpurun_out/bench_512.json
