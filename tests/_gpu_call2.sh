#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
fmt='
import sys, json
for ln in sys.stdin:
    if ln.startswith("{"):
        d = json.loads(ln); r = d["roofline"]
        print("value %.3f Gz/s  ms/step %.3f  kernels %s" % (d["value"]/1e9, d["ms_per_step"], {k: round(v,3) for k,v in r["kernel_ms_per_step"].items() if v}))
    else: print(ln.rstrip())
'
for lib in guacho_b200/libguacho_gx.so $EXTRA_LIBS; do
  echo "=== $lib"
  GUACHO_GX_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | python -c "$fmt"
done
