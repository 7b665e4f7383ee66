#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_host_driver.py tests/test_abi.py -q 2>&1 | tail -2
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
  GUACHO_GX_LIB=$PWD/$lib python tests/_gpu_quick.py 2>&1 | grep -E "strict|fast" | cut -c1-150
  GUACHO_GX_LIB=$PWD/$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | python -c "$fmt"
done
