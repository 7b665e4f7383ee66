#!/bin/bash
# tuning helper: run the GPU tests once, then the bench for each library variant given as args
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lib in "$@"; do
  echo "=== $lib"
  GUACHO_GX_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); r = d['roofline']
        print('value %.3f Gz/s  ms/step %.3f  frac %.3f  kernels %s' % (d['value']/1e9, d['ms_per_step'], r['frac'], {k: round(v,3) for k,v in r['kernel_ms_per_step'].items()}))
    else: print(ln.rstrip())
"
done
