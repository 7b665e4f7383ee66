#!/bin/bash
# tuning helper: bench one library under several values of an env var: _bench_env.sh LIB VAR v1 v2 ...
lib=$1; var=$2; shift 2
for v in "$@"; do
  echo "=== $var=$v"
  env $var=$v GUACHO_GX_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); r = d['roofline']
        print('value %.3f Gz/s  ms/step %.3f  frac %.3f  kernels %s' % (d['value']/1e9, d['ms_per_step'], r['frac'], {k: round(v,3) for k,v in r['kernel_ms_per_step'].items() if v}))
    else: print(ln.rstrip())
"
done
