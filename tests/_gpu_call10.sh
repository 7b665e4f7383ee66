#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -k "not full_orszag" 2>&1 | tail -6 )
python bench.py --steps 20 --no-cpu-baseline --e2e-steps 0 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); r = d['roofline']
        print('value %.3f Gz/s  ms/step %.3f  kernels %s' % (d['value']/1e9, d['ms_per_step'], {k: round(v,3) for k,v in r['kernel_ms_per_step'].items() if v}))
    else: print(ln.rstrip())
"
