mkdir -p gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 300 python bench.py --problem exo --steps 10 --warmup 3 > gpurun_out/r2f/bench_exo.json 2> gpurun_out/r2f/bench_exo.err
grep -o '"value": [0-9.e+]*, "unit\|"kernel_ms_per_step": {[^}]*}\|rror.*' gpurun_out/r2f/bench_exo.json | head -4
GX_NO_BC_SHELL=1 GX_NO_VISC_COOL=1 timeout 300 python bench.py --problem exo --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | grep -o '"value": [0-9.e+]*, "unit\|"kernel_ms_per_step": {[^}]*}\|rror.*' | head -3
