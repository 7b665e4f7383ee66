mkdir -p gpurun_out/r2f
timeout 600 python -m pytest tests/test_thermal_gpu.py -m gpu -x -q 2>&1 | tail -12
timeout 300 python bench.py --problem tcond --steps 10 --warmup 3 > gpurun_out/r2f/bench_tcond.json 2> gpurun_out/r2f/bench_tcond.err
grep -o '"value": [0-9.e+]*, "unit\|"frac": [0-9.e+-]*, "traffic\|"tcond": [0-9.]*\|"substeps_per_step": [0-9.]*\|rror.*' gpurun_out/r2f/bench_tcond.json | head -5; tail -3 gpurun_out/r2f/bench_tcond.err
