mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.txt 2>&1; tail -5 gpurun_out/r2/pytest_gpu.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2/bench_v2.txt 2>&1
grep -o '"value": [0-9.e+]*\|"kernel_ms_per_step": {[^}]*}' gpurun_out/r2/bench_v2.txt; tail -3 gpurun_out/r2/bench_v2.txt | cut -c1-300
