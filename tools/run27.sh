mkdir -p gpurun_out/r2d
timeout 900 python -m pytest tests/test_exo_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --problem exo --steps 10 --warmup 3 > gpurun_out/r2d/bench_exo.json 2> gpurun_out/r2d/bench_exo.err
grep -o '"value": [0-9.e+]*, "unit\|"kernel_ms_per_step": {[^}]*}\|rror.*' gpurun_out/r2d/bench_exo.json | head -4
