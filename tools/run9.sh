mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_exo_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "exo or source or callback" 2>&1 | tail -5
timeout 600 python bench.py --problem exo --steps 5 --warmup 3 > gpurun_out/r2/bench_exo_unfused.json 2> gpurun_out/r2/bench_exo_unfused.err; tail -c 1800 gpurun_out/r2/bench_exo_unfused.json; tail -3 gpurun_out/r2/bench_exo_unfused.err
