mkdir -p gpurun_out/r2b
timeout 600 python -m pytest tests/test_thermal_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b/pytest_thermal2.txt; cat gpurun_out/r2b/pytest_thermal2.txt
timeout 300 python bench.py --problem tcond --steps 10 --warmup 3 > gpurun_out/r2b/bench_tcond.json 2> gpurun_out/r2b/bench_tcond.err; tail -c 1100 gpurun_out/r2b/bench_tcond.json; tail -3 gpurun_out/r2b/bench_tcond.err
