mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2/pytest_gpu2.txt 2>&1; tail -25 gpurun_out/r2/pytest_gpu2.txt
