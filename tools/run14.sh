mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q -s 2>&1 | tail -12
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "viscosity" 2>&1 | tail -3
