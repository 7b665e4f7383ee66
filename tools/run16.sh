timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras 2>&1 | grep -o '"value": [0-9.e+]*, "unit\|"kernel_ms_per_step": {[^}]*}\|rror.*' | head -4
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
