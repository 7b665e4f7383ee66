for lib in "$@"; do GUACHO_GX_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "ot_shipped or solvers_random or supersonic or fixture" 2>&1 | tail -3; done
for lib in guacho_b200/libguacho_gx.so "$@"; do
  echo "=== $lib"
  GUACHO_GX_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --no-extras 2>&1 | grep -o '"value": [0-9.e+]*, "unit\|"kernel_ms_per_step": {[^}]*}\|rror.*' | head -3
done
