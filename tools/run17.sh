# 2 CTAs per SM for the first-order stage (tile 32x5, 6 warps each) against the default (1 CTA, 32x11, 12 warps)
for lib in guacho_b200/libguacho_gx.so guacho_b200/libgx_2cta.so; do
  echo "=== $lib"
  GUACHO_GX_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras 2>&1 | grep -o '"value": [0-9.e+]*, "unit\|"kernel_ms_per_step": {[^}]*}\|rror.*' | head -4
done
GUACHO_GX_LIB=$PWD/guacho_b200/libgx_2cta.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "ot_shipped or solvers_random" 2>&1 | tail -3
