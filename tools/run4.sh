mkdir -p gpurun_out/r2
for lib in libguacho_gx.so libgx_ty1_15.so; do
  echo "=== $lib"
  GUACHO_GX_LIB=$PWD/guacho_b200/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | grep -o '"value": [0-9.e+]*\|"kernel_ms_per_step": {[^}]*}' | head -3
done
GUACHO_GX_LIB=$PWD/guacho_b200/libgx_ty1_15.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "not limiters" 2>&1 | tail -3
