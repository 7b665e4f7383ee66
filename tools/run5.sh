mkdir -p gpurun_out/r2
for lib in libguacho_gx.so libgx_v1.so libgx_v2.so libgx_v3.so; do
  echo "=== $lib"
  GUACHO_GX_LIB=$PWD/guacho_b200/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | grep -o '"value": [0-9.e+]*, "unit\|"kernel_ms_per_step": {[^}]*}' | head -3
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
