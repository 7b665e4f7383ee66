mkdir -p gpurun_out/r2
for lib in libguacho_gx.so libgx_bu82.so libgx_bu28.so libgx_bu42.so libgx_bu84.so libgx_bu81.so; do
  echo "=== $lib"
  GUACHO_GX_LIB=$PWD/guacho_b200/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | grep -o '"value": [0-9.e+]*, "unit\|"kernel_ms_per_step": {[^}]*}' | head -3
done > gpurun_out/r2/variants6.txt 2>&1
cat gpurun_out/r2/variants6.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
