for lib in guacho_b200/libguacho_gx.so "$@"; do
  echo "=== $lib"
  GUACHO_GX_LIB=$PWD/$lib timeout 300 python bench.py --problem exo --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | grep -o '"value": [0-9.e+]*, "unit\|"kernel_ms_per_step": {[^}]*}\|rror.*' | head -3
done
