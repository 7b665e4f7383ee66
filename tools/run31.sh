for lib in guacho_b200/libguacho_gx.so guacho_b200/libgx_tcm4.so guacho_b200/libgx_tcm2.so; do
  echo "=== $lib"
  GUACHO_GX_LIB=$PWD/$lib timeout 300 python bench.py --problem tcond --steps 10 --warmup 3 2>&1 | grep -o '"frac": [0-9.e+-]*, "traffic\|"tcond": [0-9.]*\|rror.*' | head -3
done
