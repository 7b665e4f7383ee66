mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/gpu.txt
./tools/lat > gpurun_out/r2/lat.txt 2>&1
for lib in libguacho_gx.so libgx_ty9.so libgx_ty10.so; do
  echo "=== $lib" >> gpurun_out/r2/variants1.txt
  GUACHO_GX_LIB=$PWD/guacho_b200/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 >> gpurun_out/r2/variants1.txt 2>&1
done
cat gpurun_out/r2/lat.txt
grep -o '"value": [0-9.e+]*\|"kernel_ms_per_step": {[^}]*}\|===.*' gpurun_out/r2/variants1.txt
