mkdir -p gpurun_out/final5
timeout 600 ncu --set full --clock-control none -k regex:"k_tc_march|k_tc_prim" -s 4 -c 3 -f -o gpurun_out/final5/tcond_full python bench.py --problem tcond --steps 1 --warmup 3 > gpurun_out/final5/ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_coolingh|k_viscous2|k_bc_shell|k_calcprim|k_wind" -s 8 -c 8 -f -o gpurun_out/final5/exo_ops_full python bench.py --problem exo --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/final5/ncu_exo.log 2>&1
ls -la gpurun_out/final5
