mkdir -p gpurun_out/r2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 4 -c 2 -f -o gpurun_out/r2/v2_full python bench.py --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2/ncu_v2.log 2>&1
tail -3 gpurun_out/r2/ncu_v2.log | cut -c1-200
ls -la gpurun_out/r2/
