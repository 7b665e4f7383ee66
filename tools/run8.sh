mkdir -p gpurun_out/r2
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q -s > gpurun_out/r2/multigpu_tests_8gpu.log 2>&1; tail -25 gpurun_out/r2/multigpu_tests_8gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras > gpurun_out/r2/bench_n8.json 2> gpurun_out/r2/bench_n8.err; tail -c 1500 gpurun_out/r2/bench_n8.json; tail -3 gpurun_out/r2/bench_n8.err
