mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q -s 2>&1 | tail -11
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r2/bench_n2.json 2> gpurun_out/r2/bench_n2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2/bench_n1b.json 2> gpurun_out/r2/bench_n1b.err
python - <<'PY'
import json
for f in ('bench_n2','bench_n1b'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r2/{f}.json') if l.startswith('{')][-1])
        print(f, d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['launches_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'])
    except Exception as e: print('ERR', e)
PY
