# 2-GPU weak-scaling line (overlapped peer push) with the final library + the multi-GPU tests
mkdir -p gpurun_out/evidence2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/evidence2/bench_n2.json 2> gpurun_out/evidence2/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/evidence2/bench_ref_n2.json 2> gpurun_out/evidence2/bench_ref_n2.err
python - <<'PY'
import json
for f in ('bench_n2','bench_ref_n2'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/evidence2/{f}.json') if l.startswith('{')][-1])
        print(f, round(d['value']/1e9,4), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('kernel_ms_per_step'), (d.get('e2e') or {}).get('value'), d.get('extra'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/evidence2/bench_n2.err
