# 8-GPU evidence with the final library: the 4- and 8-GPU bitwise tests and the driver's N = 8 bench line
mkdir -p gpurun_out/evidence8
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -s -k "four or eight" 2>&1 | grep -v "^$" | tail -12 > gpurun_out/evidence8/multigpu_8gpu.log; cat gpurun_out/evidence8/multigpu_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/evidence8/bench_n8.json 2> gpurun_out/evidence8/bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/evidence8/bench_n8.json') if l.startswith('{')][-1])
print(round(d['value']/1e9,3), round(d['ms_per_step'],3), d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'], {k:(round(v.get('value',0)/1e9,3), v.get('ms_per_step')) for k,v in d['extra'].items()})
PY
tail -2 gpurun_out/evidence8/bench_n8.err | cut -c1-300
