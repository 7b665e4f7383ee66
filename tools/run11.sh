mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_exo_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --problem exo --steps 5 --warmup 3 > gpurun_out/r2/bench_exo_fused.json 2> gpurun_out/r2/bench_exo_fused.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2/bench_exo_fused.json') if l.startswith('{')][-1])
    print(d['value']/1e9, d['ms_per_step'], d['config']['path'], d['roofline']['kernel_ms_per_step'], d['roofline']['whole_step']['frac'], d['e2e']['ms_per_step'], d['cpu_baseline'])
except Exception as e: print('ERR', e)
PY
tail -3 gpurun_out/r2/bench_exo_fused.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_n1_default.json 2> gpurun_out/r2/bench_n1_default.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2/bench_n1_default.json') if l.startswith('{')][-1])
    print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'], d['extra'], d['cpu_baseline'])
except Exception as e: print('ERR', e)
PY
tail -3 gpurun_out/r2/bench_n1_default.err
