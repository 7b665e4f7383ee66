mkdir -p gpurun_out/r2d
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2d/bench_n1.json 2> gpurun_out/r2d/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2d/bench_n1.json') if l.startswith('{')][-1])
print(round(d['value']/1e9,4), round(d['ms_per_step'],3), d['e2e'], d['roofline']['traffic'], d['roofline']['fp64_issue'])
PY
