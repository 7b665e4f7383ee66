mkdir -p gpurun_out/r2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 24 --csv --log-file gpurun_out/r2/launches_bench256.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/r2/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stage|k_bupdate" -s 8 -c 4 -f -o gpurun_out/r2/final_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/r2/ncu_full.log 2>&1
tail -2 gpurun_out/r2/ncu_full.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_n1_final.json 2> gpurun_out/r2/bench_n1_final.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2/bench_reference_final.json 2> gpurun_out/r2/bench_reference_final.err
python - <<'PY'
import json
for f in ('bench_n1_final','bench_reference_final'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r2/{f}.json') if l.startswith('{')][-1])
        print(f, d['value']/1e9, d['ms_per_step'], d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('kernel_ms_per_step'), d['e2e']['value'], d.get('extra'), d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
