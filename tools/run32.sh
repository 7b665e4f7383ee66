mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/final/pytest_gpu.txt; cat gpurun_out/final/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err
timeout 600 python bench.py --problem exo --steps 10 --warmup 3 > gpurun_out/final/bench_exo.json 2> gpurun_out/final/bench_exo.err
timeout 600 python bench.py --problem tcond --steps 10 --warmup 3 > gpurun_out/final/bench_tcond.json 2> gpurun_out/final/bench_tcond.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_tc -c 8 --csv --log-file gpurun_out/final/tc_launches.csv python bench.py --problem tcond --steps 1 --warmup 3 > gpurun_out/final/tc_ncu.log 2>&1
python - <<'PY'
import json
for f in ('bench_n1','bench_exo','bench_tcond'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/final/{f}.json') if l.startswith('{')][-1])
        print(f, round(d['value']/1e9,4), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), (d.get('roofline',{}).get('whole_step') or {}).get('frac'), d.get('roofline',{}).get('kernel_ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('extra') or {}).get('grid512',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
