# final round-2 evidence: full GPU test suite, launch list, ncu --set full of the four step kernels, bench lines (N = 1, reference arm, EXO, solver sweep)
mkdir -p gpurun_out/evidence
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/evidence/pytest_gpu.txt; cat gpurun_out/evidence/pytest_gpu.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 24 --csv --log-file gpurun_out/evidence/launches_bench256.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/evidence/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stage|k_bupdate" -s 8 -c 4 -f -o gpurun_out/evidence/final_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/evidence/ncu_full.log 2>&1
tail -2 gpurun_out/evidence/ncu_full.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/evidence/bench_n1_final.json 2> gpurun_out/evidence/bench_n1_final.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/evidence/bench_reference_final.json 2> gpurun_out/evidence/bench_reference_final.err
timeout 600 python bench.py --problem exo --steps 10 --warmup 3 > gpurun_out/evidence/bench_exo.json 2> gpurun_out/evidence/bench_exo.err
for s in hll hllc hlle hlld; do timeout 300 python bench.py --solver $s --grid 384 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/evidence/bench_sweep384_$s.json 2> gpurun_out/evidence/bench_sweep384_$s.err; done
python - <<'PY'
import json
for f in ('bench_n1_final','bench_reference_final','bench_exo','bench_sweep384_hll','bench_sweep384_hllc','bench_sweep384_hlle','bench_sweep384_hlld'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/evidence/{f}.json') if l.startswith('{')][-1])
        print(f, round(d['value']/1e9,4), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), (d.get('roofline',{}).get('whole_step') or {}).get('frac'), d.get('roofline',{}).get('kernel_ms_per_step'), (d.get('e2e') or {}).get('value'), d.get('extra'), d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
