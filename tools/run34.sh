mkdir -p gpurun_out/final2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 24 --csv --log-file gpurun_out/final2/launches_bench256.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/final2/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stage|k_bupdate" -s 8 -c 4 -f -o gpurun_out/final2/final_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/final2/ncu_full.log 2>&1
tail -2 gpurun_out/final2/ncu_full.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final2/bench_n1.json 2> gpurun_out/final2/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/final2/bench_n1.json') if l.startswith('{')][-1])
print(round(d['value']/1e9,4), round(d['ms_per_step'],3), d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['extra']['grid512']['value'])
PY
