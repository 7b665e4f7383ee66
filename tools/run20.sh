mkdir -p gpurun_out/r2b
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_tc -c 12 --csv --log-file gpurun_out/r2b/tc_launches.csv python bench.py --problem tcond --steps 1 --warmup 3 > gpurun_out/r2b/tc_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r2b/tc_launches.csv') if l.startswith('"')))
h = rows[0]; ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
idc = h.index("ID")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[idc], r[ki][:40]), {})[r[mi]] = r[vi]
for k, v in list(d.items())[:12]:
    print(k, {a.split('.')[0][-28:]: b for a, b in v.items()})
PY
