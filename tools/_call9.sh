python tools/recon_probe.py cfg4 small 2>&1 | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c9_launches.csv python tools/recon_probe.py cfg4 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c9_launches.csv')) if len(r)>5]
hdr=rows[0]
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
for r in rows[1:]:
    print(r[ik][:90], r[iv])
PY
