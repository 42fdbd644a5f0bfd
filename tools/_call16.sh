SKB_ADJ_MODE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c16_launches.csv python tools/time_bwd.py cfg4 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c16_launches.csv')) if len(r)>5]
hdr=rows[0]
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
tot=0
for r in rows[-34:]:
    print(r[ik][:90], r[iv]); tot+=float(r[iv])
print("sum us", tot/1000)
PY
