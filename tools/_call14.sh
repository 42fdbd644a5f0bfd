(timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -E "passed|failed|FAILED|Error|assert [0-9n]" | head -40) > gpurun_out/c14_tests.log 2>&1
cat gpurun_out/c14_tests.log
(timeout 300 python tools/time_bwd.py cfg4 cfg3b 2>&1 | tail -12) > gpurun_out/c14_time_bwd.log 2>&1
cat gpurun_out/c14_time_bwd.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c14_launches.csv python tools/time_bwd.py cfg4 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c14_launches.csv')) if len(r)>5]
hdr=rows[0]
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
out=[(r[ik][:100], r[iv]) for r in rows[1:]]
for k,v in out[-40:]:
    print(k, v)
PY
