(timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -E "passed|failed|FAILED|Error|assert [0-9n]" | head -40) > gpurun_out/c15_tests.log 2>&1
cat gpurun_out/c15_tests.log
for m in -1 1; do
echo "adjoint mode $m"
(SKB_ADJ_MODE=$m timeout 300 python tools/time_bwd.py cfg4 cfg3b 2>&1 | tail -12) > gpurun_out/c15_time_bwd_$m.log 2>&1
cat gpurun_out/c15_time_bwd_$m.log
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c15_launches.csv python tools/recon_probe.py cfg4 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c15_launches.csv')) if len(r)>5]
hdr=rows[0]
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
for r in rows[-8:]:
    print(r[ik][:100], r[iv])
PY
