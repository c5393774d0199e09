#!/bin/bash
TAG=${1:-r2g}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 120 python scripts/dev_render.py > gpurun_out/${TAG}_render.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_render_launches.csv env REPS=3 python scripts/dev_render.py > /dev/null 2>&1
tail -30 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_render.log; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2g_render_launches.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[-14:]: print(r[ki][:60], r[vi])
PY
