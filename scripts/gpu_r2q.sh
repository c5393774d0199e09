#!/bin/bash
# round 2: image output -- resident CTAs per SM of the background fill beside raster + pixel pass; per-kernel durations of one call
TAG=${1:-r2q}
mkdir -p gpurun_out
for c in 1 2 3 4 8; do for i in 1 2; do DDOPE_FILL_CTAS=$c timeout 120 python scripts/dev_render.py 2>&1 | sed "s/^/fill_ctas=$c /" >> gpurun_out/${TAG}_render.log; done; done
REPS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_render_launches.csv python scripts/dev_render.py > /dev/null 2>&1
cat gpurun_out/${TAG}_render.log
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r2q_render_launches.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]; h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
for r in rows[hi+1:][-12:]: print(r[kn][:50], r[mv])
P
