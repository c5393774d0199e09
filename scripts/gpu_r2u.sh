#!/bin/bash
# round 2: parity suite (render path on internal streams), render timing, raster occupancy A/B, hypothesis-part count sweep (warm call + flushed bench)
TAG=${1:-r2u}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for i in 1 2; do timeout 120 python scripts/dev_render.py 2>&1 | tail -1; done
ROUNDS=2 bash scripts/gpu_ab3.sh ${TAG} > /dev/null 2>&1
grep -v "^\.\|warn\|Docs\|^$\|tests/\|assert\|Consider\|/tmp" gpurun_out/${TAG}_kernels.log
for p in 1 2 3 4 5 6; do
DDOPE_PARTS=$p ITERS=200 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" | sed "s/^/parts=$p /"
DDOPE_PARTS=$p timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('parts=$p flushed value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'warm', round(d['value_l2_warm_single_call']), 'e2e', round(d['e2e']['value']))"
done
