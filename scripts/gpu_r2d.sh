#!/bin/bash
# round 2, call d: binned raster A/B
TAG=${1:-r2d}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
for i in 1 2; do
for mode in zbuffer binned; do
DDOPE_RASTER=$mode ITERS=50 TAG=$mode timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_kernels.log 2>&1
DDOPE_RASTER=$mode ITERS=200 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" | sed "s/^/$mode /" >> gpurun_out/${TAG}_kernels.log
DDOPE_RASTER=$mode DDOPE_PARTS=1 ITERS=200 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" | sed "s/^/$mode parts=1 /" >> gpurun_out/${TAG}_kernels.log
DDOPE_RASTER=$mode B=1 WIN=320 ITERS=50 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" | sed "s/^/$mode B=1 /" >> gpurun_out/${TAG}_kernels.log
DDOPE_RASTER=$mode CFG=5 B=128 ITERS=10 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/$mode cfg5 /" >> gpurun_out/${TAG}_kernels.log
DDOPE_RASTER=$mode CFG=5 B=128 ITERS=10 NO_EDGE=1 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/$mode cfg5 noedge /" >> gpurun_out/${TAG}_kernels.log
DDOPE_RASTER=$mode CFG=4 B=256 ITERS=50 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/$mode cfg4 /" >> gpurun_out/${TAG}_kernels.log
done
done
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -30 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_kernels.log; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); print('fwd', d['forward_only_ms_per_iter'], 'value', d['value'], 'hot', d['value_l2_warm_single_call'], 'e2e', d['e2e'])"
