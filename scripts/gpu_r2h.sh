#!/bin/bash
TAG=${1:-r2h}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 120 python scripts/dev_render.py > gpurun_out/${TAG}_render.log 2>&1
ITERS=50 TAG=default timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_render.log 2>&1
ITERS=200 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" >> gpurun_out/${TAG}_render.log
tail -30 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_render.log
