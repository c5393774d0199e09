#!/bin/bash
# Quick GPU pass: parity tests + per-kernel times + whole-call time (config 2 and config 1).
TAG=${1:-q}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
for i in 1 2; do
ITERS=50 TAG=default timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_kernels.log 2>&1
ITERS=200 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" >> gpurun_out/${TAG}_kernels.log
B=1 WIN=320 ITERS=50 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" >> gpurun_out/${TAG}_kernels.log
done
tail -25 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_kernels.log
