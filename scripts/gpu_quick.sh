#!/bin/bash
# Quick GPU pass: parity tests + per-kernel times.
TAG=${1:-q}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
ITERS=50 timeout 300 python scripts/dev_kernels.py > gpurun_out/${TAG}_kernels.log 2>&1
tail -25 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_kernels.log
