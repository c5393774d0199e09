#!/bin/bash
# Quick GPU pass: parity tests + per-kernel times (optionally of alternative builds in _lib/alt_*.so).
TAG=${1:-q}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
for i in 1 2; do
ITERS=50 TAG=default timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_kernels.log 2>&1
for alt in diff-dope_b200/diffdope/_lib/alt_*.so; do
  [ -f "$alt" ] && DDOPE_B200_LIB=$PWD/$alt ITERS=50 TAG=$(basename $alt) timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_kernels.log 2>&1
done
done
tail -25 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_kernels.log
