#!/bin/bash
# Quick GPU pass: parity tests + per-kernel times + whole-call time with and without programmatic dependent launch.
TAG=${1:-q}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
ITERS=50 TAG=default timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_kernels.log 2>&1
for i in 1 2; do
echo "pdl:" >> gpurun_out/${TAG}_kernels.log; ITERS=200 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" >> gpurun_out/${TAG}_kernels.log
echo "no pdl:" >> gpurun_out/${TAG}_kernels.log; DDOPE_NO_PDL=1 ITERS=200 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" >> gpurun_out/${TAG}_kernels.log
done
echo "config 1 (B=1, 320 window, 50 iters) pdl / no pdl:" >> gpurun_out/${TAG}_kernels.log
B=1 WIN=320 ITERS=50 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" >> gpurun_out/${TAG}_kernels.log
DDOPE_NO_PDL=1 B=1 WIN=320 ITERS=50 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" >> gpurun_out/${TAG}_kernels.log
tail -25 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_kernels.log
for c in "4 128" "5 128" "3 128"; do set -- $c; CFG=$1 B=$2 ITERS=20 timeout 300 python scripts/dev_configs.py 2>&1 | tail -3 >> gpurun_out/${TAG}_configs.log; done
NO_EDGE=1 CFG=5 B=128 ITERS=20 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 >> gpurun_out/${TAG}_configs.log
cat gpurun_out/${TAG}_configs.log
