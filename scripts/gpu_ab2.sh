#!/bin/bash
# A/B of every alt_*.so against the default build: per-kernel times + whole call, config 2 and config 5 (no edge)
TAG=${1:-ab}
mkdir -p gpurun_out
for i in 1 2 3; do
ITERS=50 TAG=default timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_kernels.log 2>&1
ITERS=200 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" | sed "s/^/default /" >> gpurun_out/${TAG}_kernels.log
for alt in diff-dope_b200/diffdope/_lib/alt_*.so; do
  n=$(basename $alt .so)
  DDOPE_B200_LIB=$PWD/$alt ITERS=50 TAG=$n timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_kernels.log 2>&1
  DDOPE_B200_LIB=$PWD/$alt ITERS=200 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" | sed "s/^/$n /" >> gpurun_out/${TAG}_kernels.log
done
done
CFG=5 B=128 ITERS=10 NO_EDGE=1 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/default cfg5 noedge /" >> gpurun_out/${TAG}_kernels.log
CFG=4 B=256 ITERS=50 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/default cfg4 /" >> gpurun_out/${TAG}_kernels.log
for alt in diff-dope_b200/diffdope/_lib/alt_*.so; do
  n=$(basename $alt .so)
  DDOPE_B200_LIB=$PWD/$alt CFG=5 B=128 ITERS=10 NO_EDGE=1 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/$n cfg5 noedge /" >> gpurun_out/${TAG}_kernels.log
  DDOPE_B200_LIB=$PWD/$alt CFG=4 B=256 ITERS=50 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/$n cfg4 /" >> gpurun_out/${TAG}_kernels.log
done
if [ -n "$TEST_ALT" ]; then ( DDOPE_B200_LIB=$PWD/diff-dope_b200/diffdope/_lib/$TEST_ALT timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log; fi
cat gpurun_out/${TAG}_kernels.log
