#!/bin/bash
# quick A/B of every alt_*.so against the default build on the bench workload: per-kernel times + one warm call (ROUNDS rounds); CFGS=1 adds configs 4 / 5
TAG=${1:-ab}
mkdir -p gpurun_out
for i in $(seq 1 ${ROUNDS:-2}); do
for lib in default $(ls diff-dope_b200/diffdope/_lib/alt_*.so 2>/dev/null); do
  if [ $lib = default ]; then n=default; unset DDOPE_B200_LIB; else n=$(basename $lib .so); export DDOPE_B200_LIB=$PWD/$lib; fi
  ITERS=50 TAG=$n timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_kernels.log 2>&1
  ITERS=200 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" | sed "s/^/$n /" >> gpurun_out/${TAG}_kernels.log
  if [ -n "$CFGS" ] && [ $i = 1 ]; then
    CFG=5 B=128 ITERS=10 NO_EDGE=1 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/$n cfg5 noedge /" >> gpurun_out/${TAG}_kernels.log
    CFG=5 B=128 ITERS=10 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/$n cfg5 edge /" >> gpurun_out/${TAG}_kernels.log
    CFG=4 B=256 ITERS=50 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/$n cfg4 /" >> gpurun_out/${TAG}_kernels.log
  fi
done
done
unset DDOPE_B200_LIB
if [ -n "$TEST_ALT" ]; then ( DDOPE_B200_LIB=$PWD/diff-dope_b200/diffdope/_lib/$TEST_ALT timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log; fi
cat gpurun_out/${TAG}_kernels.log
