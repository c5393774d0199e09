#!/bin/bash
# A/B per-kernel times of the default build and every diff-dope_b200/diffdope/_lib/alt_*.so (+ parity tests of the default build)
TAG=${1:-ab}
mkdir -p gpurun_out
for i in 1 2 3; do
ITERS=50 TAG=default timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_kernels.log 2>&1
[ -n "$AB_ENV" ] && env $AB_ENV ITERS=50 TAG="$AB_ENV" timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_kernels.log 2>&1
for alt in diff-dope_b200/diffdope/_lib/alt_*.so; do
  [ -f "$alt" ] && DDOPE_B200_LIB=$PWD/$alt ITERS=50 TAG=$(basename $alt) timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_kernels.log 2>&1
done
done
[ -z "$SKIP_TESTS" ] && ( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_kernels.log
