#!/bin/bash
# round 2: parity suite on the default build, then A/B (bench workload + configs 4 / 5) of the default build and every alt_*.so
TAG=${1:-r2m}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
CFGS=1 ROUNDS=2 bash scripts/gpu_ab3.sh ${TAG}
