#!/bin/bash
# round 2: parity suite, kernel times (bench workload + configs 4 / 5), e2e host sections, default bench line
TAG=${1:-r2o}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
CFGS=1 ROUNDS=2 bash scripts/gpu_ab3.sh ${TAG}
timeout 300 python scripts/dev_e2e_sections.py > gpurun_out/${TAG}_e2e_sections.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err; head -5 gpurun_out/${TAG}_e2e_sections.log
