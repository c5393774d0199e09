#!/bin/bash
TAG=${1:-r2b}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
LOSSES=all timeout 300 python scripts/dev_maskgrad_dump.py > gpurun_out/${TAG}_maskgrad.log 2>&1
tail -30 gpurun_out/${TAG}_pytest.log; cat gpurun_out/parity_measured.jsonl; tail -3 gpurun_out/${TAG}_maskgrad.log
