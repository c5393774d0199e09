#!/bin/bash
# round 2, call a: parity suite with the new full-size tests, mask-gradient dump, e2e sections
TAG=${1:-r2a}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q ) > gpurun_out/${TAG}_pytest_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_full.log
timeout 300 python scripts/dev_maskgrad_dump.py > gpurun_out/${TAG}_maskgrad.log 2>&1
timeout 300 python scripts/dev_e2e_sections.py > gpurun_out/${TAG}_e2e_sections.log 2>&1
tail -15 gpurun_out/${TAG}_pytest.log; tail -40 gpurun_out/${TAG}_pytest_full.log; cat gpurun_out/parity_measured.jsonl; tail -5 gpurun_out/${TAG}_maskgrad.log; head -5 gpurun_out/${TAG}_e2e_sections.log
