#!/bin/bash
# final pass of the round: parity suite, smoke, ncu evidence of the shipped build (full capture + counters + launch list), default bench line + reference arm
TAG=${1:-r2final}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SKIP_BENCH= bash scripts/gpu_evidence.sh ${TAG}
