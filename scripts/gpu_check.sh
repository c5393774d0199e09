#!/bin/bash
# round 2: parity suite + kernel times (bench workload, configs 4 / 5) + flushed bench line of the current build
TAG=${1:-r2w}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
CFGS=1 ROUNDS=2 bash scripts/gpu_ab3.sh ${TAG} > /dev/null 2>&1
grep -v "^\.\|warn\|Docs\|^$\|tests/\|assert\|Consider\|/tmp" gpurun_out/${TAG}_kernels.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('flushed value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'warm', round(d['value_l2_warm_single_call']), 'e2e', round(d['e2e']['value']), 'fwd', d['forward_only_ms_per_iter'], 'launches', d['gpu_launches'])"
