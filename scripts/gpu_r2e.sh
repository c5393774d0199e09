#!/bin/bash
# round 2, call e: edge single-shading, graph replay, render overlap
TAG=${1:-r2e}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
for i in 1 2; do
ITERS=50 TAG=default timeout 300 python scripts/dev_kernels.py >> gpurun_out/${TAG}_kernels.log 2>&1
B=1 WIN=320 ITERS=50 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" | sed "s/^/graph B=1 /" >> gpurun_out/${TAG}_kernels.log
DDOPE_GRAPH=0 B=1 WIN=320 ITERS=50 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" | sed "s/^/nograph B=1 /" >> gpurun_out/${TAG}_kernels.log
B=4 WIN=640 ITERS=50 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" | sed "s/^/graph B=4 /" >> gpurun_out/${TAG}_kernels.log
DDOPE_GRAPH=0 B=4 WIN=640 ITERS=50 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" | sed "s/^/nograph B=4 /" >> gpurun_out/${TAG}_kernels.log
CFG=5 B=128 ITERS=10 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/cfg5 /" >> gpurun_out/${TAG}_kernels.log
CFG=3 B=128 ITERS=20 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/cfg3 obj3 /" >> gpurun_out/${TAG}_kernels.log
done
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -30 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_kernels.log; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); print('fwd', d['forward_only_ms_per_iter'], 'value', d['value'], 'hot', d['value_l2_warm_single_call'], 'e2e', d['e2e']['value'])
for c in d['configs']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in c.items() if k!='workload'})"
tail -3 gpurun_out/${TAG}_bench.err
