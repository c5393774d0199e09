#!/bin/bash
# round 2, call f: one-launch multi-object, render fill stream, EDGE occupancy A/B
TAG=${1:-r2f}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
for i in 1 2; do
B=16 ITERS=100 timeout 300 python scripts/dev_multiobj.py >> gpurun_out/${TAG}_multi.log 2>&1
B=16 ITERS=100 NO_EDGE=1 timeout 300 python scripts/dev_multiobj.py >> gpurun_out/${TAG}_multi.log 2>&1
B=4 ITERS=100 NO_EDGE=1 timeout 300 python scripts/dev_multiobj.py >> gpurun_out/${TAG}_multi.log 2>&1
B=128 ITERS=50 timeout 300 python scripts/dev_multiobj.py >> gpurun_out/${TAG}_multi.log 2>&1
CFG=5 B=128 ITERS=10 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/edge3 cfg5 /" >> gpurun_out/${TAG}_kernels.log
CFG=3 B=128 ITERS=20 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/edge3 cfg3 obj3 /" >> gpurun_out/${TAG}_kernels.log
DDOPE_B200_LIB=$PWD/diff-dope_b200/diffdope/_lib/alt_edge4.so CFG=5 B=128 ITERS=10 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/edge4 cfg5 /" >> gpurun_out/${TAG}_kernels.log
DDOPE_B200_LIB=$PWD/diff-dope_b200/diffdope/_lib/alt_edge4.so CFG=3 B=128 ITERS=20 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/edge4 cfg3 obj3 /" >> gpurun_out/${TAG}_kernels.log
done
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -30 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_multi.log gpurun_out/${TAG}_kernels.log; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); print('fwd', d['forward_only_ms_per_iter'], 'value', d['value'], 'hot', d['value_l2_warm_single_call'], 'e2e', d['e2e']['value'])"
tail -3 gpurun_out/${TAG}_bench.err
