#!/bin/bash
# round 2: render path with concurrent fill (parity suite + timing), part-stream priorities A/B on the bench line
TAG=${1:-r2p}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
for i in 1 2 3; do timeout 120 python scripts/dev_render.py >> gpurun_out/${TAG}_render.log 2>&1; done
cat gpurun_out/${TAG}_render.log
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench_prio0_$i.json 2>> gpurun_out/${TAG}_bench.err
DDOPE_PART_PRIO=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench_prio1_$i.json 2>> gpurun_out/${TAG}_bench.err
done
for v in 0 1; do
DDOPE_PART_PRIO=$v ITERS=200 timeout 300 python scripts/dev_time.py 2>&1 | grep "ms/iter" | sed "s/^/prio$v /"
DDOPE_PART_PRIO=$v CFG=4 B=256 ITERS=50 timeout 300 python scripts/dev_configs.py 2>&1 | tail -1 | sed "s/^/prio$v cfg4 /"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p_bench_prio*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), d['ms_per_step'], round(d['value_l2_warm_single_call']), d['forward_only_ms_per_iter'], round(d['e2e']['value']))
P
