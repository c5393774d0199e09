#!/bin/bash
# 8-GPU call: 2-rank NCCL test, bench weak at N = 1, 2, 4, 8 (20 steps, as the driver runs it) and strong at N = 8
TAG=${1:-r2z}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/${TAG}_bench1_k20.json 2> gpurun_out/${TAG}_bench1_k20.err
for N in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$N bench.py --gpus $N --steps 20 --warmup 3 --no-configs > gpurun_out/${TAG}_bench${N}_k20.json 2> gpurun_out/${TAG}_bench${N}_k20.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 --steps 20 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench8_strong.json 2> gpurun_out/${TAG}_bench8_strong.err
tail -3 gpurun_out/${TAG}_pytest.log
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'N', d['n_gpus'], d['scaling'], 'value %.0f' % d['value'], 'warm %.0f' % d['value_l2_warm_single_call'], 'e2e %.0f' % d['e2e'].get('value'), d['e2e'].get('job_ms_max_over_ranks'), d['e2e'].get('multi_gpu_check'), d['e2e'].get('error'))
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json', '.err')).read()[-1500:])
PY
