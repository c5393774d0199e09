#!/bin/bash
# 2-GPU call: NCCL hardware test + bench at N=2 (weak and strong) with 20 and 200 steps
TAG=${1:-r2i}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
for K in 20 200; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps $K --warmup 3 --no-configs > gpurun_out/${TAG}_bench2_k$K.json 2> gpurun_out/${TAG}_bench2_k$K.err
timeout 600 python bench.py --gpus 1 --steps $K --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/${TAG}_bench1_k$K.json 2> gpurun_out/${TAG}_bench1_k$K.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench2_cfgs.json 2> gpurun_out/${TAG}_bench2_cfgs.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29715 bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench2_strong.json 2> gpurun_out/${TAG}_bench2_strong.err
tail -8 gpurun_out/${TAG}_pytest.log
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/' + "${TAG}" + '_bench*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'N', d['n_gpus'], 'K', d['steps'], 'value %.0f' % d['value'], 'hot %.0f' % d['value_l2_warm_single_call'], 'e2e', d['e2e'].get('value'), d['e2e'].get('ms_per_step'), d['e2e'].get('multi_gpu_check'), d['e2e'].get('error'))
        if d.get('configs'):
            for c in d['configs']:
                if c.get('config') == 4: print('   cfg4', c['hypotheses_per_gpu'], c['value'], c['ms_per_iter'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json', '.err')).read()[-1500:])
PY
