#!/bin/bash
# 8-GPU call: bench weak / strong at N=8 (+ N=1 on the same box for the ratio)
TAG=${1:-r2j}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/${TAG}_bench1_k20.json 2> gpurun_out/${TAG}_bench1_k20.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench8_k20.json 2> gpurun_out/${TAG}_bench8_k20.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus 8 --steps 200 --warmup 3 --no-configs > gpurun_out/${TAG}_bench8_k200.json 2> gpurun_out/${TAG}_bench8_k200.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29725 bench.py --gpus 8 --steps 20 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench8_strong.json 2> gpurun_out/${TAG}_bench8_strong.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29727 bench.py --gpus 4 --steps 20 --warmup 3 --no-configs > gpurun_out/${TAG}_bench4_k20.json 2> gpurun_out/${TAG}_bench4_k20.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2j_bench*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'N', d['n_gpus'], 'K', d['steps'], 'value %.0f' % d['value'], 'hot %.0f' % d['value_l2_warm_single_call'], 'e2e', d['e2e'].get('value'), d['e2e'].get('ms_per_step'), d['e2e'].get('multi_gpu_check'), d['e2e'].get('error'))
        if d.get('configs'):
            for c in d['configs']:
                print('   cfg', c.get('config'), c['hypotheses_per_gpu'], c.get('value'), c.get('value_per_gpu'), c['ms_per_iter'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json', '.err')).read()[-1500:])
PY
