#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), ncu launch list + full capture of the two hot kernels.
TAG=${1:-r01b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-raw > gpurun_out/${TAG}_bench_e2e_raw.json 2> gpurun_out/${TAG}_bench_e2e_raw.err  # A/B of the e2e upload format
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
DDOPE_PARTS=1 ITERS=6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|pixel_kernel|iter_kernel" -s 35 -c 3 -f -o gpurun_out/${TAG}_full python scripts/dev_kernels.py > gpurun_out/${TAG}_full.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json; python -c "import json,sys; d=json.load(open(sys.argv[1])); print('e2e raw:', d['e2e'])" gpurun_out/${TAG}_bench_e2e_raw.json; cat gpurun_out/${TAG}_bench_ref.json
