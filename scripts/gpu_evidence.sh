#!/bin/bash
# round 2: ncu evidence of the shipped build (full capture of the three per-iteration kernels with source, launch list of the bench command),
# the counters file bench.py quotes, and the complete default bench line + reference arm
TAG=${1:-r2l}
mkdir -p gpurun_out
DDOPE_PARTS=1 ITERS=6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|pixel_kernel|iter_kernel" -s 35 -c 3 -f -o gpurun_out/${TAG}_full python scripts/dev_kernels.py > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
python scripts/ncu_counters.py gpurun_out/${TAG}_raw.csv gpurun_out/${TAG}_counters.json > /dev/null 2>> gpurun_out/${TAG}_full.log
cp gpurun_out/${TAG}_counters.json profiles/r02_counters.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_launches.log 2>&1
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
tail -c 3000 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
fi
tail -3 gpurun_out/${TAG}_full.log
