#!/bin/bash
TAG=${1:-r2c}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench20.json 2> gpurun_out/${TAG}_bench20.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs --e2e-float > gpurun_out/${TAG}_bench20f.json 2> gpurun_out/${TAG}_bench20f.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs --scaling strong > gpurun_out/${TAG}_bench_strong.json 2> gpurun_out/${TAG}_bench_strong.err
tail -25 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench20.json gpurun_out/${TAG}_bench20f.json gpurun_out/${TAG}_bench_strong.json; tail -3 gpurun_out/${TAG}_bench_strong.err
