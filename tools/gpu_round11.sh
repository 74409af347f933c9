#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c2 --scale 0.1 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_s01.json 2> gpurun_out/bench_s01.err
echo "bench s01 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_s01.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_s01.err
