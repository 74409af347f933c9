#!/bin/bash
# quick GPU iteration: parity tests + reduced and full bench (no e2e, no CPU baseline)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -n 6 gpurun_out/pytest_gpu.log
B="python bench.py --scale 0.1 --steps 3 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/bench_s01.json 2> gpurun_out/bench_s01.err
echo "scale 0.1: $(python -c "import json;d=json.load(open('gpurun_out/bench_s01.json'));print(d['kernels_ms'], d['value'])")"
timeout 900 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_full_noe2e.json 2> gpurun_out/bench_full.err
echo "bench full rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_full_noe2e.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pileup_kernel -s 3 -c 1 -f -o gpurun_out/prof_pileup $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
