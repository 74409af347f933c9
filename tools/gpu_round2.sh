#!/bin/bash
# GPU session 2: full pytest -m gpu, bench flow at reduced and full C2 shape.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --scale 0.02 --samples 200 --steps 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
echo "bench small rc=$?"; cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench full rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
timeout 300 python bench.py --impl reference --steps 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
