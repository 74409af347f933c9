#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_q.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'])"
timeout 900 python bench.py --workload c3 --scale 0.03 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3q.json 2> gpurun_out/bench_c3q.err
echo "bench c3 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c3q.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'])"
timeout 900 python bench.py --workload c1 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c1q.json 2> gpurun_out/bench_c1q.err
echo "bench c1 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c1q.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'])"
