#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -n 6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload c3 --scale 0.03 --steps 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench c3 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c3.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['config']['reads_per_shard'])"; tail -3 gpurun_out/bench_c3.err
timeout 900 python bench.py --workload c5 --scale 0.3 --steps 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "bench c5 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c5.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))"; tail -3 gpurun_out/bench_c5.err
timeout 900 python tools/e2e_compare.py --preset c1 --scale 0.5 > gpurun_out/e2e_c1.json 2> gpurun_out/e2e_c1.err
echo "e2e c1 rc=$?"; cat gpurun_out/e2e_c1.json | cut -c1-1800; tail -3 gpurun_out/e2e_c1.err
timeout 900 python tools/e2e_compare.py --preset c2 --scale 0.02 --samples 400 --work /tmp/msnv_e2e2 > gpurun_out/e2e_c2.json 2> gpurun_out/e2e_c2.err
echo "e2e c2 rc=$?"; cat gpurun_out/e2e_c2.json | cut -c1-1800; tail -3 gpurun_out/e2e_c2.err
