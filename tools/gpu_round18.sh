#!/bin/bash
# final validation of the round: parity suite, default bench, reference arm, smoke, BAM end-to-end comparisons, c3 alone
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_default.json'));print(d['kernels_ms'], d['value'], d['roofline'], d['e2e'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])"; tail -2 gpurun_out/bench_default.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
timeout 900 python bench.py --workload c3 --scale 0.03 --steps 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench c3 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c3.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e'])"
timeout 900 python tools/e2e_compare.py --preset c1 --scale 0.5 > gpurun_out/e2e_c1.json 2> gpurun_out/e2e_c1.err; echo "e2e c1 rc=$?"; cut -c1-1800 gpurun_out/e2e_c1.json
timeout 900 python tools/e2e_compare.py --preset c2 --scale 0.02 --samples 400 --work /tmp/msnv_e2e2 > gpurun_out/e2e_c2.json 2> gpurun_out/e2e_c2.err; echo "e2e c2 rc=$?"; cut -c1-1800 gpurun_out/e2e_c2.json
