#!/bin/bash
# round 2, final state: full GPU suite, the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r2aj_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/r2aj_pytest.log | cut -c1-400
timeout 600 python bench.py > gpurun_out/r2aj_bench_default.json 2> gpurun_out/r2aj_bench_default.err
echo "bench default rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2aj_bench_default.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['roofline']['frac'], d['roofline']['traffic']); e=d['e2e']; print(e['value'], e['seconds'], e['breakdown_s'])"
