#!/bin/bash
# round 2: mate pass with the verdict-only rule; snpCall decodes its first window while the CUDA context comes up: full suite, bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r2af_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/r2af_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2af_bench_default.json 2> gpurun_out/r2af_bench_default.err
echo "bench default rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2af_bench_default.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['roofline']['frac'], d['roofline']['traffic']); e=d['e2e']; print(e['value'], e['seconds'], e['breakdown_s'], e['trace'])"
timeout 900 python bench.py --steps 3 --e2e-bam-gb 5.0 --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2af_bench_c2_5g.json 2> gpurun_out/r2af_bench_c2_5g.err
echo "bench c2 5g rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2af_bench_c2_5g.json'));e=d['e2e']; print(e['value'], e['seconds'], e['breakdown_s'], e['trace'])"
