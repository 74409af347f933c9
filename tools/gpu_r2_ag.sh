#!/bin/bash
# round 2: snpCall start-up (first window decoded while the CUDA context comes up, staged afterwards): program-level tests, bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_metasnv_e2e.py tests/test_gpu_windows.py -x -q -m gpu --timeout 300 > gpurun_out/r2ag_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/r2ag_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2ag_bench_default.json 2> gpurun_out/r2ag_bench_default.err
echo "bench default rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2ag_bench_default.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['roofline']['frac'], d['roofline']['traffic']); e=d['e2e']; print(e['value'], e['seconds'], e['breakdown_s'], e['trace'])"
timeout 900 python bench.py --steps 3 --e2e-bam-gb 5.0 --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2ag_bench_c2_5g.json 2> gpurun_out/r2ag_bench_c2_5g.err
echo "bench c2 5g rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2ag_bench_c2_5g.json'));e=d['e2e']; print(e['value'], e['seconds'], e['breakdown_s'], e['trace'])"
