#!/bin/bash
# round 2: records expanded on the device (expand_kernel): parity, and what it does to the host's decode time
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 8 gpurun_out/r2t_pytest.log | cut -c1-300
timeout 1200 python bench.py --steps 2 --e2e-bam-gb 5.0 --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2t_bench_c2_5g_raw.json 2> gpurun_out/r2t_bench_c2_5g_raw.err
echo "bench raw rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2t_bench_c2_5g_raw.json'));e=d['e2e'];print(e['value'], e['seconds'], e['breakdown_s'], e['h2d_bytes_per_step'])"
MSNV_RAW=0 timeout 1200 python bench.py --steps 2 --e2e-bam-gb 5.0 --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2t_bench_c2_5g_host.json 2> gpurun_out/r2t_bench_c2_5g_host.err
echo "bench host-aligned rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2t_bench_c2_5g_host.json'));e=d['e2e'];print(e['value'], e['seconds'], e['breakdown_s'], e['h2d_bytes_per_step'])"
timeout 900 python bench.py --workload c1 --steps 3 --e2e-bam-gb 1.0 --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2t_bench_c1_raw.json 2> gpurun_out/r2t_bench_c1_raw.err
echo "bench c1 raw rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2t_bench_c1_raw.json'));e=d['e2e'];print(e['value'], e['seconds'], e['breakdown_s'])"
MSNV_RAW=0 timeout 900 python bench.py --workload c1 --steps 3 --e2e-bam-gb 1.0 --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2t_bench_c1_host.json 2> gpurun_out/r2t_bench_c1_host.err
echo "bench c1 host rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2t_bench_c1_host.json'));e=d['e2e'];print(e['value'], e['seconds'], e['breakdown_s'])"
