#!/bin/bash
# round 2: everything new on the host side: new GPU tests, bench legs (from-BAM e2e, coverage pass, windowed c3 / c5)
mkdir -p gpurun_out
export MSNV_VERBOSE=1
nproc > gpurun_out/r2m_env.txt; free -g >> gpurun_out/r2m_env.txt; nvidia-smi -L >> gpurun_out/r2m_env.txt
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 12 gpurun_out/r2m_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 3 --e2e-bam-gb 1.0 --no-e2e-h2d > gpurun_out/r2m_bench_c2.json 2> gpurun_out/r2m_bench_c2.err
echo "bench c2 rc=$?"; tail -n 3 gpurun_out/r2m_bench_c2.err | cut -c1-300; python -c "import json;d=json.load(open('gpurun_out/r2m_bench_c2.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d.get('e2e'))[:1500])"
timeout 600 python bench.py --workload cov --steps 3 > gpurun_out/r2m_bench_cov.json 2> gpurun_out/r2m_bench_cov.err
echo "bench cov rc=$?"; tail -n 3 gpurun_out/r2m_bench_cov.err | cut -c1-300; cut -c1-1800 gpurun_out/r2m_bench_cov.json
timeout 900 python bench.py --workload c3 --scale 0.125 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/r2m_bench_c3.json 2> gpurun_out/r2m_bench_c3.err
echo "bench c3 rc=$?"; tail -n 3 gpurun_out/r2m_bench_c3.err | cut -c1-300; python -c "import json;d=json.load(open('gpurun_out/r2m_bench_c3.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['config']['windows_per_shard'], d['roofline']['frac'])"
timeout 900 python bench.py --workload c5 --scale 1.0 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/r2m_bench_c5.json 2> gpurun_out/r2m_bench_c5.err
echo "bench c5 rc=$?"; tail -n 3 gpurun_out/r2m_bench_c5.err | cut -c1-300; python -c "import json;d=json.load(open('gpurun_out/r2m_bench_c5.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['config']['windows_per_shard'], d['roofline']['frac'])"
