#!/bin/bash
# round 2: prefetch off + call kernel at 3 CTAs per SM: parity, timing, traffic; the default bench line and the reference arm
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/r2q_pytest.log | cut -c1-300
B="python bench.py --steps 1 --no-e2e --no-e2e-h2d --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pileup_kernel|mate_kernel|call_kernel' -s 9 -c 3 -f -o gpurun_out/r2q_prof_c2 $B --samples 200 > gpurun_out/r2q_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cov_scan_kernel' -s 2 -c 1 -f -o gpurun_out/r2q_prof_cov python bench.py --workload cov --steps 1 --cov-samples 2 > gpurun_out/r2q_ncu_cov.log 2>&1
echo "ncu cov rc=$?"
S=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2q_bench_default.json 2> gpurun_out/r2q_bench_default.err
echo "bench default rc=$? in $(( $(date +%s) - S )) s"; python -c "import json;d=json.load(open('gpurun_out/r2q_bench_default.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['seconds'], d['e2e_h2d'].get('value'), d['cpu_baseline']['value'])"
S=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/r2q_bench_ref.json 2> gpurun_out/r2q_bench_ref.err
echo "bench ref rc=$? in $(( $(date +%s) - S )) s"; cut -c1-700 gpurun_out/r2q_bench_ref.json
