#!/bin/bash
# round 2, first light of the persistent pileup kernel: parity tests, then timings and a staging sweep
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 15 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
echo "bench c2 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2a_bench_c2.json'));print(d['kernels_ms'], d['value'], d['ms_per_step'], d['roofline']['frac'])"; grep "msnv:" gpurun_out/r2a_bench_c2.err | tail -1
timeout 600 python tools/variant_sweep.py --settings "::,3::,4:3200:,4:3456:,2::,5:2560:160" > gpurun_out/r2a_sweep_c2.txt 2> gpurun_out/r2a_sweep_c2.err
echo "sweep rc=$?"; cat gpurun_out/r2a_sweep_c2.txt; grep "msnv:" gpurun_out/r2a_sweep_c2.err | sort | uniq -c
timeout 600 python bench.py --workload c4 --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err
echo "bench c4 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2a_bench_c4.json'));print(d['kernels_ms'], d['value'], d['ms_per_step'], d['roofline']['frac'])"; grep "msnv:" gpurun_out/r2a_bench_c4.err | tail -1
