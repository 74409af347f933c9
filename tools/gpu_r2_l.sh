#!/bin/bash
# round 2: scatter kernel + coalesced mate pass: parity, timing at C2 / C4, ncu (full set + shared-atomic counters)
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 8 gpurun_out/r2l_pytest.log | cut -c1-300
timeout 900 python tools/variant_sweep.py --settings ":::::,4:3200::::" > gpurun_out/r2l_sweep_c2.txt 2> gpurun_out/r2l_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2l_sweep_c2.txt; grep "msnv:" gpurun_out/r2l_sweep_c2.err | sort | uniq -c | cut -c1-250
timeout 900 python tools/variant_sweep.py --preset c4 --settings ":::::" > gpurun_out/r2l_sweep_c4.txt 2> gpurun_out/r2l_sweep_c4.err
echo "sweep c4 rc=$?"; cat gpurun_out/r2l_sweep_c4.txt; grep "msnv:" gpurun_out/r2l_sweep_c4.err | sort | uniq -c | cut -c1-250
timeout 600 python bench.py --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/r2l_bench_c2.json 2> gpurun_out/r2l_bench_c2.err
echo "bench c2 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2l_bench_c2.json'));print(d['kernels_ms'], d['value'], d['ms_per_step'], d['roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pileup_kernel|mate_kernel' -s 6 -c 2 -f -o gpurun_out/r2l_prof_c2 python bench.py --steps 1 --no-e2e --no-cpu-baseline --samples 200 > gpurun_out/r2l_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
