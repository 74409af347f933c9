#!/bin/bash
# GPU session: parity tests, full-shape bench, ncu launch list + full capture of the pileup kernel (reduced shape).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench full rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
B="python bench.py --scale 0.1 --steps 2 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pileup|call_kernel|index_kernel|scan_kernel|compact|gather' -c 60 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pileup_kernel -s 3 -c 1 -f -o gpurun_out/prof_pileup $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:index_kernel -s 6 -c 2 -f -o gpurun_out/prof_index $B > gpurun_out/ncu_index.log 2>&1
ls -la gpurun_out
