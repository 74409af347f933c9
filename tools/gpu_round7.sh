#!/bin/bash
# parity (incl. metaSNV.py end to end), official-style bench runs, other workloads, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -n 6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c2.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e'], d['setup_s'])"
timeout 300 python bench.py --impl reference --steps 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-300
for w in c1 c4; do
  timeout 900 python bench.py --workload $w --steps 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))"
done
B="python bench.py --scale 0.1 --steps 2 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pileup|call_kernel|index_kernel|scan_kernel|compact|gather' -c 60 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pileup_kernel|call_kernel' -s 6 -c 2 -f -o gpurun_out/prof_pileup_call $B > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
