#!/bin/bash
# full validation + measurement: parity suite, official bench (c2), other workloads, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c2.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['setup_s'], d['clocks'])"; tail -2 gpurun_out/bench_c2.err
for w in c1 c4; do
  timeout 900 python bench.py --workload $w --steps 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
timeout 900 python bench.py --workload c3 --scale 0.03 --steps 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench c3 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c3.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'])"
timeout 900 python bench.py --workload c5 --scale 0.15 --steps 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "bench c5 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c5.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'])"
B="python bench.py --steps 2 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pileup|call_kernel|index_kernel|mark_kernel|scan_kernel|compact|gather' -c 60 --csv --log-file gpurun_out/launches_c2_full.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'pileup_kernel|call_kernel' -s 6 -c 2 -f -o gpurun_out/prof_c2_full_v3 $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
