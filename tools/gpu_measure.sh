#!/bin/bash
# measurement session: official-style bench (c2), reference arm, other workloads, ncu launch list + full-scale capture
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c2.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['setup_s'])"
timeout 300 python bench.py --impl reference --steps 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?"
for w in c1 c4 c5; do
  timeout 900 python bench.py --workload $w --steps 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))"
done
timeout 900 python bench.py --workload c3 --scale 0.03 --steps 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench c3 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c3.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['config']['reads_per_shard'])"; tail -3 gpurun_out/bench_c3.err
B="python bench.py --steps 2 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pileup|call_kernel|index_kernel|scan_kernel|compact|gather' -c 60 --csv --log-file gpurun_out/launches_c2_full.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'pileup_kernel|call_kernel' -s 6 -c 2 -f -o gpurun_out/prof_c2_full $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
