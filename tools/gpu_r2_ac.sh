#!/bin/bash
# round 2: evidence for the gather kernel: full GPU suite, bench lines, launch list, DRAM traffic at full size, ncu full sets
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r2ac_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/r2ac_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2ac_bench_default.json 2> gpurun_out/r2ac_bench_default.err
echo "bench default rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2ac_bench_default.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['roofline']['frac'], d['roofline']['kernel'], d['e2e']['value'], d['e2e_h2d'].get('value'), d['cpu_baseline']['value'])"
for W in "c3 --scale 0.125" "c5 --scale 1.0"; do
  for K in gather scatter; do
    N=$(echo $W | cut -d' ' -f1)
    MSNV_PILEUP=$K timeout 600 python bench.py --workload $W --steps 2 --no-e2e --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2ac_bench_${N}_$K.json 2> gpurun_out/r2ac_bench_${N}_$K.err
    echo "bench $N $K rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2ac_bench_${N}_$K.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['config']['windows_per_shard'], d['roofline']['frac'])"
  done
done
timeout 600 python bench.py --workload c4 --steps 3 --no-e2e --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2ac_bench_c4.json 2> gpurun_out/r2ac_bench_c4.err
echo "bench c4 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2ac_bench_c4.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['roofline']['frac'], d['roofline']['kernel'])"
timeout 600 python bench.py --workload c1 --steps 5 --no-e2e --no-cpu-baseline > gpurun_out/r2ac_bench_c1.json 2> gpurun_out/r2ac_bench_c1.err
echo "bench c1 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2ac_bench_c1.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['roofline']['frac'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4001 -c 70 --csv --log-file gpurun_out/r2ac_launches_c2.csv python bench.py --steps 2 --no-e2e --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2ac_launches_c2.log 2>&1
echo "launch list rc=$?"; grep -c "pileup_gather_kernel" gpurun_out/r2ac_launches_c2.csv
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'pileup_.*kernel|mate_kernel|fix_clear_kernel|call_kernel' -s 12 -c 4 --csv --log-file gpurun_out/r2ac_traffic_c2.csv python bench.py --steps 1 --no-e2e --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2ac_traffic_c2.log 2>&1
echo "traffic c2 rc=$?"; grep -v "^==" gpurun_out/r2ac_traffic_c2.csv | cut -d, -f5,13- | tail -14
B="python bench.py --steps 1 --no-e2e --no-e2e-h2d --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pileup_gather_kernel|mate_kernel' -s 6 -c 2 -f -o gpurun_out/r2ac_prof_c2 $B --samples 200 > gpurun_out/r2ac_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"; tail -2 gpurun_out/r2ac_ncu_c2.log
