#!/bin/bash
# One gpurun session that validates and measures the whole path (what the round-end numbers in profiles/ come from):
#   /usr/local/graft/bin/gpurun --timeout 2300 -- "bash tools/gpu_validate.sh"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_default.json'));print(d['kernels_ms'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'])"; tail -2 gpurun_out/bench_default.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
for w in c1 c4; do
  timeout 900 python bench.py --workload $w --steps 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
timeout 900 python bench.py --workload c3 --scale 0.03 --steps 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench c3 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c3.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'])"
timeout 900 python bench.py --workload c5 --scale 0.15 --steps 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "bench c5 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c5.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'])"
timeout 900 python tools/e2e_compare.py --preset c1 --scale 0.5 > gpurun_out/e2e_c1.json 2> gpurun_out/e2e_c1.err; echo "e2e c1 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/e2e_c1.json')); print({k:d[k] for k in ('gpu_pipe_first_process_s','gpu_pipe_s','cpu_pipe_s','identical','speedup_e2e_from_bam','qacompute_gpu_s','qacompute_ref_s')}); print(d['gpu_perf'])"
timeout 900 python tools/e2e_compare.py --preset c2 --scale 0.02 --samples 400 --work /tmp/msnv_e2e2 > gpurun_out/e2e_c2.json 2> gpurun_out/e2e_c2.err; echo "e2e c2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/e2e_c2.json')); print({k:d[k] for k in ('gpu_pipe_first_process_s','gpu_pipe_s','cpu_pipe_s','identical','speedup_e2e_from_bam')}); print(d['gpu_perf'])"
B="python bench.py --steps 2 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pileup|call_kernel|index_kernel|mark_kernel|scan_kernel|compact|gather' -c 60 --csv --log-file gpurun_out/launches_c2_full.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'pileup_kernel' -s 3 -c 1 -f -o gpurun_out/prof_c2_full_final $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
