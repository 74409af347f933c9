#!/bin/bash
# round 2: evidence for profiles/ - launch list, ncu full sets (C2, C4, call, coverage scan), prefetch ablation, bench lines
mkdir -p gpurun_out
export MSNV_VERBOSE=1
B="python bench.py --steps 1 --no-e2e --no-e2e-h2d --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2p_launches_c2.csv python bench.py --steps 2 --no-e2e --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2p_launches_c2.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pileup_kernel|mate_kernel' -s 6 -c 2 -f -o gpurun_out/r2p_prof_c2 $B --samples 200 > gpurun_out/r2p_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pileup_kernel|mate_kernel' -s 6 -c 2 -f -o gpurun_out/r2p_prof_c4 $B --workload c4 --scale 0.1 > gpurun_out/r2p_ncu_c4.log 2>&1
echo "ncu c4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'call_kernel' -s 3 -c 1 -f -o gpurun_out/r2p_prof_call $B --samples 200 > gpurun_out/r2p_ncu_call.log 2>&1
echo "ncu call rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cov_scan_kernel' -s 8 -c 1 -f -o gpurun_out/r2p_prof_cov python bench.py --workload cov --steps 1 --cov-samples 2 > gpurun_out/r2p_ncu_cov.log 2>&1
echo "ncu cov rc=$?"
timeout 900 python tools/variant_sweep.py --settings ":::::,:::::64" > gpurun_out/r2p_sweep_c2.txt 2> gpurun_out/r2p_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2p_sweep_c2.txt
timeout 900 python tools/variant_sweep.py --preset c4 --settings ":::::,:::::64" > gpurun_out/r2p_sweep_c4.txt 2> gpurun_out/r2p_sweep_c4.err
echo "sweep c4 rc=$?"; cat gpurun_out/r2p_sweep_c4.txt
timeout 1200 python bench.py --steps 3 --e2e-bam-gb 5.0 --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2p_bench_c2_5g.json 2> gpurun_out/r2p_bench_c2_5g.err
echo "bench c2 5g rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2p_bench_c2_5g.json'));print(json.dumps(d.get('e2e'))[:2500])"
/usr/bin/time -v timeout 1200 python bench.py > gpurun_out/r2p_bench_default.json 2> gpurun_out/r2p_bench_default.err
echo "bench default rc=$?"; grep -E "Elapsed|Maximum resident" gpurun_out/r2p_bench_default.err; cut -c1-600 gpurun_out/r2p_bench_default.json
/usr/bin/time -v timeout 900 python bench.py --impl reference > gpurun_out/r2p_bench_ref.json 2> gpurun_out/r2p_bench_ref.err
echo "bench ref rc=$?"; grep -E "Elapsed" gpurun_out/r2p_bench_ref.err; cut -c1-900 gpurun_out/r2p_bench_ref.json
timeout 600 python bench.py --workload c4 --steps 3 --no-e2e --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2p_bench_c4.json 2> gpurun_out/r2p_bench_c4.err
echo "bench c4 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2p_bench_c4.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['roofline']['frac'])"
timeout 600 python bench.py --workload c1 --steps 5 --e2e-bam-gb 0.5 --no-cpu-baseline > gpurun_out/r2p_bench_c1.json 2> gpurun_out/r2p_bench_c1.err
echo "bench c1 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2p_bench_c1.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e_h2d']['value'])"
