#!/bin/bash
# round 2: memory checker over the new kernels (expand, mate, coverage scan, windows), then the final bench lines
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_abi.py -x -q -m gpu -k "expand or cov_run" > gpurun_out/r2v_memcheck_abi.log 2>&1
echo "memcheck abi rc=$?"; tail -n 4 gpurun_out/r2v_memcheck_abi.log | cut -c1-200
for b in snpCall qaCompute; do mv metasnv_b200/bin/$b metasnv_b200/bin/$b.real; printf '#!/bin/bash\nexec compute-sanitizer --tool memcheck --error-exitcode 9 --log-file /tmp/san.$$.log "$(dirname "$0")/%s.real" "$@"\n' $b > metasnv_b200/bin/$b; chmod +x metasnv_b200/bin/$b; done
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hand_written or golden_fixture or (in_windows and c4)" > gpurun_out/r2v_memcheck_programs.log 2>&1
echo "memcheck programs rc=$?"; tail -n 4 gpurun_out/r2v_memcheck_programs.log | cut -c1-200; cat /tmp/san.*.log 2>/dev/null | grep -c "ERROR SUMMARY: 0 errors"; cat /tmp/san.*.log 2>/dev/null | grep "ERROR SUMMARY" | grep -v ": 0 errors" | head
for b in snpCall qaCompute; do mv metasnv_b200/bin/$b.real metasnv_b200/bin/$b; done
S=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2v_bench_default.json 2> gpurun_out/r2v_bench_default.err
echo "bench default rc=$? in $(( $(date +%s) - S )) s"; python -c "import json;d=json.load(open('gpurun_out/r2v_bench_default.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['e2e']['seconds'], d['e2e_h2d'].get('value'), d['cpu_baseline']['value'])"
timeout 900 python bench.py --workload c1 --steps 5 --e2e-bam-gb 0.5 > gpurun_out/r2v_bench_c1.json 2> gpurun_out/r2v_bench_c1.err
echo "bench c1 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2v_bench_c1.json'));print(d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d.get('e2e_metasnv'))"
