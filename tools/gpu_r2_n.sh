#!/bin/bash
# round 2: byte-wide mismatch screen in the scatter loop, coverage scan with small CTAs, budgets of the windowed runs
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 12 gpurun_out/r2n_pytest.log | cut -c1-300
timeout 900 python tools/variant_sweep.py --settings ":::::,:::::1,3:::::" > gpurun_out/r2n_sweep_c2.txt 2> gpurun_out/r2n_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2n_sweep_c2.txt; grep "msnv:" gpurun_out/r2n_sweep_c2.err | sort | uniq -c | cut -c1-250
timeout 900 python tools/variant_sweep.py --preset c4 --settings ":::::,::::128:" > gpurun_out/r2n_sweep_c4.txt 2> gpurun_out/r2n_sweep_c4.err
echo "sweep c4 rc=$?"; cat gpurun_out/r2n_sweep_c4.txt; grep "msnv:" gpurun_out/r2n_sweep_c4.err | sort | uniq -c | cut -c1-250
timeout 900 python bench.py --steps 3 --e2e-bam-gb 1.0 --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2n_bench_c2.json 2> gpurun_out/r2n_bench_c2.err
echo "bench c2 rc=$?"; grep -v "^msnv: pileup" gpurun_out/r2n_bench_c2.err | tail -n 3 | cut -c1-300; python -c "import json;d=json.load(open('gpurun_out/r2n_bench_c2.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d.get('e2e'))[:1200])"
timeout 600 python bench.py --workload cov --steps 3 > gpurun_out/r2n_bench_cov.json 2> gpurun_out/r2n_bench_cov.err
echo "bench cov rc=$?"; tail -n 3 gpurun_out/r2n_bench_cov.err | cut -c1-300; python -c "import json;d=json.load(open('gpurun_out/r2n_bench_cov.json'));print(d['value'], d['kernels_ms'], d['roofline']['frac'], d['roofline']['whole_pass'], d['e2e'])"
timeout 900 python bench.py --workload c3 --scale 0.125 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/r2n_bench_c3.json 2> gpurun_out/r2n_bench_c3.err
echo "bench c3 rc=$?"; grep -v "^msnv: pileup" gpurun_out/r2n_bench_c3.err | tail -n 3 | cut -c1-300; python -c "import json;d=json.load(open('gpurun_out/r2n_bench_c3.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['config']['windows_per_shard'], d['roofline']['frac'])"
timeout 900 python bench.py --workload c5 --scale 1.0 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/r2n_bench_c5.json 2> gpurun_out/r2n_bench_c5.err
echo "bench c5 rc=$?"; grep -v "^msnv: pileup" gpurun_out/r2n_bench_c5.err | tail -n 3 | cut -c1-300; python -c "import json;d=json.load(open('gpurun_out/r2n_bench_c5.json'));print(d['value'], d['ms_per_step'], d['kernels_ms'], d['config']['windows_per_shard'], d['roofline']['frac'])"
