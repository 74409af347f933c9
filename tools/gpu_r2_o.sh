#!/bin/bash
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/r2o_pytest.log | cut -c1-300
timeout 900 python tools/variant_sweep.py --settings ":::::" > gpurun_out/r2o_sweep_c2.txt 2> gpurun_out/r2o_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2o_sweep_c2.txt
timeout 900 python bench.py --steps 3 --e2e-bam-gb 1.0 --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2o_bench_c2_1g.json 2> gpurun_out/r2o_bench_c2_1g.err
echo "bench c2 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2o_bench_c2_1g.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d.get('e2e'))[:1200])"
timeout 1200 python bench.py --steps 3 --e2e-bam-gb 5.0 --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2o_bench_c2_5g.json 2> gpurun_out/r2o_bench_c2_5g.err
echo "bench c2 5g rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2o_bench_c2_5g.json'));print(json.dumps(d.get('e2e'))[:1400])"
