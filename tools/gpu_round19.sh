#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exotic or hand or golden" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_quick.log
timeout 900 python tools/e2e_compare.py --preset c1 --scale 0.5 > gpurun_out/e2e_c1.json 2> gpurun_out/e2e_c1.err; echo "e2e c1 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/e2e_c1.json')); print({k:d[k] for k in ('gpu_pipe_first_process_s','gpu_pipe_s','cpu_pipe_s','identical','speedup_e2e_from_bam','qacompute_gpu_s','qacompute_ref_s')}); print(d['gpu_perf'])"
timeout 900 python tools/e2e_compare.py --preset c2 --scale 0.02 --samples 400 --work /tmp/msnv_e2e2 > gpurun_out/e2e_c2.json 2> gpurun_out/e2e_c2.err; echo "e2e c2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/e2e_c2.json')); print({k:d[k] for k in ('gpu_pipe_first_process_s','gpu_pipe_s','cpu_pipe_s','identical','speedup_e2e_from_bam')}); print(d['gpu_perf'])"
