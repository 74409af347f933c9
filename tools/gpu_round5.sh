#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -n 6 gpurun_out/pytest_gpu.log
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared metasnv_b200/csrc/gpu/msnv_gpu.cu"
$NV -DMSNV_PILEUP_MIN_CTAS=5 -o /tmp/lib_c5.so
$NV -DMSNV_PILEUP_MIN_CTAS=8 -o /tmp/lib_c8.so
$NV -DMSNV_PILEUP_MIN_CTAS=4 -o /tmp/lib_c4.so
B="python bench.py --scale 0.1 --steps 3 --no-e2e --no-cpu-baseline"
for v in default c5 c8 c4 q1536 q4096; do
  unset MSNV_LIB MSNV_CHUNK_Q4
  case $v in c5|c8|c4) export MSNV_LIB=/tmp/lib_$v.so;; q1536) export MSNV_CHUNK_Q4=1536;; q4096) export MSNV_CHUNK_Q4=4096;; esac
  timeout 300 $B > gpurun_out/bench_s01_$v.json 2> gpurun_out/bench_s01_$v.err
  echo "variant $v: $(python -c "import json;d=json.load(open('gpurun_out/bench_s01_$v.json'));print(d['kernels_ms'], d['value'])")"
done
unset MSNV_LIB MSNV_CHUNK_Q4
timeout 900 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_full_noe2e.json 2> gpurun_out/bench_full.err
echo "bench full rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_full_noe2e.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pileup_kernel -s 3 -c 1 -f -o gpurun_out/prof_pileup $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
