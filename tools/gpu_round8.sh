#!/bin/bash
mkdir -p gpurun_out
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared metasnv_b200/csrc/gpu/msnv_gpu.cu"
$NV -DMSNV_TILE=1024 -o /tmp/lib_t1024.so 2>&1 | grep -E "error" 
$NV -DMSNV_TILE=256 -o /tmp/lib_t256.so 2>&1 | grep -E "error"
B="python bench.py --scale 0.1 --steps 3 --no-e2e --no-cpu-baseline"
for v in default t1024s t1024l t256; do
  unset MSNV_LIB MSNV_CHUNK_Q4 MSNV_PILEUP_VARIANT
  case $v in t1024s) export MSNV_LIB=/tmp/lib_t1024.so MSNV_PILEUP_VARIANT=s;; t1024l) export MSNV_LIB=/tmp/lib_t1024.so MSNV_PILEUP_VARIANT=l;; t256) export MSNV_LIB=/tmp/lib_t256.so;; esac
  timeout 300 $B > gpurun_out/bench_s01_$v.json 2> gpurun_out/bench_s01_$v.err
  echo "variant $v: $(python -c "import json;d=json.load(open('gpurun_out/bench_s01_$v.json'));print(d['kernels_ms'], d['value'], d['hits_per_shard'])")"
  tail -2 gpurun_out/bench_s01_$v.err
done
