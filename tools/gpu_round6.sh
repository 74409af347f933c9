#!/bin/bash
mkdir -p gpurun_out
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared metasnv_b200/csrc/gpu/msnv_gpu.cu"
$NV -DMSNV_CHUNK_READS=127 -DMSNV_CHUNK_SEGS=256 -DMSNV_PILEUP_MIN_CTAS=8 -o /tmp/lib_r127c8.so
$NV -DMSNV_CHUNK_READS=127 -DMSNV_CHUNK_SEGS=256 -DMSNV_PILEUP_MIN_CTAS=7 -o /tmp/lib_r127c7.so
$NV -DMSNV_CHUNK_READS=127 -DMSNV_CHUNK_SEGS=256 -DMSNV_PILEUP_MIN_CTAS=6 -o /tmp/lib_r127c6.so
$NV -DMSNV_PILEUP_MIN_CTAS=6 -o /tmp/lib_c6.so
B="python bench.py --scale 0.1 --steps 3 --no-e2e --no-cpu-baseline"
for v in default r127c8 r127c7 r127c6 c6 r127c8q1792 r127c8q1536; do
  unset MSNV_LIB MSNV_CHUNK_Q4
  case $v in r127c8q1792) export MSNV_LIB=/tmp/lib_r127c8.so MSNV_CHUNK_Q4=1792;; r127c8q1536) export MSNV_LIB=/tmp/lib_r127c8.so MSNV_CHUNK_Q4=1536;; default) ;; *) export MSNV_LIB=/tmp/lib_$v.so;; esac
  timeout 300 $B > gpurun_out/bench_s01_$v.json 2> gpurun_out/bench_s01_$v.err
  echo "variant $v: $(python -c "import json;d=json.load(open('gpurun_out/bench_s01_$v.json'));print(d['kernels_ms']['ms_pileup'], d['value'], d['hits_per_shard'])")"
done
