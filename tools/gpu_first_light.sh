#!/bin/bash
# First-light GPU session: environment facts, parity on small shapes, a memcheck pass.
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
  echo "nproc=$(nproc)"; free -g | head -2
} > gpurun_out/env.txt 2>&1
timeout 300 python tools/parity_run.py --preset c1 --scale 0.02 --samples 6 --split --cov --text --work /tmp/p1 > gpurun_out/p1.log 2>&1
echo "p1 rc=$?" >> gpurun_out/p1.log
# memcheck on the same tiny case (direct mode)
( cd /tmp/p1 && timeout 600 bash -c "/root/repo/metasnv_b200/bin/samtools mpileup -f ref.fa -B -b all_samples | compute-sanitizer --tool memcheck --error-exitcode 9 /root/repo/metasnv_b200/bin/snpCall -f ref.fa -i /tmp/p1/san.indiv > /tmp/p1/san.called" ) > gpurun_out/sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
timeout 600 python tools/parity_run.py --preset c1 --scale 0.2 --samples 40 --split --cov --work /tmp/p2 > gpurun_out/p2.log 2>&1
echo "p2 rc=$?" >> gpurun_out/p2.log
timeout 600 python tools/parity_run.py --preset c4 --scale 0.005 --samples 3 --work /tmp/p3 > gpurun_out/p3.log 2>&1
echo "p3 rc=$?" >> gpurun_out/p3.log
timeout 600 python tools/parity_run.py --preset c5 --scale 0.004 --samples 6 --text --work /tmp/p4 > gpurun_out/p4.log 2>&1
echo "p4 rc=$?" >> gpurun_out/p4.log
tail -n 40 gpurun_out/p1.log gpurun_out/sanitizer.log gpurun_out/p2.log gpurun_out/p3.log gpurun_out/p4.log
