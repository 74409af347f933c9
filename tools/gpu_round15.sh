#!/bin/bash
mkdir -p gpurun_out
export MSNV_PILEUP_VARIANT=x2 MSNV_CHUNK_Q4=3072
B="python bench.py --scale 0.5 --steps 2 --no-e2e --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'pileup_kernel' -s 3 -c 1 -f -o gpurun_out/prof_v3b $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
unset MSNV_PILEUP_VARIANT MSNV_CHUNK_Q4
timeout 600 python tools/variant_sweep.py --preset c4 --settings default,x2,x2:4096,x6:4096,l:4096,x5:4096 > gpurun_out/sweep_c4.txt 2> gpurun_out/sweep_c4.err; cat gpurun_out/sweep_c4.txt; tail -3 gpurun_out/sweep_c4.err
timeout 600 python tools/variant_sweep.py --preset c3 --scale 0.03 --settings default,x2,x2:3072,x2:2048,x3:2048,x4:2048 > gpurun_out/sweep_c3.txt 2> gpurun_out/sweep_c3.err; cat gpurun_out/sweep_c3.txt; tail -3 gpurun_out/sweep_c3.err
