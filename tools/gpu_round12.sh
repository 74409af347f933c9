#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/variant_sweep.py --settings default,l,s,x2,x3,x4,x5,x6,l:3072,l:3584,x4:3072,x4:3584,x5:2560,x5:3072,s:2560,s:3072 > gpurun_out/sweep_c2.txt 2> gpurun_out/sweep_c2.err; cat gpurun_out/sweep_c2.txt; tail -3 gpurun_out/sweep_c2.err
B="python bench.py --steps 2 --no-e2e --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'pileup_kernel' -s 3 -c 1 -f -o gpurun_out/prof_v3_c2_full $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
