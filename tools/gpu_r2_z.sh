#!/bin/bash
# round 2: ncu full set of the gather kernel at the C2 shape (200 samples)
mkdir -p gpurun_out
B="python bench.py --steps 1 --no-e2e --no-e2e-h2d --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pileup_gather_kernel' -s 3 -c 1 -f -o gpurun_out/r2z_prof_c2 $B --samples 200 > gpurun_out/r2z_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"; tail -3 gpurun_out/r2z_ncu_c2.log
