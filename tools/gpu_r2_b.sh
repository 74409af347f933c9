#!/bin/bash
# round 2: ncu captures of the persistent pileup kernel (C2 shape at 200 samples, C4 shape at scale 0.1)
mkdir -p gpurun_out
export MSNV_VERBOSE=1
B="python bench.py --steps 1 --no-e2e --no-cpu-baseline"
MSNV_PILEUP_CTAS=4 MSNV_CHUNK_Q4=3200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pileup_kernel' -s 3 -c 1 -f -o gpurun_out/r2b_prof_c2 $B --samples 200 > gpurun_out/r2b_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"; grep "msnv:" gpurun_out/r2b_ncu_c2.log | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pileup_kernel' -s 3 -c 1 -f -o gpurun_out/r2b_prof_c4 $B --workload c4 --scale 0.1 > gpurun_out/r2b_ncu_c4.log 2>&1
echo "ncu c4 rc=$?"; grep "msnv:" gpurun_out/r2b_ncu_c4.log | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'call_kernel' -s 3 -c 1 -f -o gpurun_out/r2b_prof_call $B --samples 200 > gpurun_out/r2b_ncu_call.log 2>&1
echo "ncu call rc=$?"
