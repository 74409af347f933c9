#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/variant_sweep.py --settings 1,1::0,1::888,default > gpurun_out/sweep_pf_c2.txt 2> gpurun_out/sweep_pf_c2.err; cat gpurun_out/sweep_pf_c2.txt; tail -2 gpurun_out/sweep_pf_c2.err
timeout 600 python tools/variant_sweep.py --preset c4 --settings 2,2::0,4::0,4 > gpurun_out/sweep_pf_c4.txt 2> gpurun_out/sweep_pf_c4.err; cat gpurun_out/sweep_pf_c4.txt; tail -2 gpurun_out/sweep_pf_c4.err
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
MSNV_PF_DIST=0 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'pileup_kernel' -s 3 -c 1 --csv --log-file gpurun_out/traffic_pf0.csv $B > gpurun_out/ncu_t0.log 2>&1
echo "ncu rc=$?"; tail -4 gpurun_out/traffic_pf0.csv | cut -c1-300
