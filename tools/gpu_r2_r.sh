#!/bin/bash
# round 2: launch list of the timed steps at the full C2 shape, DRAM traffic of the pileup phase at full size (C2, C4)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4001 -c 70 --csv --log-file gpurun_out/r2r_launches_c2.csv python bench.py --steps 2 --no-e2e --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2r_launches_c2.log 2>&1
echo "launch list rc=$?"; grep -c pileup_kernel gpurun_out/r2r_launches_c2.csv
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'pileup_kernel|mate_kernel|fix_clear_kernel|call_kernel' -s 12 -c 4 --csv --log-file gpurun_out/r2r_traffic_c2.csv python bench.py --steps 1 --no-e2e --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2r_traffic_c2.log 2>&1
echo "traffic c2 rc=$?"; grep -v "^==" gpurun_out/r2r_traffic_c2.csv | cut -d, -f5,13- | tail -14
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'pileup_kernel|mate_kernel|fix_clear_kernel|call_kernel' -s 12 -c 4 --csv --log-file gpurun_out/r2r_traffic_c4.csv python bench.py --workload c4 --steps 1 --no-e2e --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2r_traffic_c4.log 2>&1
echo "traffic c4 rc=$?"; grep -v "^==" gpurun_out/r2r_traffic_c4.csv | cut -d, -f5,13- | tail -14
