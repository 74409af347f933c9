#!/bin/bash
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 15 gpurun_out/r2e_pytest.log | cut -c1-300
timeout 900 python tools/variant_sweep.py --no-hits --settings ":::::,:::::1,:::::2,:::::4,:::::8,:::::3,:::::7,:::::15,:::::16,:::::32,:::::64" > gpurun_out/r2e_ablate_c2.txt 2> gpurun_out/r2e_ablate_c2.err
echo "ablate c2 rc=$?"; cat gpurun_out/r2e_ablate_c2.txt
timeout 900 python tools/variant_sweep.py --no-hits --preset c4 --settings "::96:::,::128:::,::160:::,::255:::,::96:::1,::96:::131,::96:::135,::96:::143,::255:::143" > gpurun_out/r2e_ablate_c4.txt 2> gpurun_out/r2e_ablate_c4.err
echo "ablate c4 rc=$?"; cat gpurun_out/r2e_ablate_c4.txt; grep "msnv:" gpurun_out/r2e_ablate_c4.err | sort | uniq -c | cut -c1-250
