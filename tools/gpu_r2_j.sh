#!/bin/bash
# round 2: where the floor of the persistent pileup kernel comes from (stage count, early release, ablations)
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_abi.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/r2j_pytest.log | cut -c1-300
timeout 1200 python tools/variant_sweep.py --settings ":::::,:::::21,:::::29,:::::53,:::::61,:::::125,::::::3,:::::21:3,:::::61:3,::::::4,:::::61:4,:3312::::,:3312:::::3,:3312:::61:3" > gpurun_out/r2j_sweep_c2.txt 2> gpurun_out/r2j_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2j_sweep_c2.txt; grep "msnv:" gpurun_out/r2j_sweep_c2.err | sort | uniq -c | cut -c1-250
