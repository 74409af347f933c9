#!/bin/bash
# round 2: gather kernel with the two-deep software pipeline and a sleeping producer: parity, sweep, ncu
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py tests/test_gpu_windows.py -x -q -m gpu -k "hand_written or golden_fixture or full_size or device_synth or windows" --timeout 200 > gpurun_out/r2ab_quick.log 2>&1
rc=$?; echo "quick rc=$rc"; tail -n 15 gpurun_out/r2ab_quick.log | cut -c1-300
timeout 600 python tools/variant_sweep.py --settings "::::::gather::1000,::::::gather::0,::::::gather::200,::::::gather::3000,::::::scatter::1000,::::::scatter::0" > gpurun_out/r2ab_sweep_c2.txt 2> gpurun_out/r2ab_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2ab_sweep_c2.txt; grep "msnv:" gpurun_out/r2ab_sweep_c2.err | uniq -c | cut -c1-250
timeout 400 python tools/variant_sweep.py --preset c4 --settings "::::::::1000,::::::::0" > gpurun_out/r2ab_sweep_c4.txt 2> gpurun_out/r2ab_sweep_c4.err
echo "sweep c4 rc=$?"; cat gpurun_out/r2ab_sweep_c4.txt
unset MSNV_VERBOSE
B="python bench.py --steps 1 --no-e2e --no-e2e-h2d --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pileup_gather_kernel' -s 3 -c 1 -f -o gpurun_out/r2ab_prof_c2 $B --samples 200 > gpurun_out/r2ab_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"; tail -2 gpurun_out/r2ab_ncu_c2.log
