#!/bin/bash
# round 2: gather kernel without the branch in the loop, parked producer; mate pass with per-lane pair geometry
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -x -q -m gpu -k "hand_written or golden_fixture or full_size or device_synth" --timeout 200 > gpurun_out/r2aa_quick.log 2>&1
rc=$?; echo "quick rc=$rc"; tail -n 15 gpurun_out/r2aa_quick.log | cut -c1-300
timeout 600 python tools/variant_sweep.py --settings "::::::gather::1000,::::::gather::0,::::::gather::300,::::::gather::4000,:::20:::gather::1000,::::::scatter::1000,::::::scatter::0" > gpurun_out/r2aa_sweep_c2.txt 2> gpurun_out/r2aa_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2aa_sweep_c2.txt; grep "msnv:" gpurun_out/r2aa_sweep_c2.err | uniq -c | cut -c1-250
timeout 400 python tools/variant_sweep.py --preset c4 --settings "::::::::1000,::::::::0" > gpurun_out/r2aa_sweep_c4.txt 2> gpurun_out/r2aa_sweep_c4.err
echo "sweep c4 rc=$?"; cat gpurun_out/r2aa_sweep_c4.txt; grep "msnv:" gpurun_out/r2aa_sweep_c4.err | uniq -c | cut -c1-250
