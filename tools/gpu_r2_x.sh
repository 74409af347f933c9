#!/bin/bash
# round 2: gather kernel with the set-up on the consumer warps (a thread per read, two barriers), pairs of quads per thread
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -x -q -m gpu -k "hand_written or golden_fixture or full_size or device_synth" --timeout 200 > gpurun_out/r2x_quick.log 2>&1
rc=$?; echo "quick rc=$rc"; tail -n 15 gpurun_out/r2x_quick.log | cut -c1-300
timeout 400 python tools/variant_sweep.py --settings "::::::gather,::::::scatter,:::::1:gather,:::::3:gather,:::::8:gather,:::::11:gather,3::::::gather" > gpurun_out/r2x_sweep_c2.txt 2> gpurun_out/r2x_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2x_sweep_c2.txt; grep "msnv:" gpurun_out/r2x_sweep_c2.err | sort | uniq -c | cut -c1-250
timeout 400 python tools/variant_sweep.py --preset c4 --settings "::::::gather,::::::scatter,::::128::gather" > gpurun_out/r2x_sweep_c4.txt 2> gpurun_out/r2x_sweep_c4.err
echo "sweep c4 rc=$?"; cat gpurun_out/r2x_sweep_c4.txt; grep "msnv:" gpurun_out/r2x_sweep_c4.err | sort | uniq -c | cut -c1-250
