#!/bin/bash
# round 2: mate pass with the verdict-only rule (no quality words rebuilt, mask only for partial quads)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -x -q -m gpu -k "hand_written or golden_fixture or device_synth or forms or full_size" --timeout 200 > gpurun_out/r2ae_quick.log 2>&1
rc=$?; echo "quick rc=$rc"; tail -n 6 gpurun_out/r2ae_quick.log | cut -c1-300
timeout 300 python tools/variant_sweep.py --settings "::::::" > gpurun_out/r2ae_sweep_c2.txt 2> gpurun_out/r2ae_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2ae_sweep_c2.txt
timeout 300 python tools/variant_sweep.py --preset c4 --settings "::::::" > gpurun_out/r2ae_sweep_c4.txt 2> gpurun_out/r2ae_sweep_c4.err
echo "sweep c4 rc=$?"; cat gpurun_out/r2ae_sweep_c4.txt
