#!/bin/bash
# round 2: gather kernel: more bytes in flight (stages of the ring, chunk sizes, CTA shapes); mate pass with the cheap reject
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -x -q -m gpu -k "hand_written or golden_fixture or full_size or device_synth" --timeout 200 > gpurun_out/r2y_quick.log 2>&1
rc=$?; echo "quick rc=$rc"; tail -n 15 gpurun_out/r2y_quick.log | cut -c1-300
MSNV_STAGES=3 MSNV_MAX_READS=60 MSNV_CHUNK_Q4=1664 timeout 300 python -m pytest tests/test_gpu_abi.py -x -q -m gpu -k "full_size" --timeout 200 > gpurun_out/r2y_quick_split.log 2>&1
echo "quick split rc=$?"; tail -n 5 gpurun_out/r2y_quick_split.log | cut -c1-300
timeout 600 python tools/variant_sweep.py --settings "::::::gather:2,4:1664:60::::gather:3,4:1104:40::::gather:4,4:1664:60::::gather:2,2::::256::gather:4,2::::::gather:4,3:2208:80::::gather:3,3::::::gather:2,2::::256::gather:3" > gpurun_out/r2y_sweep_c2.txt 2> gpurun_out/r2y_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2y_sweep_c2.txt; grep "msnv:" gpurun_out/r2y_sweep_c2.err | uniq -c | cut -c1-250
timeout 400 python tools/variant_sweep.py --preset c4 --settings "::::::gather:2,::::::gather:3,3::::::gather:2" > gpurun_out/r2y_sweep_c4.txt 2> gpurun_out/r2y_sweep_c4.err
echo "sweep c4 rc=$?"; cat gpurun_out/r2y_sweep_c4.txt; grep "msnv:" gpurun_out/r2y_sweep_c4.err | uniq -c | cut -c1-250
