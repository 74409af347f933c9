#!/bin/bash
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/r2d_pytest.log
MSNV_CONSUMERS=256 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -x -q -m gpu > gpurun_out/r2d_pytest256.log 2>&1
echo "pytest(256 consumers) rc=$?"; tail -n 3 gpurun_out/r2d_pytest256.log
timeout 900 python tools/variant_sweep.py --settings "::::,::::256,2::::256,4:3200:::256" > gpurun_out/r2d_sweep_c2.txt 2> gpurun_out/r2d_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2d_sweep_c2.txt; grep "msnv:" gpurun_out/r2d_sweep_c2.err | sort | uniq -c
timeout 900 python tools/variant_sweep.py --preset c4 --settings "::::,::96::,::::256,::96::256,::64::256,::255::256" > gpurun_out/r2d_sweep_c4.txt 2> gpurun_out/r2d_sweep_c4.err
echo "sweep c4 rc=$?"; cat gpurun_out/r2d_sweep_c4.txt; grep "msnv:" gpurun_out/r2d_sweep_c4.err | sort | uniq -c
