#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_quick.log
timeout 900 python tools/variant_sweep.py --settings default,default::31,default::15,default::7,default::3,default::2,default::1,s,x3,x4,x6 > gpurun_out/sweep_c2.txt 2> gpurun_out/sweep_c2.err; cat gpurun_out/sweep_c2.txt; tail -3 gpurun_out/sweep_c2.err
