#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_quick.log
timeout 900 python tools/variant_sweep.py --scale ${SCALE:-0.5} --settings default,x6:3584,x2:3072,x3:3072,x2:3072:1,x2:3072:2,x2:3072:3 > gpurun_out/sweep_dbg.txt 2> gpurun_out/sweep_dbg.err; cat gpurun_out/sweep_dbg.txt; tail -3 gpurun_out/sweep_dbg.err
