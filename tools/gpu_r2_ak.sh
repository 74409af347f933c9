#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hand_written_case or leading" --timeout 50 > gpurun_out/r2ak_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/r2ak_pytest.log | cut -c1-300
