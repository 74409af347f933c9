#!/bin/bash
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cigars_that_open or pairing_rules or exotic" --timeout 40 > gpurun_out/r2al_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/r2al_pytest.log | cut -c1-300
