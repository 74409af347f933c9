#!/bin/bash
# round 2: the kernel-forms agreement test, the driver's smoke()
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_abi.py -x -q -m gpu -k "forms" --timeout 300 > gpurun_out/r2ad_forms.log 2>&1
echo "forms rc=$?"; tail -n 15 gpurun_out/r2ad_forms.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ad_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 3 gpurun_out/r2ad_smoke.log | cut -c1-300
