#!/bin/bash
# round 2: the gather form of the pileup kernel: parity suite with it as the default, then both kernels side by side
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hand_written or golden_fixture" --timeout 120 > gpurun_out/r2w_quick.log 2>&1
rc=$?; echo "quick rc=$rc"; tail -n 15 gpurun_out/r2w_quick.log | cut -c1-300
if [ $rc -ne 0 ]; then
  MSNV_PILEUP=scatter timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hand_written or golden_fixture" --timeout 120 > gpurun_out/r2w_quick_scatter.log 2>&1
  echo "quick scatter rc=$?"; tail -n 5 gpurun_out/r2w_quick_scatter.log | cut -c1-300
fi
timeout 700 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 15 gpurun_out/r2w_pytest.log | cut -c1-300
timeout 400 python tools/variant_sweep.py --settings "::::::gather,::::::scatter,:::::1:gather,:::::8:gather,:::::9:gather,3::::::gather" > gpurun_out/r2w_sweep_c2.txt 2> gpurun_out/r2w_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2w_sweep_c2.txt; grep "msnv:" gpurun_out/r2w_sweep_c2.err | sort | uniq -c | cut -c1-250
timeout 400 python tools/variant_sweep.py --preset c4 --settings "::::::gather,::::::scatter,::::128::gather" > gpurun_out/r2w_sweep_c4.txt 2> gpurun_out/r2w_sweep_c4.err
echo "sweep c4 rc=$?"; cat gpurun_out/r2w_sweep_c4.txt; grep "msnv:" gpurun_out/r2w_sweep_c4.err | sort | uniq -c | cut -c1-250
