#!/bin/bash
# round 2, second pass: parity, then sweeps of the staging choices at the C2 and C4 shapes
mkdir -p gpurun_out
export MSNV_VERBOSE=1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/r2c_pytest.log
timeout 900 python tools/variant_sweep.py --settings ":::,:::2000,:::20000,3:::,4:3200::,5:2560:160:" > gpurun_out/r2c_sweep_c2.txt 2> gpurun_out/r2c_sweep_c2.err
echo "sweep c2 rc=$?"; cat gpurun_out/r2c_sweep_c2.txt; grep "msnv:" gpurun_out/r2c_sweep_c2.err | sort | uniq -c
timeout 900 python tools/variant_sweep.py --preset c4 --settings ":::,:::2000,::96:,::160:,::255:,3::128:" > gpurun_out/r2c_sweep_c4.txt 2> gpurun_out/r2c_sweep_c4.err
echo "sweep c4 rc=$?"; cat gpurun_out/r2c_sweep_c4.txt; grep "msnv:" gpurun_out/r2c_sweep_c4.err | sort | uniq -c
