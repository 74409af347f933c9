#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/e2e_startup_ab.py > gpurun_out/r2ah_ab_2gb.txt 2> gpurun_out/r2ah_ab_2gb.err
echo "ab rc=$?"; cat gpurun_out/r2ah_ab_2gb.txt | cut -c1-700
