#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_counts.py c2 0.003 16 > gpurun_out/dbg.txt 2>&1
tail -60 gpurun_out/dbg.txt
