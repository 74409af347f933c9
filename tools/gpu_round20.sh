#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/startup2.txt 2>&1
import os, sys, time
sys.path.insert(0, os.getcwd())
from metasnv_b200 import harness as H
d="/tmp/st2"; H.synth(d, "c2", 0.02, 400)
for k in range(4):
    t0=time.time(); rc,err=H.run_product_snpcall(d, d+"/o", env=dict(os.environ, MSNV_VERBOSE="1")); print("run",k,"wall",time.time()-t0); print(err[-900:])
PY
cat gpurun_out/startup2.txt | cut -c1-200
