#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/startup.txt 2>&1
import os, sys, time, subprocess
sys.path.insert(0, os.getcwd())
from metasnv_b200 import harness as H
from metasnv_b200.paths import bin_path
d="/tmp/st"; H.synth(d, "c1", 0.05, 16)
for k in range(2):
    t0=time.time(); rc,err=H.run_product_snpcall(d, d+"/o", env=dict(os.environ, MSNV_VERBOSE="1")); print("run",k,"wall",time.time()-t0); print(err)
t0=time.time(); rc,err=H.run_product_snpcall(d, d+"/o2", env=dict(os.environ, MSNV_VERBOSE="1", MSNV_CLEAN_EXIT="1")); print("clean exit wall",time.time()-t0); print(err)
bam=open(d+"/all_samples").readline().strip()
for k in range(2):
    t0=time.time(); r=H.run_qacompute(bin_path("qaCompute"), bam, d+"/g.cov"); print("qaCompute wall", time.time()-t0)
t0=time.time(); subprocess.run(["python","-c","import ctypes,time;t=time.time();l=ctypes.CDLL('metasnv_b200/lib/libmsnv_gpu.so');l.msnv_device_count.restype=ctypes.c_int;print('count',l.msnv_device_count(),time.time()-t);h=ctypes.c_void_p();t=time.time();print(l.msnv_create(0,ctypes.byref(h)),time.time()-t)"]); print("py ctx", time.time()-t0)
print(subprocess.run(["nvidia-smi","-q","-d","PERFORMANCE"],capture_output=True,text=True).stdout[:300])
print(subprocess.run(["nvidia-smi","--query-gpu=persistence_mode","--format=csv"],capture_output=True,text=True).stdout)
PY
cat gpurun_out/startup.txt | cut -c1-200
timeout 900 python bench.py --workload c3 --scale 0.03 --steps 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench c3 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c3.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'])"
timeout 900 python bench.py --workload c5 --scale 0.15 --steps 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "bench c5 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_c5.json'));print(d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))"; tail -3 gpurun_out/bench_c5.err
