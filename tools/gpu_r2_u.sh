#!/bin/bash
# round 2, two GPUs: the shipped multi-GPU path (two splits, two snpCall processes, two devices) and bench.py under torchrun
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2u_env.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_gpus or windows" > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/r2u_pytest.log | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --no-e2e-h2d --e2e-bam-gb 2.0 > gpurun_out/r2u_bench_n2.json 2> gpurun_out/r2u_bench_n2.err
echo "bench n2 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2u_bench_n2.json'));print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['seconds'])"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c3 --scale 0.25 --bins 2 --steps 2 --no-e2e --no-e2e-h2d --no-cpu-baseline > gpurun_out/r2u_bench_c3_n2.json 2> gpurun_out/r2u_bench_c3_n2.err
echo "bench c3 n2 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2u_bench_c3_n2.json'));print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['sharding'], d['config']['windows_per_shard'])"
