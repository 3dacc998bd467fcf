#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02_bench_2gpu_final.json 2> $O/r02_bench_2gpu_final.err
timeout 600 python -m pytest tests/test_gpu_resident.py -m gpu -q > $O/q_pytest_2gpu.txt 2>&1
ls -la $O > $O/ls.txt
