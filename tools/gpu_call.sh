#!/bin/bash
# development helper: what one gpurun call runs (edit per experiment); here: the bench on all GPUs of the box
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
N=$(nvidia-smi -L | wc -l)
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-extras > $O/r02_bench_${N}gpu_final.json 2> $O/r02_bench_${N}gpu_final.err ) 2> $O/r02_bench_${N}gpu_final.time
