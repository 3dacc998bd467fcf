#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
P=tools/probes
{
for x in nofast_256_8_b2 v1_256_8_b2; do timeout 120 $P/$x 4 2 110; timeout 120 $P/$x 4 2 220; timeout 120 $P/$x 4 2 110 64; done
} > $O/probes.txt 2>&1
timeout 600 python bench.py --tracks-total 8192 --steps 3 --warmup 2 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1
ls -la $O > $O/ls.txt
