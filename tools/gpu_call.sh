#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --tracks-total 8192 --steps 3 --warmup 2 --no-cpu-baseline --long-tracks-total 1024 --extra-steps 2 > $O/bench.json 2> $O/bench.err
timeout 2000 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1
ls -la $O > $O/ls.txt
