#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_cluster.py -m gpu -q -x > $O/q_pytest.txt 2>&1
timeout 900 python bench.py --tracks-total 1184 --steps 2 --warmup 3 --no-cpu-baseline --long-tracks-total 2368 --extra-steps 2 > $O/q_bench.json 2> $O/q_bench.err
ls -la $O > $O/ls.txt
