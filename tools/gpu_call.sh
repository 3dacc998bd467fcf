#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_parity.py tests/test_gpu_chains.py tests/test_gpu_fuzz.py tests/test_gpu_adversarial.py -m gpu -q > $O/q_pytest.txt 2>&1
timeout 900 python bench.py --tracks-total 8192 --steps 3 --warmup 3 --no-cpu-baseline --long-tracks-total 1184 --extra-steps 2 > $O/q_bench.json 2> $O/q_bench.err
ls -la $O > $O/ls.txt
