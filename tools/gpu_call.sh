#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py --tracks-total 8192 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/q_bench.json 2> $O/q_bench.err
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chains.py tests/test_gpu_adversarial.py -m gpu -q > $O/q_pytest.txt 2>&1
ls -la $O > $O/ls.txt
