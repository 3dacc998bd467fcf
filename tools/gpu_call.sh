#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py --tracks-total 1184 --steps 2 --warmup 3 --no-cpu-baseline --long-tracks-total 0 --extra-steps 2 > $O/q_bench.json 2> $O/q_bench.err
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chains.py tests/test_gpu_debug_build.py tests/test_gpu_fuzz.py -m gpu -q > $O/q_pytest.txt 2>&1
ls -la $O > $O/ls.txt
