#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_cluster.py -m gpu -q -x > $O/q_pytest.txt 2>&1
timeout 900 python bench.py --tracks-total 8192 --steps 3 --warmup 3 --no-cpu-baseline --long-tracks-total 1184 --extra-steps 2 > $O/q_bench.json 2> $O/q_bench.err
RL_LIB_VARIANT=_sc timeout 900 python bench.py --tracks-total 8192 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/q_bench_sc.json 2> $O/q_bench_sc.err
ls -la $O > $O/ls.txt
