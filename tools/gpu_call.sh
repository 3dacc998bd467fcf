#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_cluster.py -m gpu -q -x > $O/q_pytest.txt 2>&1
RL_LIB_VARIANT=_ph timeout 600 python tools/phase_report.py --tracks 148 --n 16384 --m 7447 > $O/phase_cluster.txt 2>&1
timeout 900 python bench.py --tracks-total 1184 --steps 2 --warmup 3 --no-cpu-baseline --long-tracks-total 1184 --extra-steps 2 > $O/q_bench.json 2> $O/q_bench.err
ls -la $O > $O/ls.txt
