#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_debug_build.py tests/test_gpu_cluster.py tests/test_geom_stage.py -m gpu -x -q > $O/pytest_new.txt 2>&1
echo "rc=$?" >> $O/pytest_new.txt
timeout 900 python bench.py --tracks-total 2368 --steps 2 --warmup 1 --no-cpu-baseline --long-tracks-total 1024 --extra-steps 2 > $O/bench_small.json 2> $O/bench_small.err
echo "rc=$?" >> $O/bench_small.err
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1
ls -la $O > $O/ls.txt
