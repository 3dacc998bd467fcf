#!/bin/bash
# development helper: what one gpurun call runs (edit per experiment); this is the round's final validation
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu_r02_final.txt 2>&1
( time timeout 600 python bench.py > $O/r02_bench_final.json 2> $O/r02_bench_final.err ) 2> $O/r02_bench_final.time
ls -la $O > $O/ls.txt
