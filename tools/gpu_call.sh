#!/bin/bash
# development helper: what one gpurun call runs (edit per experiment); this is the round's final validation
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu_r02_final.txt 2>&1
timeout 900 python bench.py --tracks-total 8192 --steps 3 --warmup 3 --no-cpu-baseline --long-tracks-total 1184 --extra-steps 2 > $O/q_bench.json 2> $O/q_bench.err
ls -la $O > $O/ls.txt
