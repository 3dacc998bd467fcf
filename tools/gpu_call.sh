#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu_r02_final.txt 2>&1
( time timeout 1200 python bench.py > $O/r02_bench_final.json 2> $O/r02_bench_final.err ) 2> $O/r02_bench_final.time
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err ) 2> $O/r02_bench_reference.time
timeout 900 ncu --set full --import-source on --clock-control none -k regex:solve_kernel -s 3 -c 1 -f -o $O/r02g_sweep_kernel python tools/sweep_probe.py 0 > $O/r02g_ncu.log 2>&1
ls -la $O > $O/ls.txt
