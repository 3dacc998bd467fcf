#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu_r02d.txt 2>&1
( time timeout 1200 python bench.py > $O/r02_bench_final.json 2> $O/r02_bench_final.err ) 2> $O/r02_bench_final.time
timeout 900 ncu --set full --import-source on --clock-control none -k regex:solve_cluster_kernel -c 1 -f -o $O/r02f_cluster python tools/phase_report.py --tracks 148 --n 16384 --m 7447 > $O/r02f_ncu.log 2>&1
ls -la $O > $O/ls.txt
