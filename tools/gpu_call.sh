#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --import-source on --clock-control none -k regex:solve_kernel -s 3 -c 1 -f -o $O/r02_sweep_kernel python tools/sweep_probe.py 0 > $O/sweep_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:solve_kernel -s 1 -c 1 -f -o $O/r02_n252_kernel python tools/phase_report.py --tracks 4096 --n 252 --m 115 > $O/n252_ncu.log 2>&1
ls -la $O > $O/ls.txt
