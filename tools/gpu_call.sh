#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python tools/sweep_probe.py 0 2 3 4 6 7 8 2>&1 | cut -c1-60,150-230 > $O/chain_sweep.txt
ls -la $O > $O/ls.txt
