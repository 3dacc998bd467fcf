#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu_r02c.txt 2>&1
ls -la $O > $O/ls.txt
