#!/bin/bash
# scratch driver for one gpurun call (development); every step under its own timeout, logs into gpurun_out/
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1
for c in 4 8 16; do
  timeout 600 python bench.py --tracks-total 8192 --steps 3 --warmup 2 --no-cpu-baseline --no-extras --solve-chunks $c > $O/bench_c$c.json 2> $O/bench_c$c.err
done
( time timeout 1500 python bench.py --steps 3 --warmup 2 ) > $O/bench_full.json 2> $O/bench_full.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/bench_ref.json 2> $O/bench_ref.err
ls -la $O > $O/ls.txt
