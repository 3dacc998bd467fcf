#!/bin/bash
# development helper: what one gpurun call runs (edit per experiment); here: how the pipeline chunks of rl_solve_batch
# are spread over the kernel streams (rl_set_option "chunk_streams"), full 65,536 tracks
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for m in 2 3 1; do
  timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --chunk-streams $m > $O/cstreams_$m.json 2> $O/cstreams_$m.err
done
ls -la $O > $O/ls.txt
