#!/bin/bash
# development helper: what one gpurun call runs (edit per experiment); here: the pipelined geometry batch
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_geom_stage.py tests/test_abi_and_host.py -m gpu -q > $O/pytest_geom_pipe.txt 2>&1
timeout 300 python bench.py --tracks-total 8192 --long-tracks-total 0 --no-cpu-baseline --steps 2 --warmup 3 > $O/geom_pipe_bench.json 2> $O/geom_pipe_bench.err
