#!/bin/bash
# builds tools/probes/* : PGD loop probe in several configurations (development)
cd "$(dirname "$0")/.."
mkdir -p tools/probes
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
b() { # name variant T K minb
  $NV -DRL_PGD_VARIANT=$2 -DPROBE_V=$2 -DPROBE_T=$3 -DPROBE_K=$4 -DPROBE_MINB=$5 -o tools/probes/$1 tools/pgd_probe.cu || echo "build $1 failed"
}
for spec in "$@"; do b $spec & done
wait
ls tools/probes
