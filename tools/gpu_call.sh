#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/c_ab.txt
for v in "" _c10 _c12; do
  echo "variant '$v'" >> $O/c_ab.txt
  RL_LIB_VARIANT=$v timeout 300 python tools/sweep_probe.py 0 2>&1 | tail -1 | cut -c1-200 >> $O/c_ab.txt
done
ls -la $O > $O/ls.txt
