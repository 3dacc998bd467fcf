#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/k4_ab.txt
for v in _k8 ""; do
  for nm in "300 136" "500 227" "700 318" "1000 455"; do
    set -- $nm
    echo "variant '$v' N=$1" >> $O/k4_ab.txt
    RL_LIB_VARIANT=$v timeout 300 python tools/phase_report.py --tracks 4096 --n $1 --m $2 2>&1 | head -1 >> $O/k4_ab.txt
  done
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chains.py tests/test_gpu_fuzz.py tests/test_gpu_adversarial.py -m gpu -q > $O/q_pytest.txt 2>&1
ls -la $O > $O/ls.txt
