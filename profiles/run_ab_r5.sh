#!/bin/bash
# A/B on one box, two interleaved rounds: bash profiles/run_ab_r5.sh <variant to run the GPU tests on> <variants...>   (libraries from profiles/ab_build.py)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; : > $O/ab_r5.log
T=$1; shift
for round in 1 2; do
  for v in "$@"; do
    echo "== $v (round $round)" >> $O/ab_r5.log
    NERFQ_LIB=profiles/_ab/$v/libnerfq.so timeout 120 python profiles/time_mlp.py >> $O/ab_r5.log 2>&1
  done
done
cat $O/ab_r5.log
NERFQ_LIB=$PWD/profiles/_ab/$T/libnerfq.so timeout 600 python -m pytest tests -m gpu -q -x > $O/ab_r5_pytest.log 2>&1; tail -5 $O/ab_r5_pytest.log
