#!/bin/bash
# A/B on one box: bash profiles/run_ab_r4.sh <variant to run the GPU tests on> <variants...>   (libraries from profiles/ab_build.py)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; : > $O/ab_r4.log
T=$1; shift
for round in 1 2 3; do
  for v in "$@"; do
    echo "== $v (round $round)" >> $O/ab_r4.log
    NERFQ_LIB=profiles/_ab/$v/libnerfq.so timeout 120 python profiles/time_mlp.py >> $O/ab_r4.log 2>&1
  done
done
cat $O/ab_r4.log
NERFQ_LIB=$PWD/profiles/_ab/$T/libnerfq.so timeout 900 python -m pytest tests -m gpu -q > $O/ab_r4_pytest.log 2>&1; tail -15 $O/ab_r4_pytest.log
