#!/bin/bash
# A/B on one box: backward with the job's addresses pinned in registers (NERFQ_BWD_PIN), saved activations a whole job ahead (NERFQ_BWD_HH=5),
# and the 32 / 112 register split (last, under its own short timeout: a wrong split blocks in setmaxnreg.inc).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; : > $O/ab_r6.log
for round in 1 2; do
  for v in base pin hh5pin; do
    echo "== $v (round $round)" >> $O/ab_r6.log
    NERFQ_LIB=profiles/_ab/$v/libnerfq.so timeout 120 python profiles/time_mlp.py >> $O/ab_r6.log 2>&1
  done
done
NERFQ_LIB=$PWD/profiles/_ab/hh5pin/libnerfq.so timeout 300 python -m pytest tests -m gpu -q -x > $O/ab_r6_pytest.log 2>&1; tail -3 $O/ab_r6_pytest.log
echo "== hh5pin112" >> $O/ab_r6.log
NERFQ_LIB=profiles/_ab/hh5pin112/libnerfq.so timeout -s KILL 60 python profiles/time_mlp.py >> $O/ab_r6.log 2>&1; echo "rc=$?" >> $O/ab_r6.log
cat $O/ab_r6.log
