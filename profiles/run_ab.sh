#!/bin/bash
# A/B timing of library variants built by profiles/ab_build.py (same box, interleaved, two rounds), then the GPU tests on the in-tree build.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/ab_times.log
for round in 1 2; do
  for v in "$@"; do
    echo "== $v (round $round)" >> $O/ab_times.log
    NERFQ_LIB=profiles/_ab/$v/libnerfq.so timeout 200 python profiles/time_mlp.py >> $O/ab_times.log 2>&1
  done
done
cat $O/ab_times.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/ab_pytest.log 2>&1; tail -5 $O/ab_pytest.log
