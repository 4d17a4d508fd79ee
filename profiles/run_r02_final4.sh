#!/bin/bash
# Last validation of the round at HEAD: GPU tests, smoke(), default bench line.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2_pytest_final4.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest_final4.log; tail -3 $O/r2_pytest_final4.log
timeout 200 python __graft_entry__.py --smoke > $O/r2_smoke_final4.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2_smoke_final4.log
timeout 400 python bench.py > $O/r2_bench_final4.json 2> $O/r2_bench_final4.err; echo "bench rc=$?"
grep -h "^{" $O/r2_bench_final4.json | cut -c1-400
