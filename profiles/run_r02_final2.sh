#!/bin/bash
# Round-2 closing single-GPU session: GPU tests, smoke(), both bench arms, launch list of one cfg3 view.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest_final2.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest_final2.log; tail -3 $O/r2_pytest_final2.log
timeout 300 python __graft_entry__.py --smoke > $O/r2_smoke_final2.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2_smoke_final2.log
timeout 600 python bench.py > $O/r2_bench_final2.json 2> $O/r2_bench_final2.err; echo "bench rc=$?"
timeout 600 python bench.py --mode cfg3 > $O/r2_bench_cfg3_final2.json 2> $O/r2_bench_cfg3_final2.err; echo "cfg3 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_ref_final2.json 2> $O/r2_bench_ref_final2.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_cfg3.csv python bench.py --mode cfg3 --steps 1 --warmup 1 > $O/ncu_r02_launches_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"
grep -h "^{" $O/r2_bench_final2.json $O/r2_bench_cfg3_final2.json $O/r2_bench_ref_final2.json | cut -c1-260
