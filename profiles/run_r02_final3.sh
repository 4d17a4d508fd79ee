#!/bin/bash
# Round-2 closing session after the forward-kernel work: GPU tests, smoke(), bench (cfg2, cfg3), MLP kernel timing, ncu of the two forward kernels and the backward.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest_final3.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest_final3.log; tail -3 $O/r2_pytest_final3.log
timeout 300 python __graft_entry__.py --smoke > $O/r2_smoke_final3.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2_smoke_final3.log
timeout 600 python bench.py > $O/r2_bench_final3.json 2> $O/r2_bench_final3.err; echo "bench rc=$?"
timeout 600 python bench.py --mode cfg3 > $O/r2_bench_cfg3_final3.json 2> $O/r2_bench_cfg3_final3.err; echo "cfg3 rc=$?"
timeout 300 python profiles/time_mlp.py > $O/r2_time_mlp_final3.log 2>&1; cat $O/r2_time_mlp_final3.log
bash profiles/run_r02_ncu.sh mlp > $O/r2_ncu_final3.log 2>&1; tail -4 $O/r2_ncu_final3.log
grep -h "^{" $O/r2_bench_final3.json $O/r2_bench_cfg3_final3.json | cut -c1-300
