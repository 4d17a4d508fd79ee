#!/bin/bash
# Round-2 final single-GPU session: GPU tests, smoke(), the bench line, MLP kernel timing, ray-kernel timing, ncu captures.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r2_pytest_final.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest_final.log; tail -3 $O/r2_pytest_final.log
timeout 300 python __graft_entry__.py --smoke > $O/r2_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2_smoke_final.log
timeout 600 python bench.py > $O/r2_bench_final.json 2> $O/r2_bench_final.err; echo "bench rc=$?"
timeout 300 python profiles/time_mlp.py > $O/r2_time_mlp_final.log 2>&1; cat $O/r2_time_mlp_final.log
timeout 300 python profiles/prof_ray_kernels.py > $O/r2_ray_kernels_timing_final.log 2>&1
bash profiles/run_r02_ncu.sh all > $O/r2_ncu_final.log 2>&1; tail -6 $O/r2_ncu_final.log
