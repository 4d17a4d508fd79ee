#!/bin/bash
# Round-2 GPU session A (one B200): tests, bench, MLP timing + traces, ncu launch list, ncu --set full of the MLP and ray kernels.
# Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest5.log
timeout 600 python bench.py > $O/r2_bench2.json 2> $O/r2_bench2.err; echo "bench rc=$?"
tail -3 $O/r2_pytest5.log
timeout 300 python profiles/time_mlp.py > $O/r2_time_mlp_v7.log 2>&1
timeout 300 python profiles/trace_mlp_fwd.py > $O/r2_trace_v7.log 2>&1
timeout 300 python profiles/prof_mlp_bwd.py > $O/r2_trace_bwd_v7.log 2>&1
timeout 300 python profiles/prof_ray_kernels.py > $O/r2_ray_kernels_timing.log 2>&1
# ncu: full sets with source for the three MLP kernels (one launch each), then the ray-side kernels
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp3_forward -s 2 -c 1 -f -o $O/r02_fwd_nosave python profiles/prof_mlp_fwd.py > $O/ncu_r02_fwd_nosave.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp3_forward -s 4 -c 1 -f -o $O/r02_fwd_save python profiles/prof_mlp_fwd.py > $O/ncu_r02_fwd_save.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp3_backward_kernel -s 2 -c 1 -f -o $O/r02_bwd python profiles/prof_mlp_bwd.py > $O/ncu_r02_bwd.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"composite|sample_fine|quantize|dequantize|coarse_depths|absmax|to8b" -c 40 -f -o $O/r02_ray_kernels python profiles/prof_ray_kernels.py --once > $O/ncu_r02_ray.log 2>&1
# launch list of the bench (eager so that every launch is visible)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv python bench.py --eager --steps 2 --warmup 3 > $O/ncu_r02_launches.log 2>&1
ls -la $O | tail -20
