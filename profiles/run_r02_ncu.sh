#!/bin/bash
# ncu captures of the round-2 kernels (one B200): full sets with source for the three MLP kernels, the ray-side kernels at the
# cfg3 chunk size, and the launch list of an eager bench run.  Output: gpurun_out/r02_*.ncu-rep / .csv (summarised by profiles/summarize_ncu.py).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp3_forward -s 2 -c 1 -f -o $O/r02_fwd_nosave python profiles/prof_mlp_fwd.py > $O/ncu_r02_fwd_nosave.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp3_forward -s 4 -c 1 -f -o $O/r02_fwd_save python profiles/prof_mlp_fwd.py > $O/ncu_r02_fwd_save.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp3_backward_kernel -s 2 -c 1 -f -o $O/r02_bwd python profiles/prof_mlp_bwd.py > $O/ncu_r02_bwd.log 2>&1
if [ "$1" != "mlp" ]; then
timeout 600 ncu --set full --clock-control none -k regex:"composite|sample_fine|quantize|dequantize|coarse_depths|absmax|to8b" -c 40 -f -o $O/r02_ray_kernels python profiles/prof_ray_kernels.py --once > $O/ncu_r02_ray.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv python bench.py --eager --steps 2 --warmup 3 > $O/ncu_r02_launches.log 2>&1
fi
ls -la $O/*.ncu-rep
