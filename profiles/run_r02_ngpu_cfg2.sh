#!/bin/bash
# N-GPU check of the headline bench line only (weak scaling, dp_parity, exposed exchange): bash profiles/run_r02_ngpu_cfg2.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_bench_cfg2_${N}gpu_final2.json 2> $O/r2_bench_cfg2_${N}gpu_final2.err; echo "cfg2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --impl reference --steps 1 --warmup 1 > $O/r2_bench_ref_${N}gpu_final2.json 2> $O/r2_bench_ref_${N}gpu_final2.err; echo "ref rc=$?"
grep -h "^{" $O/r2_bench_cfg2_${N}gpu_final2.json $O/r2_bench_ref_${N}gpu_final2.json | cut -c1-400
tail -3 $O/r2_bench_cfg2_${N}gpu_final2.err
