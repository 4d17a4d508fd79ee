#!/bin/bash
# Round-2 multi-GPU session (N GPUs of one box): headline bench (weak scaling, DP parity, all-reduce cost), cfg3, cfg4, cfg5.
cd "$(dirname "$0")/.."
N=${1:-8}
O=gpurun_out
mkdir -p $O
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > $O/r2_bench_${2}_${N}gpu.json 2> $O/r2_bench_${2}_${N}gpu.err; echo "$2 rc=$?"; }
run 29521 cfg2 --steps 20 --warmup 5
run 29522 cfg3 --mode cfg3
run 29523 cfg4 --mode cfg4 --steps 2
run 29524 cfg5 --mode cfg5 --steps 5
grep -h "^{" $O/r2_bench_*_${N}gpu.json | cut -c1-200
