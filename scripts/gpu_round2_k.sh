#!/bin/bash
# 2-GPU call: configs[4] (PPO continuous) with the LL-word exchange (exchange regions 32-byte aligned)
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29531 bench.py --gpus 2 --config ppo_continuous --steps 2 --warmup 1 > $O/r02_bench_ppo_continuous_n2.json 2> $O/r02_bench_cont_n2.err; echo "cont rc=$?"
tail -c 1500 $O/r02_bench_ppo_continuous_n2.json; grep -v "^\[W\|^W0\|OMP_NUM\|^\*\*\*" $O/r02_bench_cont_n2.err | tail -5 | cut -c1-300
