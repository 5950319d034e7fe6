#!/bin/bash
# 2-GPU call: configs[4] (PPO continuous) and configs[3] (Ape-X, sharded PER) bench lines with the LL-word exchange
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 170 $TR --master-port 29531 bench.py --gpus 2 --config ppo_continuous --steps 2 --warmup 1 > $O/r02_bench_ppo_continuous_n2.json 2> $O/r02_bench_cont_n2.err; echo "cont rc=$?"
tail -c 1500 $O/r02_bench_ppo_continuous_n2.json; tail -3 $O/r02_bench_cont_n2.err | cut -c1-300
timeout 120 $TR --master-port 29533 bench.py --gpus 2 --config apex --steps 3 --warmup 1 > $O/r02_bench_apex_n2.json 2> $O/r02_bench_apex_n2.err; echo "apex rc=$?"
tail -c 1500 $O/r02_bench_apex_n2.json; tail -3 $O/r02_bench_apex_n2.err | cut -c1-300
