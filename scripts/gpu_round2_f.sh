#!/bin/bash
# 2-GPU call: in-kernel gradient exchange check (short = asserted vs NCCL path, long = bit-equal weights across ranks),
# the -m gpu 2-rank test, bench N=2 of configs[1] and configs[4]
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
JB_NENV=64 JB_NEPOCH=1 timeout 400 $TR --master-port 29521 scripts/multigpu_check.py > $O/r02_multigpu_check_n2_short.txt 2>&1; echo "mg2 short rc=$?"
grep -v "^\[W\|^W0\|^\*\*\*\|^$" $O/r02_multigpu_check_n2_short.txt | tail -25 | cut -c1-300
timeout 400 $TR --master-port 29522 scripts/multigpu_check.py > $O/r02_multigpu_check_n2.txt 2>&1; echo "mg2 long rc=$?"
grep "rank 0" $O/r02_multigpu_check_n2.txt | tail -8 | cut -c1-300
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q 2>&1 | tail -3
JB_BENCH_TRACE=1 timeout 600 $TR --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err; echo "bench2 rc=$?"
tail -c 2200 $O/r02_bench_n2.json; grep -v "^\[W\|^W0\|^\*\*\*" $O/r02_bench_n2.err | tail -12 | cut -c1-300
timeout 900 $TR --master-port 29524 bench.py --gpus 2 --config ppo_continuous --steps 2 --warmup 1 > $O/r02_bench_ppo_continuous_n2.json 2> $O/r02_bench_cont_n2.err; echo "cont2 rc=$?"
tail -c 2200 $O/r02_bench_ppo_continuous_n2.json; grep -v "^\[W\|^W0\|^\*\*\*" $O/r02_bench_cont_n2.err | tail -8 | cut -c1-300
