#!/bin/bash
# 2-GPU call: exchange check (short asserted / long), exchange timeline, bench N=2
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
JB_NENV=64 JB_NEPOCH=1 timeout 400 $TR --master-port 29521 scripts/multigpu_check.py > $O/r02_multigpu_check_n2_short.txt 2>&1; echo "mg2 short rc=$?"
grep "^rank\|Error\|assert" $O/r02_multigpu_check_n2_short.txt | tail -14 | cut -c1-330
timeout 400 $TR --master-port 29522 scripts/multigpu_check.py > $O/r02_multigpu_check_n2.txt 2>&1; echo "mg2 long rc=$?"
grep "^rank 0\|Error\|assert" $O/r02_multigpu_check_n2.txt | tail -8 | cut -c1-330
timeout 300 $TR --master-port 29525 scripts/perf_trace_mg.py > $O/r02_trace_exchange_n2.txt 2>&1; echo "trace rc=$?"
grep -v "^\[W\|^W0\|^\*\*\*\|^$\|OMP_NUM" $O/r02_trace_exchange_n2.txt | head -60 | cut -c1-200
timeout 600 $TR --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err; echo "bench2 rc=$?"
tail -c 1500 $O/r02_bench_n2.json
