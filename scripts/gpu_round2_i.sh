#!/bin/bash
# N-GPU call (N = $1): exchange check (short), exchange timeline, bench line
set -u
N=${1:-4}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
JB_NENV=64 JB_NEPOCH=1 timeout 400 $TR --master-port 29521 scripts/multigpu_check.py > $O/r02_multigpu_check_n${N}_short.txt 2>&1; echo "mg short rc=$?"
grep "^rank 0\|Error\|assert" $O/r02_multigpu_check_n${N}_short.txt | tail -8 | cut -c1-330
timeout 300 $TR --master-port 29525 scripts/perf_trace_mg.py > $O/r02_trace_exchange_n${N}.txt 2>&1; echo "trace rc=$?"
grep -v "^\[W\|^W0\|^\*\*\*\|^$\|OMP_NUM" $O/r02_trace_exchange_n${N}.txt | head -24 | cut -c1-200
JB_BENCH_TRACE=1 timeout 600 $TR --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_bench_n${N}.json 2> $O/r02_bench_n${N}.err; echo "bench rc=$?"
tail -c 1800 $O/r02_bench_n${N}.json; grep "FAILED\|Error\|Traceback" -A12 $O/r02_bench_n${N}.err | head -40 | cut -c1-300
