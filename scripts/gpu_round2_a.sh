#!/bin/bash
# 2-GPU call: exchange check + N=2 bench, then GPU0: pytest -m gpu, GPU1: ckpt fixtures + per-kernel ncu list
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
JB_NENV=64 JB_NEPOCH=1 timeout 400 $TR --master-port 29521 scripts/multigpu_check.py > gpurun_out/r02_mg2_short.txt 2>&1; echo "mg2 short rc=$?"
grep -v "^\[W\|^W0\|^\*\*\*\|^$" gpurun_out/r02_mg2_short.txt | tail -25
timeout 400 $TR --master-port 29522 scripts/multigpu_check.py > gpurun_out/r02_mg2_long.txt 2>&1; echo "mg2 long rc=$?"
grep "rank 0" gpurun_out/r02_mg2_long.txt | tail -8
JB_BENCH_TRACE=1 timeout 400 $TR --master-port 29523 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench2 rc=$?"
tail -c 1800 gpurun_out/r02_bench_n2.json; grep -v "^\[W\|^W0\|^\*\*\*" gpurun_out/r02_bench_n2.err | tail -15
(CUDA_VISIBLE_DEVICES=0 timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_a.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_a.txt) &
PT=$!
CUDA_VISIBLE_DEVICES=1 timeout 300 python scripts/make_ckpt_fixtures.py 2>&1 | tail -6
CUDA_VISIBLE_DEVICES=1 timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size --nvtx --nvtx-include "jb/" -f -o gpurun_out/r02_kernels python scripts/ncu_kernels.py > gpurun_out/r02_ncu_kernels.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r02_ncu_kernels.log
ncu -i gpurun_out/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv 2>/dev/null; wc -l gpurun_out/r02_kernels_raw.csv
wait $PT
tail -30 gpurun_out/r02_pytest_a.txt
