#!/bin/bash
# 1-GPU call: pytest -m gpu, ckpt fixtures, per-kernel ncu list, bench (default config) + functional runs of the other configs
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_b.txt 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r02_pytest_b.txt
timeout 300 python scripts/make_ckpt_fixtures.py 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --config ppo_continuous --n-envs 512 --n-step 64 --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_func_cont.json 2> gpurun_out/r02_func_cont.err; echo "cont rc=$?"
tail -c 1200 gpurun_out/r02_func_cont.json; tail -8 gpurun_out/r02_func_cont.err
timeout 400 python bench.py --config rainbow_frames --buffer 20000 --rounds 4 --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_func_rainbow.json 2> gpurun_out/r02_func_rainbow.err; echo "rainbow rc=$?"
tail -c 1200 gpurun_out/r02_func_rainbow.json; tail -8 gpurun_out/r02_func_rainbow.err
timeout 400 python bench.py --config apex --n-envs 32 --batch 64 --buffer 20000 --rounds 1 --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_func_apex.json 2> gpurun_out/r02_func_apex.err; echo "apex rc=$?"
tail -c 1200 gpurun_out/r02_func_apex.json; tail -8 gpurun_out/r02_func_apex.err
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size --nvtx --nvtx-include "jb/" -f -o gpurun_out/r02_kernels python scripts/ncu_kernels.py > gpurun_out/r02_ncu_kernels.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r02_ncu_kernels.log
ncu -i gpurun_out/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv 2>/dev/null; wc -l gpurun_out/r02_kernels_raw.csv
