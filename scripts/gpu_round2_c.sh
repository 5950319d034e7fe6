#!/bin/bash
# 1-GPU call: tensor-core forward phase of the persistent kernel: PPO tests, timeline, bench, per-kernel ncu list
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_c.txt 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r02_pytest_c.txt | cut -c1-300
JB_FUSED_TC=1 timeout 200 python scripts/perf_trace.py > gpurun_out/r02_trace_tc.txt 2>&1; echo "trace rc=$?"; head -45 gpurun_out/r02_trace_tc.txt
JB_FUSED_TC=0 timeout 200 python scripts/perf_trace.py > gpurun_out/r02_trace_ffma.txt 2>&1; head -3 gpurun_out/r02_trace_ffma.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_bench_n1_tc.json 2> gpurun_out/r02_bench_n1_tc.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_n1_tc.json; tail -3 gpurun_out/r02_bench_n1_tc.err
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size --nvtx --profile-from-start off -f -o gpurun_out/r02_kernels python scripts/ncu_kernels.py > gpurun_out/r02_ncu_kernels.log 2>&1; echo "ncu rc=$?"
tail -4 gpurun_out/r02_ncu_kernels.log
ncu -i gpurun_out/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv 2>/dev/null; wc -l gpurun_out/r02_kernels_raw.csv
