#!/bin/bash
# 1-GPU measurement bundle (round 2): tests, timelines, bench lines of the 1-GPU configurations, per-kernel ncu list,
# full ncu capture + launch list of the dominant kernel.  Outputs -> gpurun_out/r02_*.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_pytest.txt 2>&1; echo "pytest rc=$?"
tail -25 $O/r02_pytest.txt | cut -c1-400
JB_FUSED_TC=1 timeout 200 python scripts/perf_trace.py > $O/r02_trace_tc.txt 2>&1; echo "trace rc=$?"; head -42 $O/r02_trace_tc.txt
JB_FUSED_TC=0 timeout 200 python scripts/perf_trace.py > $O/r02_trace_ffma.txt 2>&1; head -2 $O/r02_trace_ffma.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/r02_clocks_bench.csv &
SMI=$!
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?"
kill $SMI
tail -c 2600 $O/r02_bench_n1.json; tail -3 $O/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_n1_reference_arm.json 2>/dev/null; tail -c 500 $O/r02_bench_n1_reference_arm.json
timeout 900 python bench.py --config rainbow_frames --steps 3 --warmup 1 > $O/r02_bench_rainbow_frames_n1.json 2> $O/r02_bench_rainbow.err; echo "rainbow rc=$?"
tail -c 1800 $O/r02_bench_rainbow_frames_n1.json; tail -4 $O/r02_bench_rainbow.err
timeout 900 python bench.py --config apex --steps 3 --warmup 1 > $O/r02_bench_apex_n1.json 2> $O/r02_bench_apex.err; echo "apex rc=$?"
tail -c 1800 $O/r02_bench_apex_n1.json; tail -4 $O/r02_bench_apex.err
timeout 1200 python bench.py --config ppo_continuous --steps 2 --warmup 1 > $O/r02_bench_ppo_continuous_n1.json 2> $O/r02_bench_cont.err; echo "cont rc=$?"
tail -c 1800 $O/r02_bench_ppo_continuous_n1.json; tail -4 $O/r02_bench_cont.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 900 ncu --clock-control none --metrics $M --nvtx --profile-from-start off -f -o $O/r02_kernels python scripts/ncu_kernels.py > $O/r02_ncu_kernels.log 2>&1; echo "ncu rc=$?"
tail -3 $O/r02_ncu_kernels.log
ncu -i $O/r02_kernels.ncu-rep --page raw --csv > $O/r02_kernels_raw.csv 2>/dev/null; wc -l $O/r02_kernels_raw.csv
N_ENVS=4096 T=128 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches_bench.csv python scripts/ncu_ppo.py > $O/r02_ncu_list.log 2>&1; tail -1 $O/r02_ncu_list.log
T=128 timeout 900 ncu --set full --clock-control none --import-source on -k regex:ppo_epoch -c 1 -f -o $O/r02_ppo_epoch python scripts/ncu_fused.py > $O/r02_ncu_full.log 2>&1; tail -2 $O/r02_ncu_full.log
ncu -i $O/r02_ppo_epoch.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,smsp__inst_executed.sum,sm__inst_executed_pipe_tc.sum > $O/r02_ppo_epoch_raw.csv 2>/dev/null; tail -2 $O/r02_ppo_epoch_raw.csv | cut -c1-600
