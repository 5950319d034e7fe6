#!/bin/bash
# 1-GPU measurement bundle, second pass (after the tensor-core engine became opt-in): tests, timelines of both engines, the
# headline bench line + reference arm, configs[4] at N=1, ncu full capture + launch list of the dominant (FFMA-engine) kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_pytest.txt 2>&1; echo "pytest rc=$?"
tail -12 $O/r02_pytest.txt | cut -c1-400
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | cut -c1-400
JB_FUSED_TC=1 timeout 200 python scripts/perf_trace.py > $O/r02_trace_tc.txt 2>&1; echo "trace rc=$?"; head -30 $O/r02_trace_tc.txt
JB_FUSED_TC=0 timeout 200 python scripts/perf_trace.py > $O/r02_trace_ffma.txt 2>&1; head -36 $O/r02_trace_ffma.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/r02_clocks_bench.csv &
SMI=$!
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?"
kill $SMI
tail -c 3000 $O/r02_bench_n1.json; tail -3 $O/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_n1_reference_arm.json 2>/dev/null; tail -c 400 $O/r02_bench_n1_reference_arm.json
timeout 1500 python bench.py --config ppo_continuous --steps 2 --warmup 1 > $O/r02_bench_ppo_continuous_n1.json 2> $O/r02_bench_cont.err; echo "cont rc=$?"
tail -c 2000 $O/r02_bench_ppo_continuous_n1.json; tail -4 $O/r02_bench_cont.err
N_ENVS=4096 T=128 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches_bench.csv python scripts/ncu_ppo.py > $O/r02_ncu_list.log 2>&1; tail -1 $O/r02_ncu_list.log
T=128 timeout 900 ncu --set full --clock-control none --import-source on -k regex:ppo_epoch -c 1 -f -o $O/r02_ppo_epoch_ffma python scripts/ncu_fused.py > $O/r02_ncu_full.log 2>&1; tail -2 $O/r02_ncu_full.log
ncu -i $O/r02_ppo_epoch_ffma.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,smsp__inst_executed.sum > $O/r02_ppo_epoch_ffma_raw.csv 2>/dev/null; tail -2 $O/r02_ppo_epoch_ffma_raw.csv | cut -c1-600
