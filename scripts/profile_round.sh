#!/bin/bash
# Round-end measurement bundle (run under gpurun on ONE B200): tests, smoke, bench, ncu launch list of one
# bench-shaped PPO iteration, ncu --set full capture of the dominant kernel.  Outputs -> gpurun_out/.
set -u
R=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$R.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
kill $SMI
tail -c 2500 gpurun_out/bench_$R.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$R.json 2>/dev/null
tail -c 600 gpurun_out/bench_ref_$R.json
# launch list of one full iteration at bench shape (fused kernel = 1 launch per epoch)
N_ENVS=4096 T=128 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench_$R.csv python scripts/ncu_ppo.py > gpurun_out/ncu_list_$R.log 2>&1
tail -1 gpurun_out/ncu_list_$R.log
# full capture of the dominant kernel at bench shape (2048 steps/launch)
T=128 timeout 900 ncu --set full --clock-control none --import-source on -k regex:ppo_epoch -c 1 -o gpurun_out/ppo_epoch_$R python scripts/ncu_fused.py > gpurun_out/ncu_full_$R.log 2>&1
tail -2 gpurun_out/ncu_full_$R.log
# per-kernel DRAM traffic + headline numbers of that capture
ncu -i gpurun_out/ppo_epoch_$R.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,smsp__inst_executed.sum > gpurun_out/ppo_epoch_raw_$R.csv 2>/dev/null
tail -3 gpurun_out/ppo_epoch_raw_$R.csv
timeout 120 python scripts/perf_trace.py > gpurun_out/trace_$R.txt 2>&1; head -3 gpurun_out/trace_$R.txt
timeout 300 python scripts/perf_ppo.py 2>&1 | tail -6
