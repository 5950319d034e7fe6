#!/bin/bash
# 2-GPU call: two-rank exchange test (discrete + continuous networks), the fixed collector test, configs[4] at N=2
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_multigpu_gpu.py "tests/test_ac_gpu.py::test_replay_collector_runs_the_three_agents_on_pendulum" -q -m gpu > $O/r02_pytest_mg.txt 2>&1; echo "mg rc=$?"
tail -25 $O/r02_pytest_mg.txt | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 170 $TR --master-port 29531 bench.py --gpus 2 --config ppo_continuous --steps 2 --warmup 1 --no-e2e > $O/r02_bench_ppo_continuous_n2.json 2> $O/r02_bench_cont_n2.err; echo "cont rc=$?"
tail -c 1200 $O/r02_bench_ppo_continuous_n2.json; grep -v "^\[W\|^W0\|OMP_NUM\|^\*\*\*" $O/r02_bench_cont_n2.err | tail -4 | cut -c1-300
