#!/bin/bash
# last 1-GPU call of round 2: the whole GPU suite with the CUDA-graph learn() of DDPG/TD3/SAC, then the SAC bench line again
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python -m pytest tests -m gpu -q > $O/r02_pytest_final.txt 2>&1; echo "pytest rc=$?"
grep "^FAILED\|^ERROR\|passed\|failed\|^E  " $O/r02_pytest_final.txt | head -30 | cut -c1-260
timeout 100 python bench.py --config sac_hopper --steps 5 --warmup 3 > $O/r02_bench_sac_hopper_graph_n1.json 2> $O/r02_bench_sac.err; echo "sac rc=$?"
tail -c 1500 $O/r02_bench_sac_hopper_graph_n1.json; grep -v "^\[W\|^W0" $O/r02_bench_sac.err | tail -4 | cut -c1-300
