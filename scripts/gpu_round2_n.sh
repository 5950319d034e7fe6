#!/bin/bash
# final 1-GPU call of round 2: whole GPU suite + smoke(), checkpoint fixtures of DDPG/TD3/SAC, the SAC (8f-4) bench line
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 500 python -m pytest tests -m gpu -q > $O/r02_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 $O/r02_pytest.txt | cut -c1-300
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-300
timeout 120 python scripts/make_ckpt_fixtures.py --ac-only 2>&1 | tail -4 | cut -c1-200
timeout 300 python bench.py --config sac_hopper --steps 5 --warmup 3 > $O/r02_bench_sac_hopper_n1.json 2> $O/r02_bench_sac.err; echo "sac rc=$?"
tail -c 2500 $O/r02_bench_sac_hopper_n1.json; grep -v "^\[W\|^W0" $O/r02_bench_sac.err | tail -6 | cut -c1-300
