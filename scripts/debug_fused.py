import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import gen_inputs as G
import test_ppo_gpu as T
from helpers import load_golden
for name in ["ppo_continuous_small", "ppo_discrete_h512", "ppo_continuous_h512"]:
    case = G.PPO_CASES[name]
    gold = load_golden(name)
    for rep in range(3):
        a1, r1, _ = T._run_cuda(case, False, use_fused=True)
        a2, r2, _ = T._run_cuda(case, False, use_fused=False)
        print(name, rep, {k: (round(r1[k], 6), round(r2[k], 6), round(float(gold["result." + k]), 6)) for k in r1})
        d = max((a1.network.p[k] - a2.network.p[k]).abs().max().item() for k in a1.network.p)
        print("   max param diff fused vs multi-launch:", d)
