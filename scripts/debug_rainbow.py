import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import gen_inputs as G
from helpers import q_oracle_inputs
from oracle import nets, dqn as odqn
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_dqn_gpu as T
case = G.Q_CASES["rainbow_small"]
params, tparams, batch, hp, optim, inp = q_oracle_inputs(case)
ref = odqn.dist_learn(params, tparams, batch, hp, optim)
agent = T._make(case)
dev = "cuda"
noise = [[(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b in layers] for layers in inp["noise"]]
x = torch.from_numpy(inp["state"]).to(dev)
lg = agent.network.forward(x, True, tag="t.", noise=noise[0])
rl = nets.rainbow_network(params, batch["state"], hp["noise"][0], case["A"], case["K"])
print("logits diff", (lg.cpu() - rl).abs().max().item())
agent2, prio = T._run(case)
print("KL diff", (agent2.network._buf("t.kl", (case["B"],)).cpu() - ref["KL"]).abs().max().item())
print("KL mine", agent2.network._buf("t.kl", (case["B"],)).cpu()[:6].tolist())
print("KL ref ", ref["KL"][:6].tolist())
print("done0", inp["done"][:6, 0, 0].tolist(), "w", inp["weights"][:4])
print("loss", agent2._stats[0].item(), ref["loss"])
