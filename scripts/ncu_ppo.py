"""Short PPO run without CUDA graphs for `ncu` launch lists / full captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jorldy_b200.core import Agent, Env
from jorldy_b200.core.collect import RolloutCollector
N = int(os.environ.get("N_ENVS", 4096)); T = int(os.environ.get("T", 8)); B = 256
env = Env("cartpole", num_envs=N, seed=0)
agent = Agent("ppo", state_size=4, action_size=2, hidden_size=512, batch_size=B, n_step=T, n_epoch=int(os.environ.get("EPOCHS", 3)),
              optim_config={"name": "adam", "lr": 2.5e-4}, device="cuda", run_step=10**9, use_cuda_graph=False,
              use_fused=os.environ.get("FUSED", "1") == "1")
col = RolloutCollector(env, agent, use_cuda_graph=False)
for it in range(int(os.environ.get("ITERS", 1))):
    ro = col.collect()
    agent.learn_rollout(ro)
torch.cuda.synchronize()
print("done")
