"""One short launch of the persistent PPO kernel for `ncu --set full --import-source on -k regex:ppo_epoch -c 1`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jorldy_b200.core import Agent, Env
from jorldy_b200.core.collect import RolloutCollector

N, T, B = 4096, int(os.environ.get("T", 8)), 256
env = Env("cartpole", num_envs=N, seed=0)
agent = Agent("ppo", state_size=4, action_size=2, hidden_size=512, batch_size=B, n_step=T, n_epoch=1,
              optim_config={"name": "adam", "lr": 2.5e-4}, device="cuda", run_step=10**9, use_fused=True)
col = RolloutCollector(env, agent, use_cuda_graph=False); col.collect()
agent.learn_rollout(col.rollout)
torch.cuda.synchronize()
print("done", agent._fused.keys())
