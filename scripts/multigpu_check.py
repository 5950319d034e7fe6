"""2-GPU functional check (torchrun): PPO data-parallel learner (identical weights on every rank after
learn), Ape-X sharded PER (global IS normalisation), NCCL all-reduce inside the captured graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from jorldy_b200.core import Agent, Env, parallel
from jorldy_b200.core.collect import RolloutCollector, ReplayCollector
# ---- PPO dp ----
env = Env("cartpole", num_envs=512, seed=0, id=rank, device=dev)
agent = Agent("ppo", state_size=4, action_size=2, hidden_size=512, batch_size=256, n_step=32, n_epoch=2, device=dev,
              run_step=10**6, seed=100 + rank)
agent.rng_stream_base = rank << 32
parallel.attach(agent, world)
col = RolloutCollector(env, agent)
for it in range(3):
    res = agent.learn_rollout(col.collect())
flat = agent.network.flat.clone()
ref = flat.clone(); dist.broadcast(ref, src=0)
assert torch.equal(flat, ref), "weights diverged across ranks"
obs0 = env.obs.clone(); o0 = obs0.clone(); dist.broadcast(o0, src=0)
assert rank == 0 or not torch.equal(obs0, o0), "ranks should own different env shards"
print(f"rank {rank}: PPO dp ok", {k: round(v, 4) for k, v in res.items()}, flush=True)
# ---- Ape-X sharded PER ----
env2 = Env("cartpole", num_envs=64, seed=1, id=rank, device=dev)
ax = Agent("ape_x", state_size=4, action_size=2, hidden_size=128, network="dueling", buffer_size=8192, batch_size=64,
           start_train_step=64, target_update_period=50, run_step=10**6, n_step=3, num_workers=64 * world, device=dev,
           optim_config={"name": "rmsprop", "lr": 1e-4, "eps": 1.5e-7, "centered": True}, learn_period=4, seed=7)
parallel.attach(ax, world)
rc = ReplayCollector(env2, ax, update_period=8)
step = 0
for it in range(30):
    step, res = rc.run_round(step)
assert ax.num_learn > 0
f2 = ax.network.flat.clone(); r2 = f2.clone(); dist.broadcast(r2, src=0)
assert torch.equal(f2, r2), "Ape-X weights diverged"
tr, w, idx, stats = ax.memory.sample_device(0.5, 64)
wm = w.max().clone(); dist.all_reduce(wm, op=dist.ReduceOp.MAX)
assert abs(wm.item() - 1.0) < 1e-12, wm.item()
print(f"rank {rank}: Ape-X sharded PER ok, learns={ax.num_learn}, local items={ax.memory.size}, max w (global)={wm.item():.3f}", flush=True)
dist.barrier(); dist.destroy_process_group()
