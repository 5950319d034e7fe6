"""Multi-GPU functional check (torchrun, >= 2 ranks): PPO data-parallel learner with the in-kernel gradient exchange
(equal to the NCCL path to round-off, global critic means, identical weights on every rank after learn), Ape-X sharded
PER (global IS normalisation)."""
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
NE, NEP = int(os.environ.get("JB_NENV", 512)), int(os.environ.get("JB_NEPOCH", 2))
env = Env("cartpole", num_envs=NE, seed=0, id=rank, device=dev)


def trio(**kw):
    """The same learner three ways: persistent kernel with the in-kernel exchange, CUDA graphs + NCCL, eager + NCCL."""
    mk = lambda **k2: Agent("ppo", state_size=4, action_size=2, hidden_size=512, batch_size=256, n_step=32, n_epoch=NEP, device=dev,
                            run_step=10**6, seed=100 + rank, optim_config={"name": "adam", "lr": 2.5e-4}, **kw, **k2)
    agent = mk()
    agent.rng_stream_base = rank << 32
    parallel.attach(agent, world)
    ref_agent = mk(use_fused=False)
    parallel.attach(ref_agent, world)
    ref_agent.p2p = None
    ref_agent.network.load_state_dict(agent.network.state_dict())
    eag = mk(use_fused=False, use_cuda_graph=False)
    parallel.attach(eag, world); eag.p2p = None
    eag.network.load_state_dict(agent.network.state_dict())
    return agent, ref_agent, eag


def learn3(agents, ro):
    out = []
    for ag in agents:
        ro.t = ag.n_step
        torch.manual_seed(1234)
        out.append(ag.learn_rollout(ro))
        torch.cuda.synchronize()
    return out


n_steps_run = NEP * (NE * 32 // 256)
# (1) exchange arithmetic: with the clip range wide open both critic candidates coincide (v_clip == v), so the fused kernel's
# GLOBAL critic means and the NCCL path's per-rank ones select the same gradient and the three learners must agree to
# fp32 round-off.  (Adam normalises every coordinate's step to ~lr, so coordinates whose gradient is pure round-off random-walk
# apart between ANY two summation orders: the element-wise comparison is asserted for short runs only.)
agent, ref_agent, eag = trio(epsilon_clip=1e9)
print(f"rank {rank}: in-kernel gradient exchange {'ON' if agent.p2p else 'off (NCCL all-reduce)'}", flush=True)
col = RolloutCollector(env, agent)
ro = col.collect()
res, res_ref, res_eag = learn3((agent, ref_agent, eag), ro)
d = (agent.network.flat - ref_agent.network.flat).abs().max().item()
print(f"rank {rank}: [no clipping] graph vs eager max|dW| = {(ref_agent.network.flat - eag.network.flat).abs().max().item():.3e}; "
      f"fused(p2p={bool(agent.p2p)}) vs graph+NCCL = {d:.3e} after {n_steps_run} steps", flush=True)
if n_steps_run <= 8:
    assert d < 5e-5, d
for k in res:           # (critic_loss is REPORTED globally by the fused kernel, per rank by the NCCL path: compared in (2))
    if k != "critic_loss":
        assert abs(res[k] - res_ref[k]) < 5e-3 * max(1.0, abs(res_ref[k])), (k, res[k], res_ref[k])
# (2) the reference's clip range: the fused kernel evaluates critic_loss = max(mean, mean) over the GLOBAL minibatch (identical
# on every rank, equal to the reference's semantics); the NCCL path evaluates it per rank
agent, ref_agent, eag = trio()
col = RolloutCollector(env, agent)
ro = col.collect()
res, res_ref, _ = learn3((agent, ref_agent, eag), ro)
cl = torch.tensor([res["critic_loss"], res_ref["critic_loss"]], dtype=torch.float64, device=dev)
lo, hi = cl.clone(), cl.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert lo[0].item() == hi[0].item(), "the fused kernel's critic loss must be the same global number on every rank"
print(f"rank {rank}: [clip 0.1] critic_loss fused (global) = {res['critic_loss']:.6f}; NCCL path (per rank) = {res_ref['critic_loss']:.6f} "
      f"in [{lo[1].item():.6f}, {hi[1].item():.6f}]", {k: round(v, 4) for k, v in res.items()}, flush=True)
for k in ("actor_loss", "entropy_loss", "mean_ret"):
    assert abs(res[k] - res_ref[k]) < 2e-2 * max(1.0, abs(res_ref[k])), (k, res[k], res_ref[k])
for it in range(3):
    res = agent.learn_rollout(col.collect())
flat = agent.network.flat.clone()
ref = flat.clone(); dist.broadcast(ref, src=0)
assert torch.equal(flat, ref), "weights diverged across ranks"
obs0 = env.obs.clone(); o0 = obs0.clone(); dist.broadcast(o0, src=0)
assert rank == 0 or not torch.equal(obs0, o0), "ranks should own different env shards"
print(f"rank {rank}: PPO dp ok (fused path: {bool(agent._fused)})", {k: round(v, 4) for k, v in res.items()}, flush=True)
# ---- PPO dp, continuous policy (obs 11 / act 3: a flat parameter buffer that is NOT a multiple of 8 floats) ----
envc = Env("hopper", num_envs=64, seed=2, id=rank, device=dev)
mkc = lambda **k2: Agent("ppo", state_size=11, action_size=3, hidden_size=512, batch_size=256, n_step=16, n_epoch=1, device=dev,
                         network="continuous_policy_value", run_step=10**6, seed=300 + rank, epsilon_clip=1e9,
                         optim_config={"name": "adam", "lr": 3e-4}, **k2)
ac, rc_ = mkc(), mkc(use_fused=False)
ac.rng_stream_base = rank << 32
parallel.attach(ac, world)
parallel.attach(rc_, world); rc_.p2p = None
rc_.network.load_state_dict(ac.network.state_dict())
assert (ac.p2p is not None) == (agent.p2p is not None), "the continuous network must get the same exchange path as the discrete one"
roc = RolloutCollector(envc, ac).collect()
resc, resr = learn3((ac, rc_), roc)
dc = (ac.network.flat - rc_.network.flat).abs().max().item()
fc = ac.network.flat.clone(); f0 = fc.clone(); dist.broadcast(f0, src=0)
assert torch.equal(fc, f0), "continuous PPO weights diverged across ranks"
assert dc < 5e-5, dc
print(f"rank {rank}: PPO continuous dp ok (p2p={bool(ac.p2p)}, num_flat % 8 = {ac.network.num_flat % 8}): fused vs graph+NCCL max|dW| = {dc:.3e} "
      f"after {64 * 16 // 256} steps", flush=True)
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
dist.barrier(); torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
