"""Exploratory timing breakdown of the PPO CartPole pipeline (not the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jorldy_b200.core import Agent, Env
from jorldy_b200.core.collect import RolloutCollector

N = int(os.environ.get("N_ENVS", 4096)); T = 128; B = int(os.environ.get("BATCH", 256))
env = Env("cartpole", num_envs=N, seed=0)
agent = Agent("ppo", state_size=4, action_size=2, hidden_size=512, batch_size=B, n_step=T, n_epoch=3,
              optim_config={"name": "adam", "lr": 2.5e-4}, device="cuda", run_step=10**9)
col = RolloutCollector(env, agent)

def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

t_col = timed(col.collect)
ro = col.collect()
t_learn = timed(lambda: agent.learn_rollout(col.rollout), n=2)
print(f"collect {t_col:.2f} ms  ({N*T/t_col/1e3:.2f} M env-steps/s alone)")
print(f"learn   {t_learn:.2f} ms  ({N*T*3/t_learn/1e3:.2f} M learner transitions/s alone)")
print(f"total   {t_col+t_learn:.2f} ms -> {N*T/(t_col+t_learn)/1e3:.3f} M env-steps/s")
print("episodes finished:", env.stats.tolist())
# per-minibatch step time
st = agent._st
agent._cursor.zero_()     # learn() left the cursor at the end of the permutation
g = agent._graph_for(st, B)
agent._cursor.zero_()
t_g = timed(lambda: (agent._cursor.zero_(), g.replay()), n=20)
print(f"minibatch step (graph of 16): {t_g/16*1000:.1f} us/step")
# individual kernels
from jorldy_b200._lib import C
from jorldy_b200.core.dev import ptr, stream_ptr
net = agent.network
x = torch.randn(B, 512, device="cuda"); y = torch.empty(B, 512, device="cuda"); dw = torch.empty(512, 512, device="cuda"); db = torch.empty(512, device="cuda")
def k_fwd(): C.jb_linear_fwd(ptr(x), ptr(net.p["l.weight"]), ptr(net.p["l.bias"]), ptr(y), B, 512, 512, 1, stream_ptr())
def k_dx(): C.jb_linear_bwd_dx(ptr(x), ptr(net.p["l.weight"]), ptr(y), B, 512, 512, ptr(x), stream_ptr())
def k_dw(): C.jb_linear_bwd_dw(ptr(x), ptr(y), ptr(dw), ptr(db), B, 512, 512, stream_ptr())
def k_adam(): agent.optimizer.step(max_norm=1.0)
for name, fn in [("gemm fwd", k_fwd), ("gemm dx", k_dx), ("gemm dw", k_dw), ("sumsq+adam", k_adam)]:
    def many():
        for _ in range(50): fn()
    print(f"{name}: {timed(many, n=3)/50*1000:.2f} us")
xb = torch.randn(16384, 4, device="cuda"); ob = torch.empty(16384, 3, device="cuda")
t = timed(lambda: net.forward_rows(xb, ob), n=5)
print(f"forward_rows 16384: {t*1000:.1f} us  ({2*16384*(4*512+512*512+512*3)/t/1e9:.2f} TFLOP/s)")
