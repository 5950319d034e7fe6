"""Intra-step timeline of the persistent PPO kernel: clock64 stamps of the last step (debug trace, JB_FUSED_SKIP=256)."""
import sys, os, ctypes
os.environ["JB_FUSED_SKIP"] = "256"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jorldy_b200.core import Agent, Env
from jorldy_b200.core.collect import RolloutCollector
from jorldy_b200._lib import C

N, T, B = 4096, 32, int(os.environ.get("B", 256))
env = Env("cartpole", num_envs=N, seed=0)
agent = Agent("ppo", state_size=4, action_size=2, hidden_size=512, batch_size=B, n_step=T, n_epoch=1,
              optim_config={"name": "adam", "lr": 2.5e-4}, device="cuda", run_step=10**9, use_fused=True)
col = RolloutCollector(env, agent, use_cuda_graph=False); col.collect()
agent.learn_rollout(col.rollout); col.rollout.t = T
st = agent._st; fr = agent._fused[B]
n = N * T // B
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
agent._cursor.zero_(); e0.record(); fr.run(st, n); e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1)/n*1000:.1f} us/step over {n} steps")
tr = np.zeros((256, 48), np.int64)
C.jb_ppo_fused_trace(tr.ctypes.data_as(ctypes.c_void_p))
names = {0: "step start", 1: "P1 h1 generated", 2: "P1 panel landed", 3: "P1 mma", 4: "P1 reduce", 5: "P1 end", 6: "bar1",
         7: "row phase done", 8: "P3 job0 start", 9: "P3 job1 start", 10: "P3 job2 start", 11: "P3 job3+ start",
         12: "JB staged", 13: "JB dh2 gen", 14: "JB mma", 15: "JB reduce", 16: "JB end",
         17: "JA staged", 18: "JA dh2 gen", 19: "JA mma(last)", 20: "JA reduce(last)", 21: "JA end",
         22: "P3 jobs end", 28: "P1 stash issued", 29: "P1 stash landed", 31: "row: head outputs", 32: "row: math done",
         33: "JB dW1 staged", 34: "JB dW1 stored", 35: "P1 shuffles done", 36: "row: perm issued", 23: "norm partial + p/m/v issued", 24: "bar3", 25: "P5 fold", 26: "P5 end", 27: "bar5"}
ghz = 1.965
for cta in [0, 60, 147]:
    t = tr[cta]
    print(f"--- CTA {cta}")
    order = sorted([i for i in names if t[i] > 0], key=lambda i: t[i])
    prev = t[0]
    for i in order:
        print(f"  {names[i]:22s} +{(t[i]-prev)/ghz/1000:6.2f} us   @{(t[i]-t[0])/ghz/1000:6.2f}")
        prev = t[i]
# barrier waits: arrival spread
for a_, b_, nm in [(5, 6, "bar1"), (23, 24, "bar3"), (26, 27, "bar5")]:
    arr = tr[:148, a_] - tr[:148, 0]; dep = tr[:148, b_] - tr[:148, 0]
    print(f"{nm}: arrive min {arr.min()/ghz/1000:.2f} max {arr.max()/ghz/1000:.2f} (cta {arr.argmax()}) | depart-arrive min {(dep - arr).min()/ghz/1000:.2f} us")
