"""Launches every kernel of the hot path ONCE OR TWICE at its BASELINE size, for an ncu capture of all of them
(profiles/r02_kernels.md).  Run under ncu on one B200:

  ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,\
launch__registers_per_thread,launch__grid_size,launch__block_size \
      --profile-from-start off -o gpurun_out/r02_kernels python scripts/ncu_kernels.py

Every section runs once UNPROFILED (allocates workspaces; ncu flushes caches per replay anyway), then once between
cudaProfilerStart/Stop, so that only OUR launches (and the few torch fills next to them) are captured; the section
name is printed in order, and an NVTX range carries it into the report.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jorldy_b200.core import Agent, Env  # noqa: E402
from jorldy_b200.core.buffer import PERBuffer  # noqa: E402
from jorldy_b200.core.dev import C, ptr, stream_ptr  # noqa: E402

dev = torch.device("cuda")
ONLY = set(os.environ.get("SECTIONS", "").split(",")) - {""}


class section:
    def __init__(self, name):
        self.name = name

    def __call__(self, fn):
        if ONLY and self.name not in ONLY:
            return fn
        fn()                                      # warm-up / allocation pass, not captured
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("jbsec_" + self.name)
        torch.cuda.cudart().cudaProfilerStart()
        fn()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        torch.cuda.nvtx.range_pop()
        print("section", self.name, "done", flush=True)
        return fn


# ---- env kernels -----------------------------------------------------------------------------------------------------
for n in (4096, 1 << 20):
    env = Env("cartpole", num_envs=n, seed=0, device=dev)
    env.reset_device()
    act = torch.randint(0, 2, (n,), device=dev)

    @section(f"cartpole_step_n{n}")
    def _():
        env.step_device(act)

fenv = Env("breakout", num_envs=1024, seed=0, device=dev)
fenv.reset_device()


@section("frames_step_n1024")
def _():
    fenv.step_device(None)


# ---- GAE -------------------------------------------------------------------------------------------------------------
for (N, T) in ((4096, 128), (8192, 2048)):
    r = torch.rand(N * T, device=dev); d = (torch.rand(N * T, device=dev) < 0.01).float()
    v = torch.randn(N * T, device=dev); lv = torch.randn(N, device=dev)
    adv = torch.empty(N * T, device=dev); ret = torch.empty(N * T, device=dev)

    @section(f"gae_{N}x{T}")
    def _():
        C.jb_gae(ptr(r), ptr(d), ptr(v), 0, ptr(lv), N, T, 0.99, 0.95, 1, ptr(adv), ptr(ret), stream_ptr())
    del r, d, v, adv, ret

# ---- PPO: act / pre-pass at 4096 rows, one multi-launch minibatch step at B = 256, persistent kernel ----------------
penv = Env("cartpole", num_envs=4096, seed=0, device=dev)
penv.reset_device()
pa = Agent("ppo", state_size=4, action_size=2, hidden_size=512, batch_size=256, n_step=128, n_epoch=1,
           optim_config={"name": "adam", "lr": 2.5e-4}, device=dev, run_step=10 ** 9, use_cuda_graph=False, use_fused=False)


@section("ppo_act_4096")
def _():
    pa.act_device(penv.obs, True)


NT = 4096 * 128
st_state = torch.randn(NT, 4, device=dev) * 0.1
st = {"state": st_state, "action": torch.randint(0, 2, (NT,), device=dev, dtype=torch.int32),
      "adv": torch.randn(NT, device=dev), "ret": torch.randn(NT, device=dev), "value": torch.randn(NT, device=dev),
      "logp_old": torch.full((NT, 1), -0.69, device=dev), "perm": torch.randperm(NT, device=dev, dtype=torch.int32)}
out = torch.empty(16384, 3, device=dev)


@section("ppo_prepass_16384")
def _():
    pa.network.forward_rows(st_state[:16384], out)
    C.jb_ppo_prepass_discrete(ptr(out), ptr(st["action"]), 16384, 2, 3, ptr(st["value"]), ptr(st["logp_old"]), stream_ptr())


@section("ppo_minibatch_step_B256")
def _():
    pa._minibatch_step(st, st["perm"][:256], 256)


pf = Agent("ppo", state_size=4, action_size=2, hidden_size=512, batch_size=256, n_step=128, n_epoch=1,
           optim_config={"name": "adam", "lr": 2.5e-4}, device=dev, run_step=10 ** 9)
from jorldy_b200.core.agent import ppo_fused  # noqa: E402
runner = ppo_fused.FusedRunner(pf, 256)
pf._acc.zero_()


@section("ppo_epoch_kernel_256steps")
def _():
    pf._cursor.zero_()
    runner.run(st, 256)


# ---- PER sum-tree at 1 M slots ---------------------------------------------------------------------------------------
per = PERBuffer(1 << 20, 1e-3, device=dev)
per.buffer_counter = 1 << 20
per._tree[per.first_leaf_index:] = torch.rand(1 << 20, dtype=torch.float64, device=dev) + 0.01
C.jb_per_rebuild(ptr(per._tree), per.buffer_size, stream_ptr())
for B in (32, 512):
    idx = torch.randint(per.first_leaf_index, per.tree_size, (B,), device=dev)
    newp = torch.rand(B, dtype=torch.float64, device=dev)
    o_idx = torch.empty(B, dtype=torch.int64, device=dev); o_w = torch.empty(B, dtype=torch.float64, device=dev)
    o_p = torch.empty(B, dtype=torch.float64, device=dev); o_s = torch.empty(4, dtype=torch.float64, device=dev)

    @section(f"per_update_1M_B{B}")
    def _():
        per.update_priorities(idx, newp)

    @section(f"per_sample_1M_B{B}")
    def _():
        C.jb_per_sample(ptr(per._tree), per.buffer_size, per.buffer_counter, B, 0.4, 1e-3, 0, 0, 0, 1, 0, 0,
                        ptr(o_idx), ptr(o_w), ptr(o_p), ptr(o_s), 1, stream_ptr())

# ---- Rainbow (CNN, A = 4, K = 51, B = 32) and Ape-X (dueling CNN, B = 512) learn() ----------------------------------


def _fill(agent, n_step, B):
    batch = {"state": torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8, device=dev),
             "next_state": torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8, device=dev),
             "action": torch.randint(0, 4, (B, 1), device=dev),
             "reward": torch.randn(B, n_step, 1, device=dev).sign() * (torch.rand(B, n_step, 1, device=dev) < 0.1),
             "done": (torch.rand(B, n_step, 1, device=dev) < 0.01).float()}
    return batch


rb = Agent("rainbow", state_size=[4, 84, 84], action_size=4, hidden_size=512, head="cnn", buffer_size=4096, batch_size=32,
           n_step=3, alpha=0.5, beta=0.4, v_min=-1, v_max=10, num_support=51, device=dev, run_step=10 ** 7,
           optim_config={"name": "adam", "lr": 6.25e-5})
rbatch = _fill(rb, 3, 32)
rw = torch.rand(32, dtype=torch.float64, device=dev)


@section("rainbow_learn_cnn_B32")
def _():
    rb._dist_learn(rbatch, rw, 1, [None, None, None])


@section("rainbow_act_cnn_n64")
def _():
    rb.batch_size, rb.start_train_step = 0, 0
    rb.act_device(rbatch["state"].repeat(2, 1, 1, 1), True)
    rb.batch_size = 32


@section("target_copy_3M")
def _():
    rb.update_target()


ax = Agent("ape_x", state_size=[4, 84, 84], action_size=4, hidden_size=512, head="cnn", network="dueling", buffer_size=4096,
           batch_size=512, n_step=3, alpha=0.6, clip_grad_norm=40.0, num_workers=256, device=dev, run_step=10 ** 7,
           optim_config={"name": "rmsprop", "lr": 6.25e-5, "eps": 1.5e-7, "centered": True})
abatch = _fill(ax, 3, 512)
aw = torch.rand(512, dtype=torch.float64, device=dev)


@section("apex_learn_cnn_B512")
def _():
    ax._learn_batch(abatch, aw)


print("all sections done")
