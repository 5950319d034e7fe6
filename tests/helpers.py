"""Shared helpers for oracle-vs-golden and CUDA-vs-oracle tests."""
import os

import numpy as np
import torch

import gen_inputs as G

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def ppo_oracle_inputs(case):
    inp = G.ppo_case_inputs(case)
    params = {k: torch.from_numpy(v) for k, v in G.make_params(G.ppo_shapes(case), case["seed"]).items()}
    batch = {
        "state": torch.from_numpy(inp["state"]),
        "next_state": torch.from_numpy(inp["next_state"]),
        "action": torch.from_numpy(inp["action"].astype(np.float32)),
        "reward": torch.from_numpy(inp["reward"].astype(np.float32)),
        "done": torch.from_numpy(inp["done"].astype(np.float32)),
    }
    hp = {"continuous": case["continuous"], "n_step": case["T"], "gamma": case["gamma"], "lambda": case["lam"],
          "standardize": case["standardize"], "batch_size": case["batch_size"], "n_epoch": case["n_epoch"],
          "eps_clip": case["eps_clip"], "vf_coef": case["vf_coef"], "ent_coef": case["ent_coef"],
          "clip_grad_norm": case["clip_grad_norm"]}
    return params, batch, hp, inp["perms"]


def check_against_golden(gold, params_after, result, pre=None, rtol=1e-5, atol=1e-6, stat_tol=1e-5):
    for k, v in gold.items():
        if k.startswith("param."):
            got = G.subsample(np.asarray(params_after[k[6:]]))
            np.testing.assert_allclose(got, v, rtol=rtol, atol=atol, err_msg=k)
        elif k.startswith("result."):
            np.testing.assert_allclose(result[k[7:]], float(v), rtol=stat_tol, atol=stat_tol, err_msg=k)
        elif k.startswith("pre.") and pre is not None and k[4:] in pre:
            got = G.subsample(np.asarray(pre[k[4:]], dtype=np.float32))
            np.testing.assert_allclose(got, v, rtol=rtol, atol=atol, err_msg=k)
