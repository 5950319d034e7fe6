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


# ---- value-based agents --------------------------------------------------------------------------
def q_oracle_inputs(case):
    inp = G.q_case_inputs(case)
    params = {k: torch.from_numpy(v) for k, v in G.q_params(case).items()}
    tparams = {k: torch.from_numpy(v) for k, v in G.q_params(case, seed_offset=1000).items()}
    batch = {"state": torch.from_numpy(inp["state"]), "next_state": torch.from_numpy(inp["next_state"]),
             "action": torch.from_numpy(inp["action"].astype(np.float32)),
             "reward": torch.from_numpy(inp["reward"].astype(np.float32)),
             "done": torch.from_numpy(inp["done"].astype(np.float32)), "weights": inp["weights"]}
    ag = case["agent"]
    noise = None
    if inp["noise"] is not None:
        noise = [[(torch.from_numpy(a), torch.from_numpy(b)) for a, b in layers] for layers in inp["noise"]]
    hp = {"action_size": case["A"], "gamma": case["gamma"], "n_step": case.get("n_step", 1),
          "alpha": case.get("alpha", 0.0), "clip": case.get("clip"), "noise": noise,
          "net": {"discrete_q_network": "dqn", "dueling": "dueling", "noisy": "noisy", "rainbow": "rainbow"}[case["net"]],
          "double": ag in ("double", "per", "ape_x"), "loss": "wmse" if ag in ("per", "ape_x") else "huber",
          "order": {"dqn": "dqn", "dueling": "dqn", "noisy": "dqn", "double": "double", "per": "double",
                    "multistep": "nstep", "ape_x": "nstep"}.get(ag, "dqn")}
    if ag in ("c51", "rainbow"):
        hp.update(variant=ag, num_support=case["K"], v_min=case["v_min"], v_max=case["v_max"])
    if ag == "noisy" and noise is not None:
        hp["noise"] = [noise[0], None, noise[2]]
    optim = case.get("optim", {"name": "adam", "lr": case["lr"]})
    return params, tparams, batch, hp, optim, inp


def run_q_oracle(case):
    from oracle import dqn as odqn
    params, tparams, batch, hp, optim, inp = q_oracle_inputs(case)
    if case["agent"] in ("c51", "rainbow"):
        return odqn.dist_learn(params, tparams, batch, hp, optim), inp
    return odqn.td_learn(params, tparams, batch, hp, optim), inp


# ---- continuous off-policy family ----------------------------------------------------------------------------------------
def ac_oracle_inputs(case):
    inp = G.ac_case_inputs(case)
    nets = {n: {k: torch.from_numpy(v) for k, v in G.ac_params(case, n).items()} for n in G.AC_NETS}
    batch = {"state": torch.from_numpy(inp["state"]), "next_state": torch.from_numpy(inp["next_state"]),
             "action": torch.from_numpy(inp["action"]), "reward": torch.from_numpy(inp["reward"].astype(np.float32)),
             "done": torch.from_numpy(inp["done"].astype(np.float32))}
    noise = [{k: torch.from_numpy(v) for k, v in nz.items()} for nz in inp["noise"]]
    return nets, batch, noise


def run_ac_oracle(case):
    """Runs the case's learn() calls on the CPU oracle; returns (list of per-learn outputs, final network dict)."""
    from oracle import actor_critic as oac
    nets, batch, noise = ac_oracle_inputs(case)
    ag, outs, opt_state = case["agent"], [], None
    hp = {k: case[k] for k in ("gamma", "tau", "actor_lr", "critic_lr", "alpha_lr")}
    if ag == "sac":
        log_alpha = torch.zeros(1) if case["dynamic_alpha"] else torch.tensor(-2.0)
        alpha = log_alpha.exp()
        hp.update(use_dynamic_alpha=case["dynamic_alpha"], target_entropy=-case["A"])
    for i, nz in enumerate(noise):
        if ag == "ddpg":
            o = oac.ddpg_learn(nets["actor"], nets["critic1"], nets["target_actor"], nets["target_critic1"], batch, hp, opt_state)
            nets.update(actor=o["actor"], critic1=o["critic"])
        elif ag == "td3":
            hp.update(update_delay=2, target_noise_std=0.2, target_noise_clip=0.5)
            o = oac.td3_learn(nets["actor"], nets["critic1"], nets["critic2"], nets["target_actor"], nets["target_critic1"],
                              nets["target_critic2"], batch, hp, nz["target"], case["num_learn"] + i, opt_state)
            nets.update({k: o[k] for k in G.AC_NETS})
        else:
            o = oac.sac_learn(nets["actor"], nets["critic1"], nets["critic2"], nets["target_critic1"], nets["target_critic2"],
                              log_alpha, alpha, batch, hp, nz["next"], nz["actor"], opt_state)
            nets.update(actor=o["actor"], critic1=o["critic1"], critic2=o["critic2"])
            log_alpha, alpha = o["log_alpha"], o["alpha"]
        opt_state = o["opt_state"]
        outs.append(o)
    return outs, nets
