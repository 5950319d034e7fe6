"""Pins the CPU oracle (oracle/) to outputs of the UNMODIFIED reference (tests/golden/*.npz, minted
by tests/golden/make_golden.py in the build container).  CPU only."""
import numpy as np
import pytest
import torch

import gen_inputs as G
from helpers import check_against_golden, load_golden, ppo_oracle_inputs
from oracle import ppo as oppo


@pytest.mark.parametrize("name", list(G.PPO_CASES.keys()))
def test_ppo_oracle_matches_reference(name):
    case = G.PPO_CASES[name]
    torch.manual_seed(0)
    params, batch, hp, perms = ppo_oracle_inputs(case)
    out = oppo.learn(params, batch, hp, perms, lr=case["lr"])
    gold = load_golden(name)
    pre = {"value": out["value"], "next_value": out["next_value"], "adv": out["adv"], "ret": out["ret"],
           "log_prob_old": out["log_prob_old"]}
    check_against_golden(gold, {k: v.numpy() for k, v in out["params"].items()}, out["result"],
                         {k: v.numpy() for k, v in pre.items()}, rtol=1e-6, atol=1e-7, stat_tol=1e-6)
