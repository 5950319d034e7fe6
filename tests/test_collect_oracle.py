"""CPU: the collect-side oracle (oracle/collect.py) reproduces the fixtures minted from the unmodified reference
classes (tests/golden/make_golden_collect.py): act sampling with injected draws and the n-step assemblers."""
import numpy as np
import pytest
import torch

import gen_inputs as G
import make_golden_collect as MC
from helpers import load_golden
from oracle import collect as oc


def _tparams(p):
    return {k: torch.from_numpy(v) for k, v in p.items()}


@pytest.mark.parametrize("name,case_name", [("act_ppo_discrete", "ppo_discrete_small"), ("act_ppo_continuous", "ppo_continuous_small")])
def test_act_ppo_oracle_matches_reference(name, case_name):
    case, gold = G.PPO_CASES[case_name], load_golden(name)
    inp = MC.collect_inputs("act", case)
    params = _tparams(G.make_params(G.ppo_shapes(case), case["seed"]))
    a_train = oc.act_ppo(params, inp["state"], case["continuous"], True, u=inp["u"], eps=inp["eps"])
    a_eval = oc.act_ppo(params, inp["state"], case["continuous"], False)
    if case["continuous"]:
        np.testing.assert_allclose(a_train, gold["action_train"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(a_eval, gold["action_eval"], rtol=0, atol=1e-6)
    else:
        np.testing.assert_array_equal(a_train, gold["action_train"])
        np.testing.assert_array_equal(a_eval, gold["action_eval"])


@pytest.mark.parametrize("name,case_name,net", [("act_dqn", "dqn_small", "dqn"), ("act_ape_x", "ape_x_small", "dueling")])
def test_act_q_oracle_matches_reference(name, case_name, net):
    case, gold = G.Q_CASES[case_name], load_golden(name)
    inp = MC.collect_inputs("act", case)
    action, q_sel = oc.act_q(_tparams(G.q_params(case)), inp["state"], gold["eps_rows"], inp["u2"], net)
    np.testing.assert_array_equal(action, gold["action"])
    if "q" in gold:
        np.testing.assert_allclose(q_sel, gold["q"], rtol=1e-6, atol=1e-6)
    # both branches of the epsilon test are exercised
    rnd = inp["u2"][:, 0] < gold["eps_rows"]
    assert rnd.any() and (~rnd).any()


def test_act_rainbow_oracle_matches_reference():
    case, gold = G.Q_CASES["rainbow_small"], load_golden("act_rainbow")
    inp = MC.collect_inputs("act", case)
    noise = [(torch.from_numpy(a), torch.from_numpy(b)) for a, b in G.q_case_inputs(case)["noise"][0]]
    p = _tparams(G.q_params(case))
    a_train = oc.act_rainbow(p, inp["state"], case["A"], case["K"], case["v_min"], case["v_max"], noise)
    a_eval = oc.act_rainbow(p, inp["state"], case["A"], case["K"], case["v_min"], case["v_max"], None)
    np.testing.assert_array_equal(a_train, gold["action_train"])
    np.testing.assert_array_equal(a_eval, gold["action_eval"])


@pytest.mark.parametrize("name,case_name,apex", [("nstep_multistep", "multistep_small", False),
                                                 ("nstep_rainbow", "rainbow_small", False), ("nstep_ape_x", "ape_x_small", True)])
def test_nstep_oracle_matches_reference(name, case_name, apex):
    case, gold = G.Q_CASES[case_name], load_golden(name)
    inp = MC.collect_inputs("nstep", case)
    T, N = inp["state"].shape[:2]
    wins = [oc.NStepWindow(case["n_step"], apex, case["gamma"]) for _ in range(N)]
    emitted, first = {}, None
    for t in range(T):
        rows = []
        for i in range(N):
            tr = {k: inp[k][t, i:i + 1] for k in ("state", "action", "reward", "done", "next_state")}
            if apex:
                tr["q"] = inp["q"][t, i:i + 1]
            rows.append(wins[i].push(tr))
        if rows[0]:
            first = t if first is None else first
            for k in rows[0]:
                emitted.setdefault(k, []).append(np.concatenate([np.asarray(r[k]) for r in rows], axis=0))
    assert first == int(gold["first_emit"])
    assert inp["done"][:first + 1].any() and inp["done"].sum() >= 3      # windows straddle episode ends
    for k, v in emitted.items():
        got = np.stack(v)
        assert got.shape == gold[k].shape, k
        if k == "priority":
            np.testing.assert_allclose(got, gold[k], rtol=1e-12, atol=1e-12)
        else:
            np.testing.assert_array_equal(got, gold[k], err_msg=k)
