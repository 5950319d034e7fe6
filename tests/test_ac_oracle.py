"""CPU oracle of DDPG / TD3 / SAC (oracle/actor_critic.py) vs the golden files minted from the unmodified reference
classes (tests/golden/make_golden_ac.py): result dicts and post-step parameters of every network."""
import numpy as np
import pytest

import gen_inputs as G
from helpers import load_golden, run_ac_oracle


@pytest.mark.parametrize("name", list(G.AC_CASES.keys()))
def test_ac_oracle_matches_reference(name):
    case = G.AC_CASES[name]
    gold = load_golden(name)
    outs, nets = run_ac_oracle(case)
    for i, o in enumerate(outs):
        for k, v in o["result"].items():
            np.testing.assert_allclose(v, float(gold[f"result{i}.{k}"]), rtol=1e-5, atol=1e-6, err_msg=f"{k} (learn {i})")
    if case["agent"] == "td3" and case["num_learn"] % 2:
        assert float(gold["result0.actor_loss"]) == 0.0          # td3.py:129: the stale value when the actor is not updated
    checked = 0
    for k, v in gold.items():
        if k.startswith("param."):
            _, net, key = k.split(".", 2)
            np.testing.assert_allclose(G.subsample(nets[net][key].numpy()), v, rtol=1e-5, atol=1e-7, err_msg=k)
            checked += 1
    assert checked >= 16
    if case["agent"] == "sac":
        np.testing.assert_allclose(outs[-1]["log_alpha"].item(), float(gold["log_alpha"]), rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(outs[-1]["alpha"].item(), float(gold["alpha"]), rtol=1e-6)


def test_ou_noise_restates_the_reference_process():
    """agent/utils.py:8-26: X <- X + theta (mu - X) + sigma * randn(1), one draw shared by every action dimension."""
    from oracle.actor_critic import ou_step
    rs = np.random.RandomState(0)
    X = np.ones((1, 3), dtype=np.float32) * 0.0
    draws = rs.standard_normal(5)
    ref = X.copy()
    for n in draws:
        ref = ref + (1e-3 * (0.0 - ref) + 2e-3 * np.array([n]))
        X = ou_step(X, 0.0, 1e-3, 2e-3, n)
    np.testing.assert_array_equal(X, ref)
    assert X.dtype == np.float64 and np.all(X[0] == X[0, 0])


def _act_case(name):
    import importlib.util
    import os
    import torch
    spec = importlib.util.spec_from_file_location("make_golden_ac_act", os.path.join(os.path.dirname(G.__file__), "make_golden_ac_act.py"))
    mk = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mk)            # imports refimport (module-level import only; the reference itself is not needed)
    except Exception as e:                     # pragma: no cover
        pytest.skip(f"cannot import the fixture maker: {e}")
    case = G.AC_CASES[mk.CASES[name]]
    actor = {k: torch.from_numpy(v) for k, v in G.ac_params(case, "actor").items()}
    return mk, case, actor, mk.ac_act_inputs(name), load_golden(name)


def test_ddpg_act_oracle_matches_reference():
    """DDPG.act (ddpg.py:113-118) over 3 consecutive steps of 24 actors: float64 actions, OU state carried, ONE normal per step."""
    from oracle import actor_critic as oac
    mk, case, actor, inp, gold = _act_case("act_ddpg")
    X = [np.ones((1, case["A"]), np.float32) * mk.OU["mu"] for _ in range(mk.M)]
    for t in range(mk.T):
        for m in range(mk.M):
            a, X[m] = oac.act_ddpg(actor, inp["state"][t, m:m + 1], X[m], inp["scalar"][t, m], **mk.OU)
            np.testing.assert_allclose(a[0], gold["action"][t, m], rtol=1e-6, atol=1e-7)
    g = np.concatenate([oac.act_ddpg(actor, inp["state"][0, m:m + 1], None, None, training=False, **mk.OU)[0] for m in range(mk.M)])
    np.testing.assert_allclose(g, gold["greedy"], rtol=1e-6, atol=1e-7)


def test_td3_and_sac_act_oracle_match_reference():
    from oracle import actor_critic as oac
    mk, case, actor, inp, gold = _act_case("act_td3")
    for m in range(mk.M):
        a = oac.act_td3(actor, inp["state"][0, m:m + 1], inp["vector"][0, m], mk.TD3_STD)
        np.testing.assert_allclose(a[0], gold["action"][m], rtol=1e-6, atol=1e-7)
    assert np.abs(gold["action"]).max() <= 1.0 and (np.abs(gold["action"]) == 1.0).any()      # the clip is exercised
    mk, case, actor, inp, gold = _act_case("act_sac")
    for m in range(mk.M):
        a = oac.act_sac(actor, inp["state"][0, m:m + 1], inp["vector"][0, m:m + 1])
        np.testing.assert_allclose(a[0], gold["action"][m], rtol=1e-6, atol=1e-7)
    g = np.concatenate([oac.act_sac(actor, inp["state"][0, m:m + 1], None, training=False) for m in range(mk.M)])
    np.testing.assert_allclose(g, gold["greedy"], rtol=1e-6, atol=1e-7)
