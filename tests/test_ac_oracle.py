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
