"""CUDA collect side vs the reference-minted fixtures and the CPU oracle (pytest -m gpu).

Covers SURVEY.md §8 rows L1 (collect loop), D1 / X1 / R1 / G2 (act sampling with injected randomness),
D5 / R2 / X2 (n-step assemblers).  Bars: actions, indices, stored flags — bit-exact; float actions and Ape-X
actor priorities — 1e-6 absolute (fp32 tanh / exp differ from torch-CPU by an ulp).
"""
import numpy as np
import pytest
import torch

import gen_inputs as G
import make_golden_collect as MC
from helpers import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ppo_agent(case, **kw):
    from jorldy_b200.core import Agent
    net = "continuous_policy_value" if case["continuous"] else "discrete_policy_value"
    agent = Agent("ppo", state_size=case["D"], action_size=case["A"], hidden_size=case["H"], network=net,
                  optim_config={"name": "adam", "lr": case["lr"]}, run_step=1000, lr_decay=False, device=DEV,
                  batch_size=case["batch_size"], n_step=case["T"], n_epoch=case["n_epoch"], **kw)
    agent.network.load_state_dict({k: torch.from_numpy(v) for k, v in G.make_params(G.ppo_shapes(case), case["seed"]).items()})
    return agent


def _q_agent(case, **extra):
    from test_dqn_gpu import _make
    agent = _make(case)
    for k, v in extra.items():
        setattr(agent, k, v)
    return agent


# ------------------------------------------------------------------------------------------------ act (G2, D1, X1, R1)
def test_ppo_act_discrete_matches_reference():
    case, gold = G.PPO_CASES["ppo_discrete_small"], load_golden("act_ppo_discrete")
    inp = MC.collect_inputs("act", case)
    agent = _ppo_agent(case)
    s = torch.from_numpy(inp["state"]).to(DEV)
    a = agent.act_device(s, True, noise=torch.from_numpy(inp["u"]).to(DEV)).cpu().numpy()
    np.testing.assert_array_equal(a.reshape(-1, 1), gold["action_train"])
    a = agent.act_device(s, False).cpu().numpy()
    np.testing.assert_array_equal(a.reshape(-1, 1), gold["action_eval"])
    # the numpy-facing plugin call returns the reference's shape / dtype
    out = agent.act(inp["state"], training=False)["action"]
    assert out.shape == gold["action_eval"].shape and out.dtype == np.int64
    np.testing.assert_array_equal(out, gold["action_eval"])


def test_ppo_act_continuous_matches_reference():
    case, gold = G.PPO_CASES["ppo_continuous_small"], load_golden("act_ppo_continuous")
    inp = MC.collect_inputs("act", case)
    agent = _ppo_agent(case)
    s = torch.from_numpy(inp["state"]).to(DEV)
    a = agent.act_device(s, True, noise=torch.from_numpy(inp["eps"]).to(DEV)).cpu().numpy()
    np.testing.assert_allclose(a, gold["action_train"], rtol=0, atol=1e-6)
    a = agent.act_device(s, False).cpu().numpy()
    np.testing.assert_allclose(a, gold["action_eval"], rtol=0, atol=1e-6)


def test_ppo_act_sampling_law():
    """Without injection the Philox draws realise the policy's probabilities (10^5 rows, 4 sigma)."""
    case = G.PPO_CASES["ppo_discrete_small"]
    agent = _ppo_agent(case)
    s = torch.from_numpy(MC.collect_inputs("act", case)["state"][:1]).to(DEV).repeat(100000, 1).contiguous()
    a = agent.act_device(s, True).cpu().numpy()
    pi = load_golden("act_ppo_discrete")["pi"][0]
    freq = np.bincount(a, minlength=2) / a.size
    assert abs(freq[1] - pi[1]) < 4 * np.sqrt(pi[0] * pi[1] / a.size), (freq, pi)


def test_q_act_matches_reference_dqn():
    case, gold = G.Q_CASES["dqn_small"], load_golden("act_dqn")
    inp = MC.collect_inputs("act", case)
    agent = _q_agent(case, epsilon=MC.EPS_DQN)
    s = torch.from_numpy(inp["state"]).to(DEV)
    action, q_sel = agent.act_device(s, True, noise=torch.from_numpy(inp["u2"]).to(DEV))
    np.testing.assert_array_equal(action.cpu().numpy().reshape(-1, 1), gold["action"])
    # q_sel is the Q value of the action taken (used by Ape-X's actor-side priority)
    q = agent._q_values(s, True)
    np.testing.assert_array_equal(q_sel.cpu().numpy(), q.gather(1, action.view(-1, 1)).view(-1).cpu().numpy())


def test_q_act_matches_reference_ape_x():
    case, gold = G.Q_CASES["ape_x_small"], load_golden("act_ape_x")
    inp = MC.collect_inputs("act", case)
    agent = _q_agent(case, epsilon=0.4, num_workers=MC.M_ACT)
    agent.set_actor_epsilons(MC.M_ACT, total=MC.M_ACT)            # ape_x.py:166-172 for every row
    np.testing.assert_allclose(agent._eps_rows.cpu().numpy(), gold["eps_rows"].astype(np.float32), rtol=1e-6)
    s = torch.from_numpy(inp["state"]).to(DEV)
    action, q_sel = agent.act_device(s, True, noise=torch.from_numpy(inp["u2"]).to(DEV))
    np.testing.assert_array_equal(action.cpu().numpy().reshape(-1, 1), gold["action"])
    np.testing.assert_allclose(q_sel.cpu().numpy(), gold["q"], rtol=1e-5, atol=1e-6)


def test_rainbow_act_matches_reference():
    case, gold = G.Q_CASES["rainbow_small"], load_golden("act_rainbow")
    inp = MC.collect_inputs("act", case)
    agent = _q_agent(case, batch_size=0, start_train_step=0)      # past the random warm-up (rainbow.py:143)
    noise = [(torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV)) for a, b in G.q_case_inputs(case)["noise"][0]]
    s = torch.from_numpy(inp["state"]).to(DEV)
    a, _ = agent.act_device(s, True, noise=noise)
    np.testing.assert_array_equal(a.cpu().numpy().reshape(-1, 1), gold["action_train"])
    a, _ = agent.act_device(s, False)
    np.testing.assert_array_equal(a.cpu().numpy().reshape(-1, 1), gold["action_eval"])


# ---------------------------------------------------------------------------------- n-step assemblers (D5, R2, X2)
@pytest.mark.parametrize("name,case_name,apex", [("nstep_multistep", "multistep_small", False),
                                                 ("nstep_rainbow", "rainbow_small", False), ("nstep_ape_x", "ape_x_small", True)])
def test_nstep_assembler_matches_reference(name, case_name, apex):
    from jorldy_b200.core.collect import NStepAssembler
    case, gold = G.Q_CASES[case_name], load_golden(name)
    inp = MC.collect_inputs("nstep", case)
    T, N = inp["state"].shape[:2]
    asm = NStepAssembler(case["n_step"], apex, case["gamma"])
    k_emit = 0
    for t in range(T):
        tr = {"state": torch.from_numpy(inp["state"][t]).to(DEV), "action": torch.from_numpy(inp["action"][t]).to(DEV),
              "reward": torch.from_numpy(inp["reward"][t, :, 0].astype(np.float32)).to(DEV),
              "done": torch.from_numpy(inp["done"][t, :, 0].astype(np.float32)).to(DEV),
              "next_state": torch.from_numpy(inp["next_state"][t]).to(DEV)}
        if apex:
            tr["q"] = torch.from_numpy(inp["q"][t, :, 0]).to(DEV)
        out = asm.push(tr)
        if t < int(gold["first_emit"]):
            assert out is None
            continue
        assert out is not None
        for k in ("state", "action", "next_state"):
            np.testing.assert_array_equal(out[k].cpu().numpy(), gold[k][k_emit], err_msg=f"{k} step {t}")
        np.testing.assert_array_equal(out["reward"].cpu().numpy(), gold["reward"][k_emit].astype(np.float32))
        np.testing.assert_array_equal(out["done"].cpu().numpy() > 0.5, gold["done"][k_emit])
        if apex:
            assert "q" not in out
            np.testing.assert_allclose(out["priority"].cpu().numpy(), gold["priority"][k_emit], rtol=1e-5, atol=1e-6)
        k_emit += 1
    assert k_emit == gold["state"].shape[0]


# ------------------------------------------------------------------------------------------ whole collect loop (L1)
@pytest.mark.parametrize("use_graph", [False, True])
def test_rollout_collector_matches_oracle_loop(use_graph):
    """T = 8 steps of 64 CartPole actors: the resident collect (act kernel + physics kernel + rollout writes) equals
    the reference loop body (run_mode.py:68-91) replayed on the CPU with the same Philox draws."""
    from jorldy_b200.core import Env
    from jorldy_b200.core.collect import RolloutCollector
    from oracle import collect as oc
    from oracle.classic_control import CartPoleBatch
    case = dict(G.PPO_CASES["ppo_discrete_small"], T=8)
    N, T, seed = 64, 8, 5
    agent = _ppo_agent(case, seed=seed)
    params = {k: torch.from_numpy(v) for k, v in G.make_params(G.ppo_shapes(case), case["seed"]).items()}
    env = Env("cartpole", num_envs=N, seed=3, device=DEV)
    col = RolloutCollector(env, agent, n_step=T, use_cuda_graph=use_graph)
    if use_graph:
        # the capture warm-up consumes draws: rewind the env and the per-row counters to the oracle's starting point
        col.collect()
        env.episode.zero_(); env.reset_device()
        agent._row_ctr[N].zero_()
    ro = col.collect()
    torch.cuda.synchronize()
    ref = oc.rollout_loop(params, CartPoleBatch(N, seed=3, stream_base=0, auto_reset=True), T, seed)
    np.testing.assert_array_equal(ro.action.cpu().numpy(), ref["action"])
    np.testing.assert_array_equal(ro.done.cpu().numpy(), ref["done"])
    np.testing.assert_array_equal(ro.reward.cpu().numpy(), ref["reward"])
    np.testing.assert_allclose(ro.state.cpu().numpy(), ref["state"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(ro.last_next_state.cpu().numpy(), ref["last_next_state"], rtol=0, atol=1e-6)


def test_replay_collector_round_matches_oracle_loop():
    """DQN resident loop with update_period > 1: every transition of the round reaches the replay ring with ITS OWN
    action (regression test: the action workspace is reused by the next act call)."""
    from jorldy_b200.core import Env
    from jorldy_b200.core.collect import ReplayCollector
    from oracle import collect as oc
    from oracle.classic_control import CartPoleBatch
    case = dict(G.Q_CASES["dqn_small"], A=2, buffer_size=4096)
    N, P, seed = 32, 6, 9
    agent = _q_agent(case, epsilon=0.5, seed=seed, start_train_step=10 ** 9)
    params = {k: torch.from_numpy(v) for k, v in G.q_params(case).items()}
    env = Env("cartpole", num_envs=N, seed=4, device=DEV)
    rc = ReplayCollector(env, agent, update_period=P)
    step, _ = rc.run_round(0)
    assert step == P and agent.memory.size == N * P
    cenv = CartPoleBatch(N, seed=4, stream_base=0, auto_reset=True)
    obs = cenv.reset()
    mem = {k: v.cpu().numpy() for k, v in agent.memory.fields.items()}
    for t in range(P):
        u0, u1 = oc.act_uniform(seed, 0, N, t)
        a, _ = oc.act_q(params, obs, 0.5, np.stack([u0, u1], 1), "dqn")
        nobs, r, d = cenv.step(a)
        rows = slice(t * N, (t + 1) * N)              # ring order: step-major, actor-minor
        np.testing.assert_array_equal(mem["action"][rows].reshape(-1), a[:, 0], err_msg=f"action step {t}")
        np.testing.assert_allclose(mem["state"][rows], obs, rtol=0, atol=1e-6)
        np.testing.assert_allclose(mem["next_state"][rows], nobs, rtol=0, atol=1e-6)
        np.testing.assert_array_equal(mem["reward"][rows].reshape(-1), r)
        np.testing.assert_array_equal(mem["done"][rows].reshape(-1) > 0.5, d)
        obs = cenv.obs
