"""CUDA value-based learners vs the reference golden files and the CPU oracle (pytest -m gpu).

Tolerances (fp32): loss / max_Q / logits stats rtol 2e-5; gradients rtol 2e-3 atol 2e-6;
post-step parameters atol 0.1*lr (Adam / RMSprop normalise the step to ~lr);
new PER priorities rtol 2e-4 atol 1e-6 (|td|^alpha or KL^alpha computed in f32 like the reference)."""
import numpy as np
import pytest
import torch

import gen_inputs as G
from helpers import load_golden, run_q_oracle

pytestmark = pytest.mark.gpu


def _make(case):
    from jorldy_b200.core import Agent
    ag = case["agent"]
    optim = case.get("optim", {"name": "adam", "lr": case["lr"]})
    kw = dict(state_size=case["D"], action_size=case["A"], hidden_size=case["H"], optim_config=dict(optim),
              gamma=case["gamma"], buffer_size=case["buffer_size"], batch_size=case["B"], device="cuda", run_step=1000,
              lr_decay=False)
    if case.get("head"):
        kw["head"] = case["head"]
    if ag in ("multistep", "rainbow", "ape_x"):
        kw["n_step"] = case["n_step"]
    if ag in ("per", "rainbow", "ape_x"):
        kw["alpha"] = case["alpha"]
    if ag in ("c51", "rainbow"):
        kw.update(v_min=case["v_min"], v_max=case["v_max"], num_support=case["K"])
    if ag == "ape_x":
        kw.update(network="dueling", clip_grad_norm=case["clip"], num_workers=2)
    agent = Agent(ag, **kw)
    agent.network.load_state_dict({k: torch.from_numpy(v) for k, v in G.q_params(case).items()})
    agent.target_network.load_state_dict({k: torch.from_numpy(v) for k, v in G.q_params(case, seed_offset=1000).items()})
    return agent


def _run(case):
    agent = _make(case)
    inp = G.q_case_inputs(case)
    dev = "cuda"
    batch = {"state": torch.from_numpy(inp["state"]).to(dev), "next_state": torch.from_numpy(inp["next_state"]).to(dev),
             "action": torch.from_numpy(inp["action"]).to(dev), "reward": torch.from_numpy(inp["reward"]).to(dev),
             "done": torch.from_numpy(inp["done"]).to(dev)}
    ag = case["agent"]
    weights = torch.from_numpy(inp["weights"]).to(dev) if ag in ("per", "rainbow", "ape_x") else None
    noise = None
    if inp["noise"] is not None:
        noise = [[(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b in layers] for layers in inp["noise"]]
    if ag in ("c51", "rainbow"):
        prio = agent._dist_learn(batch, weights, 1 if ag == "rainbow" else 0, noise or [None, None, None])
    else:
        if noise is not None:
            agent._inject_noise = noise
        prio = agent._learn_batch(batch, weights)
    torch.cuda.synchronize()
    return agent, prio


@pytest.mark.parametrize("name", list(G.Q_CASES.keys()))
def test_q_learn_matches_reference_and_oracle(name):
    case = G.Q_CASES[name]
    agent, prio = _run(case)
    gold = load_golden(name)
    ref, inp = run_q_oracle(case)
    lr = case.get("optim", {"lr": case["lr"]})["lr"]
    st = agent._stats.cpu().numpy()
    np.testing.assert_allclose(st[0], float(gold["result.loss"]), rtol=2e-5)
    np.testing.assert_allclose(st[1], float(gold["result.max_Q"]), rtol=2e-5)
    if "result.max_logit" in gold:
        np.testing.assert_allclose(st[2], float(gold["result.max_logit"]), rtol=2e-5)
        np.testing.assert_allclose(st[3], float(gold["result.min_logit"]), rtol=2e-5)
    for k, g in ref["grads"].items():
        np.testing.assert_allclose(agent.network.g[k].cpu().numpy(), g.numpy(), rtol=2e-3, atol=2e-6, err_msg="grad " + k)
    for k, v in gold.items():
        if k.startswith("param."):
            got = G.subsample(agent.network.p[k[6:]].cpu().numpy())
            np.testing.assert_allclose(got, v, rtol=1e-4, atol=0.1 * lr, err_msg=k)
    if "prio.p" in gold:
        np.testing.assert_allclose(prio.cpu().numpy(), gold["prio.p"], rtol=2e-4, atol=1e-6)


def test_dqn_reference_bookkeeping():
    """jorldy/test/core/agent/test_dqn_agent.py:36-37 + test_per_agent.py:42-44 schedule asserts on the
    batched-N=1 plugin API: epsilon reaches epsilon_min, time_t == run_step, beta == 1.0."""
    from jorldy_b200.core import Agent
    run_step, bs = 20, 4
    for name, extra in (("dqn", {}), ("per", {"learn_period": 2}), ("rainbow", {"learn_period": 2, "n_step": 3}),
                        ("ape_x", {"learn_period": 2, "n_step": 3, "num_workers": 2, "network": "dueling"}),
                        ("multistep", {"n_step": 3}), ("c51", {}), ("noisy", {}), ("double", {}), ("dueling", {})):
        agent = Agent(name, state_size=4, action_size=3, hidden_size=32, buffer_size=100, batch_size=bs,
                      start_train_step=8, target_update_period=5, run_step=run_step, explore_ratio=0.5, **extra)
        state = np.random.random((1, 4)).astype(np.float32)
        for step in range(1, run_step + 1):
            ad = agent.act(state, True)
            assert ad["action"].shape == (1, 1)
            ns = np.random.random((1, 4)).astype(np.float32)
            tr = {"state": state, "next_state": ns, "reward": np.random.random((1, 1)),
                  "done": np.random.random((1, 1)) < 0.2}
            tr.update(ad)
            tr = agent.interact_callback(tr)
            if tr:
                agent.process([tr], step)
            state = ns
        assert agent.time_t == run_step
        if name in ("dqn", "double", "dueling", "multistep", "c51", "per"):
            assert agent.epsilon == agent.epsilon_min
        if name in ("per", "rainbow", "ape_x"):
            assert abs(agent.beta - 1.0) < 1e-9
        if name == "ape_x":
            assert agent.memory.size == run_step - 3          # test_ape_x_agent.py:47
        if name == "multistep":
            assert agent.memory.size == run_step - 3 + 1      # test_multistep_agent.py:40
        assert agent.num_learn > 0
