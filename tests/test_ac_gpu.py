"""CUDA DDPG / TD3 / SAC learners vs the reference golden files and the CPU oracle (pytest -m gpu).

Tolerances (fp32): result-dict entries rtol 1e-4 atol 2e-5; gradients rtol 2e-3 atol 2e-6 (vs oracle/actor_critic.py);
post-step parameters atol 0.1*lr (Adam normalises the step to ~lr); soft-updated target parameters atol 2e-6;
log_alpha / alpha rtol 1e-5.  The soft update itself is bit-exact."""
import numpy as np
import pytest
import torch

import gen_inputs as G
from helpers import load_golden, run_ac_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _make(case, **extra):
    from jorldy_b200.core import Agent
    optim = {"actor": "adam", "critic": "adam", "alpha": "adam", "actor_lr": case["actor_lr"], "critic_lr": case["critic_lr"],
             "alpha_lr": case["alpha_lr"]}
    kw = dict(state_size=case["D"], action_size=case["A"], hidden_size=case["H"], optim_config=optim, gamma=case["gamma"],
              buffer_size=64, batch_size=case["B"], device=DEV, run_step=1000, lr_decay=False, tau=case["tau"])
    if case["agent"] == "sac":
        kw["use_dynamic_alpha"] = case["dynamic_alpha"]
    kw.update(extra)
    agent = Agent(case["agent"], **kw)
    ld = lambda net, name: net.load_state_dict({k: torch.from_numpy(v) for k, v in G.ac_params(case, name).items()})
    ld(agent.actor, "actor")
    for i, c in enumerate(agent.critics):
        ld(c, f"critic{i + 1}")
        ld(agent.target_critics[i], f"target_critic{i + 1}")
    if hasattr(agent, "target_actor"):
        ld(agent.target_actor, "target_actor")
    return agent


def _nets(agent):
    out = {"actor": agent.actor}
    for i, c in enumerate(agent.critics):
        out[f"critic{i + 1}"] = c
        out[f"target_critic{i + 1}"] = agent.target_critics[i]
    if hasattr(agent, "target_actor"):
        out["target_actor"] = agent.target_actor
    return out


@pytest.mark.parametrize("name", list(G.AC_CASES.keys()))
def test_ac_learn_matches_reference_and_oracle(name):
    case = G.AC_CASES[name]
    gold = load_golden(name)
    outs, _ = run_ac_oracle(case)
    agent = _make(case)
    inp = G.ac_case_inputs(case)
    batch = {k: torch.from_numpy(inp[k]).to(DEV) for k in ("state", "next_state", "action", "reward", "done")}
    if case["agent"] == "td3":
        agent.num_learn = case["num_learn"]
    for i, nz in enumerate(inp["noise"]):
        agent._inject_noise = {k: torch.from_numpy(v).to(DEV) for k, v in nz.items()}
        res = agent._learn_batch(batch)
        torch.cuda.synchronize()
        for k, v in res.items():
            np.testing.assert_allclose(v, float(gold[f"result{i}.{k}"]), rtol=1e-4, atol=2e-5, err_msg=f"{k} (learn {i})")
        if i == 0:                # gradient buffers after the first learn vs autograd on the oracle
            o = outs[0]
            for net, key in (("actor", "actor_grads"), ("critic1", "critic_grads" if case["agent"] == "ddpg" else "critic1_grads"),
                             ("critic2", "critic2_grads")):
                if key in o and net in _nets(agent):
                    for k, g in o[key].items():
                        np.testing.assert_allclose(_nets(agent)[net].g[k].cpu().numpy(), g.numpy(), rtol=2e-3, atol=2e-6,
                                                   err_msg=f"grad {net}.{k}")
    nets = _nets(agent)
    lr = max(case["actor_lr"], case["critic_lr"])
    for k, v in gold.items():
        if k.startswith("param."):
            _, net, key = k.split(".", 2)
            got = G.subsample(nets[net].p[key].cpu().numpy())
            atol = 2e-6 if net.startswith("target") else 0.1 * lr
            np.testing.assert_allclose(got, v, rtol=1e-4, atol=atol, err_msg=k)
    if case["agent"] == "sac":
        np.testing.assert_allclose(agent.log_alpha.flat[0].item(), float(gold["log_alpha"]), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(agent.alpha.item(), float(gold["alpha"]), rtol=1e-5)


def test_soft_update_is_bit_exact():
    from jorldy_b200.core.dev import C, ptr, stream_ptr
    g = torch.Generator().manual_seed(3)
    p, t = torch.randn(100003, generator=g), torch.randn(100003, generator=g)
    for tau in (1e-3, 5e-3, 0.37):
        want = tau * p + (1 - tau) * t                       # ddpg.py:162 on the CPU
        td = t.to(DEV)
        C.jb_soft_update(ptr(td), ptr(p.to(DEV)), td.numel(), tau, stream_ptr())
        assert torch.equal(td.cpu(), want)


def test_philox_fill_moments_and_freshness():
    case = G.AC_CASES["td3_first"]
    agent = _make(case)
    a = agent._fill("x", (4096, 3), 5).clone()
    b = agent._fill("x", (4096, 3), 5).clone()
    assert not torch.equal(a, b)                             # the device counter advances: fresh draws per call
    assert abs(a.mean().item()) < 0.03 and abs(a.std().item() - 1.0) < 0.03
    u = agent._fill("u", (4097,), 6, kind=1, lo=-1.0, hi=1.0)
    assert u.min().item() >= -1.0 and u.max().item() < 1.0 and abs(u.mean().item()) < 0.05


def test_act_paths_match_the_oracle():
    from oracle import actor_critic as oac
    rs = np.random.RandomState(9)
    N = 33
    # DDPG: tanh(actor(s)) + clip(OU) with one normal per env and step, X carried across steps
    case = G.AC_CASES["ddpg_h512"]
    agent = _make(case, theta=0.15, sigma=0.2)
    actor = {k: torch.from_numpy(v) for k, v in G.ac_params(case, "actor").items()}
    X = [np.zeros((1, case["A"])) for _ in range(N)]
    for step in range(3):
        s = (0.7 * rs.standard_normal((N, case["D"]))).astype(np.float32)
        n = rs.standard_normal(N)
        got, _ = agent.act_device(torch.from_numpy(s).to(DEV), True, noise=torch.from_numpy(n).to(DEV))
        mu = oac.deterministic_policy(actor, torch.from_numpy(s)).numpy()
        want = np.zeros_like(mu, dtype=np.float64)
        for e in range(N):
            X[e] = oac.ou_step(X[e], 0.0, 0.15, 0.2, n[e])
            want[e] = mu[e] + X[e].clip(-1.0, 1.0)[0]
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=2e-6)
    greedy, _ = agent.act_device(torch.from_numpy(s).to(DEV), False)
    np.testing.assert_allclose(greedy.cpu().numpy(), mu, rtol=1e-5, atol=2e-6)
    # TD3: clip(tanh(actor(s)) + N(0, 0.1), -1, 1)
    case = G.AC_CASES["td3_delayed"]
    agent = _make(case, action_noise_std=0.4)
    actor = {k: torch.from_numpy(v) for k, v in G.ac_params(case, "actor").items()}
    s = (0.7 * rs.standard_normal((N, case["D"]))).astype(np.float32)
    n = rs.standard_normal((N, case["A"])).astype(np.float32)
    got, _ = agent.act_device(torch.from_numpy(s).to(DEV), True, noise=torch.from_numpy(n).to(DEV))
    want = (oac.deterministic_policy(actor, torch.from_numpy(s)).numpy() + n * np.float32(0.4)).clip(-1.0, 1.0)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=2e-6)
    # ... and uniform(-1, 1) actions for the first initial_random_step calls (td3.py:133-135)
    agent = _make(case, initial_random_step=2)
    u0 = agent.act_device(torch.from_numpy(s).to(DEV), True)[0].clone()
    u1 = agent.act_device(torch.from_numpy(s).to(DEV), True)[0].clone()
    assert agent.num_random_step == 2 and not torch.equal(u0, u1) and u0.abs().max().item() <= 1.0
    # SAC: tanh(Normal(mu, std).sample())
    case = G.AC_CASES["sac_dynamic"]
    agent = _make(case)
    mu, std = oac.continuous_policy(actor_sac := {k: torch.from_numpy(v) for k, v in G.ac_params(case, "actor").items()},
                                    torch.from_numpy(s))
    got, _ = agent.act_device(torch.from_numpy(s).to(DEV), True, noise=torch.from_numpy(n).to(DEV))
    np.testing.assert_allclose(got.cpu().numpy(), torch.tanh(mu + std * torch.from_numpy(n)).numpy(), rtol=1e-5, atol=2e-6)
    got, _ = agent.act_device(torch.from_numpy(s).to(DEV), False)
    np.testing.assert_allclose(got.cpu().numpy(), torch.tanh(mu).numpy(), rtol=1e-5, atol=2e-6)


def test_reference_bookkeeping_through_the_plugin_api():
    """jorldy/test/core/agent/test_{ddpg,td3,sac}_agent.py shape: act -> process loop on numpy transitions; learning starts at
    start_train_step, targets move only after the first learn, TD3 updates its actor every second learn."""
    from jorldy_b200.core import Agent
    run_step, bs = 24, 4
    for name, extra in (("ddpg", {}), ("td3", {"update_delay": 2}), ("sac", {"use_dynamic_alpha": True}), ("sac", {})):
        agent = Agent(name, state_size=3, action_size=2, hidden_size=32, buffer_size=100, batch_size=bs, start_train_step=8,
                      run_step=run_step, **extra)
        tgt0 = agent.target_critics[0].flat.clone()
        state = np.random.random((1, 3)).astype(np.float32)
        n_res = 0
        for step in range(1, run_step + 1):
            ad = agent.act(state, True)
            assert ad["action"].shape == (1, 2) and np.all(np.abs(ad["action"]) <= 1.0 + 1e-6 + (name == "ddpg"))
            ns = np.random.random((1, 3)).astype(np.float32)
            tr = {"state": state, "next_state": ns, "reward": np.random.random((1, 1)), "done": np.random.random((1, 1)) < 0.2}
            tr.update(ad)
            res = agent.process([agent.interact_callback(tr)], step)
            if step < 8:
                assert res == {} and torch.equal(agent.target_critics[0].flat, tgt0)
            n_res += bool(res)
            state = ns
        assert agent.num_learn == run_step - 8 + 1 == n_res           # one learn per process() from start_train_step on
        assert not torch.equal(agent.target_critics[0].flat, tgt0)
        assert all(np.isfinite(v) for v in res.values()), res
        if name == "sac":
            assert set(res) == {"critic_loss1", "critic_loss2", "actor_loss", "alpha_loss", "max_Q", "mean_Q", "alpha", "entropy"}
            if extra:
                assert abs(res["alpha"] - 1.0) > 1e-6                       # the learned temperature moves away from exp(0)
            else:
                assert abs(res["alpha"] - np.exp(-2.0)) < 1e-6              # static_log_alpha = -2 (sac.py:57)
        lr_now = agent.actor_optimizer.param_groups[0]["lr"]
        assert lr_now < agent.actor_optimizer.defaults["lr"]                # cosine decay applied after every learn


def test_checkpoint_round_trip_keeps_the_reference_layout(tmp_path):
    from jorldy_b200.core import Agent
    mk = lambda: Agent("sac", state_size=3, action_size=1, hidden_size=32, buffer_size=64, batch_size=4, start_train_step=1,
                       use_dynamic_alpha=True, run_step=100)
    a = mk()
    s = np.random.random((8, 3)).astype(np.float32)
    tr = {"state": s, "next_state": s[::-1].copy(), "reward": np.ones((8, 1)), "done": np.zeros((8, 1), dtype=bool),
          "action": a.act(s, True)["action"]}
    for step in range(1, 4):
        a.process([tr], step)
    a.save(str(tmp_path))
    ck = torch.load(str(tmp_path / "ckpt"), map_location="cpu", weights_only=False)
    assert set(ck) == {"actor", "actor_optimizer", "critic1", "critic2", "critic_optimizer1", "critic_optimizer2", "log_alpha",
                       "alpha_optimizer"}                                    # sac.py:306-319
    assert list(ck["critic1"]) == ["head.l.weight", "head.l.bias", "e.weight", "e.bias", "l.weight", "l.bias", "q.weight", "q.bias"]
    b = mk()
    b.load(str(tmp_path))
    assert torch.equal(b.actor.flat, a.actor.flat)
    assert torch.equal(b.critics[0].flat, a.critics[1].flat)                 # sac.py:329: critic2's weights land in critic1
    assert torch.equal(b.log_alpha.flat, a.log_alpha.flat)


def test_replay_collector_runs_the_three_agents_on_pendulum():
    from jorldy_b200.core import Agent, Env
    from jorldy_b200.core.collect import ReplayCollector
    for name in ("ddpg", "td3", "sac"):
        env = Env("pendulum", num_envs=32, seed=3, device=DEV)
        agent = Agent(name, state_size=3, action_size=1, hidden_size=64, buffer_size=4096, batch_size=64, start_train_step=8,
                      run_step=10 ** 5, device=DEV)           # `step` counts env steps per actor: 4 per round
        rc = ReplayCollector(env, agent, update_period=4)
        step, res = 0, {}
        for _ in range(6):
            step, res = rc.run_round(step)
        assert agent.num_learn == 5 and agent.memory.size == 32 * 4 * 6
        assert res and all(np.isfinite(v) for v in res.values()), (name, res)


def test_cuda_graph_learn_is_bit_identical_to_eager():
    """learn() replays one CUDA graph per variant (TD3: with / without the delayed actor + target update); the eager path
    launches the same kernels one by one.  Same weights, same replay contents, same minibatch indices, same Philox streams:
    results and every network must agree bit for bit, across the eager warm-up, the capture and the replays."""
    from jorldy_b200.core import Agent
    rs = np.random.RandomState(21)
    s = (0.7 * rs.standard_normal((128, 3))).astype(np.float32)
    tr = {"state": s, "next_state": (0.7 * rs.standard_normal((128, 3))).astype(np.float32),
          "action": np.tanh(rs.standard_normal((128, 2))).astype(np.float32), "reward": rs.standard_normal((128, 1)),
          "done": rs.uniform(size=(128, 1)) < 0.2}
    for name, extra in (("ddpg", {}), ("td3", {}), ("sac", {"use_dynamic_alpha": True})):
        mk = lambda g: Agent(name, state_size=3, action_size=2, hidden_size=64, buffer_size=256, batch_size=32, start_train_step=1,
                             run_step=1000, seed=11, device=DEV, use_cuda_graph=g, **extra)
        a, b = mk(True), mk(False)
        for x, y in zip(_nets(a).values(), _nets(b).values()):
            y.flat.copy_(x.flat)
        a.memory.store([tr]); b.memory.store([tr])
        for i in range(8):
            a._inject_idx = b._inject_idx = rs.randint(128, size=32)
            ra, rb = a.learn(), b.learn()
            assert ra == rb, (name, i, ra, rb)
            if name == "ddpg":
                a.update_target_soft(); b.update_target_soft()
        torch.cuda.synchronize()
        for (k, x), y in zip(_nets(a).items(), _nets(b).values()):
            assert torch.equal(x.flat, y.flat), (name, k)
        assert len(a._graphs) == (2 if name == "td3" else 1) and not b._graphs
