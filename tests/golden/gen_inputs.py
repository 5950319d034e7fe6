"""Deterministic synthetic inputs shared by make_golden.py (reference side) and the tests
(oracle / CUDA side): everything is a pure function of a case dict, so fixtures need only store
the reference's OUTPUTS."""
import numpy as np


def make_params(shapes, seed):
    """shapes: ordered dict name -> shape.  Matrices ~ N(0, 1/fan_in), biases ~ N(0, 0.1)."""
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        if len(shape) == 1:
            out[name] = (0.1 * rs.standard_normal(shape)).astype(np.float32)
        else:
            fan_in = int(np.prod(shape[1:]))
            out[name] = (rs.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32)
    return out


def ppo_case_inputs(case):
    rs = np.random.RandomState(case["seed"] + 1)
    N, T, D, A = case["N"], case["T"], case["D"], case["A"]
    NT = N * T
    state = (0.5 * rs.standard_normal((NT, D))).astype(np.float32)
    next_state = (0.5 * rs.standard_normal((NT, D))).astype(np.float32)
    if case["continuous"]:
        action = np.tanh(rs.standard_normal((NT, A))).astype(np.float32)
    else:
        action = rs.randint(0, A, size=(NT, 1)).astype(np.int64)
    reward = rs.choice([0.1, -1.0, 0.5], size=(NT, 1)).astype(np.float64)
    done = rs.uniform(size=(NT, 1)) < 0.07
    n_mb = NT // case["batch_size"] + (1 if NT % case["batch_size"] else 0)
    perms = [rs.permutation(NT) for _ in range(case["n_epoch"])]
    return dict(state=state, next_state=next_state, action=action, reward=reward, done=done, perms=perms, n_mb=n_mb)


def ppo_shapes(case):
    H, D, A = case["H"], case["D"], case["A"]
    from collections import OrderedDict
    s = OrderedDict()
    s["head.l.weight"] = (H, D); s["head.l.bias"] = (H,)
    s["l.weight"] = (H, H); s["l.bias"] = (H,)
    if case["continuous"]:
        s["mu.weight"] = (A, H); s["mu.bias"] = (A,)
        s["log_std.weight"] = (A, H); s["log_std.bias"] = (A,)
    else:
        s["pi.weight"] = (A, H); s["pi.bias"] = (A,)
    s["v.weight"] = (1, H); s["v.bias"] = (1,)
    return s


PPO_CASES = {
    "ppo_discrete_small": dict(seed=11, N=4, T=16, D=4, A=2, H=64, continuous=False, batch_size=16, n_epoch=2,
                               lr=2.5e-4, gamma=0.99, lam=0.95, eps_clip=0.1, vf_coef=1.0, ent_coef=0.01,
                               clip_grad_norm=1.0, standardize=True),
    "ppo_continuous_small": dict(seed=12, N=4, T=16, D=11, A=3, H=64, continuous=True, batch_size=32, n_epoch=2,
                                 lr=3e-4, gamma=0.99, lam=0.95, eps_clip=0.1, vf_coef=1.0, ent_coef=0.01,
                                 clip_grad_norm=1.0, standardize=True),
    "ppo_discrete_h512": dict(seed=13, N=8, T=128, D=4, A=2, H=512, continuous=False, batch_size=256, n_epoch=1,
                              lr=2.5e-4, gamma=0.99, lam=0.95, eps_clip=0.1, vf_coef=1.0, ent_coef=0.01,
                              clip_grad_norm=1.0, standardize=True),
    "ppo_continuous_h512": dict(seed=14, N=4, T=64, D=11, A=3, H=512, continuous=True, batch_size=64, n_epoch=1,
                                lr=3e-4, gamma=0.99, lam=0.95, eps_clip=0.1, vf_coef=1.0, ent_coef=0.01,
                                clip_grad_norm=1.0, standardize=False),
}

SUBSAMPLE = 61   # large tensors are stored strided by this prime


def subsample(a):
    a = np.asarray(a).reshape(-1)
    return a if a.size <= 8192 else a[::SUBSAMPLE].copy()


# ---- value-based agents --------------------------------------------------------------------------
def _head_shapes(s, case):
    H, D = case["H"], case["D"]
    if case.get("head", "mlp") == "cnn":
        s["head.conv1.weight"] = (32, 4, 8, 8); s["head.conv1.bias"] = (32,)
        s["head.conv2.weight"] = (64, 32, 4, 4); s["head.conv2.bias"] = (64,)
        s["head.conv3.weight"] = (64, 64, 3, 3); s["head.conv3.bias"] = (64,)
        return 3136
    s["head.l.weight"] = (H, D); s["head.l.bias"] = (H,)
    return H


def q_shapes(case):
    from collections import OrderedDict
    H, D, A = case["H"], case["D"], case["A"]
    K = case.get("K", 51)
    net = case["net"]
    s = OrderedDict()
    if case.get("head", "mlp") == "cnn":
        return _q_shapes_cnn(case)
    if net == "discrete_q_network":
        out = A * K if case["agent"] == "c51" else A
        s["head.l.weight"] = (H, D); s["head.l.bias"] = (H,)
        s["l.weight"] = (H, H); s["l.bias"] = (H,)
        s["q.weight"] = (out, H); s["q.bias"] = (out,)
    elif net == "dueling":
        s["head.l.weight"] = (H, D); s["head.l.bias"] = (H,)
        for n in ("l1_a", "l1_v"):
            s[f"{n}.weight"] = (H, H); s[f"{n}.bias"] = (H,)
        s["l2_a.weight"] = (A, H); s["l2_a.bias"] = (A,)
        s["l2_v.weight"] = (1, H); s["l2_v.bias"] = (1,)
    elif net == "noisy":
        for t, (i, o) in (("1", (H, H)), ("2", (H, A))):
            s[f"mu_w{t}"] = (i, o); s[f"sig_w{t}"] = (i, o); s[f"mu_b{t}"] = (o,); s[f"sig_b{t}"] = (o,)
        s["head.l.weight"] = (H, D); s["head.l.bias"] = (H,)
    elif net == "rainbow":
        for t, (i, o) in (("_a1", (H, H)), ("_v1", (H, H)), ("_a2", (H, A * K)), ("_v2", (H, K))):
            s[f"mu_w{t}"] = (i, o); s[f"sig_w{t}"] = (i, o); s[f"mu_b{t}"] = (o,); s[f"sig_b{t}"] = (o,)
        s["head.l.weight"] = (H, D); s["head.l.bias"] = (H,)
        s["l.weight"] = (H, H); s["l.bias"] = (H,)
    return s


def _q_shapes_cnn(case):
    from collections import OrderedDict
    H, A, K = case["H"], case["A"], case.get("K", 51)
    s = OrderedDict()
    if case["net"] == "discrete_q_network":
        F = _head_shapes(s, case)
        s["l.weight"] = (H, F); s["l.bias"] = (H,)
        s["q.weight"] = (A, H); s["q.bias"] = (A,)
    elif case["net"] == "dueling":
        F = _head_shapes(s, case)
        for n in ("l1_a", "l1_v"):
            s[f"{n}.weight"] = (H, F); s[f"{n}.bias"] = (H,)
        s["l2_a.weight"] = (A, H); s["l2_a.bias"] = (A,)
        s["l2_v.weight"] = (1, H); s["l2_v.bias"] = (1,)
    elif case["net"] == "rainbow":
        for t, (i, o) in (("_a1", (H, H)), ("_v1", (H, H)), ("_a2", (H, A * K)), ("_v2", (H, K))):
            s[f"mu_w{t}"] = (i, o); s[f"sig_w{t}"] = (i, o); s[f"mu_b{t}"] = (o,); s[f"sig_b{t}"] = (o,)
        F = _head_shapes(s, case)
        s["l.weight"] = (H, F); s["l.bias"] = (H,)
    return s


def q_params(case, seed_offset=0):
    """Noisy (in,out) matrices get fan_in = in (make_params uses prod(shape[1:]) which is `out`; fine —
    only determinism matters); sigma tensors are scaled down to the 0.5/sqrt(in) range."""
    p = make_params(q_shapes(case), case["seed"] + seed_offset)
    for k in p:
        if k.startswith("sig_"):
            p[k] = (0.3 * np.abs(p[k])).astype(np.float32)
    return p


def q_case_inputs(case):
    rs = np.random.RandomState(case["seed"] + 7)
    B, D, A, n = case["B"], case["D"], case["A"], case.get("n_step", 1)
    if case.get("head", "mlp") == "cnn":
        state = rs.randint(0, 256, size=(B, 4, 84, 84)).astype(np.uint8)
        next_state = rs.randint(0, 256, size=(B, 4, 84, 84)).astype(np.uint8)
    else:
        state = (0.7 * rs.standard_normal((B, D))).astype(np.float32)
        next_state = (0.7 * rs.standard_normal((B, D))).astype(np.float32)
    action = rs.randint(0, A, size=(B, 1)).astype(np.int64)
    if n == 1 and case["agent"] not in ("multistep", "rainbow", "ape_x"):
        reward = rs.choice([0.1, -1.0, 1.0, 2.5], size=(B, 1)).astype(np.float64)
        done = rs.uniform(size=(B, 1)) < 0.2
    else:
        reward = rs.choice([0.1, -1.0, 1.0, 2.5], size=(B, n, 1)).astype(np.float64)
        done = rs.uniform(size=(B, n, 1)) < 0.15
    weights = rs.uniform(0.2, 1.0, size=B)
    weights = weights / weights.max()
    indices = rs.randint(0, case.get("buffer_size", 64), size=B) + case.get("buffer_size", 64) - 1
    n_layers = {"noisy": 2, "rainbow": 4}.get(case["net"], 0)
    noise = None
    if n_layers:
        shapes = q_shapes(case)
        tags = ["1", "2"] if case["net"] == "noisy" else ["_a1", "_v1", "_a2", "_v2"]
        noise = []
        for _pass in range(3):
            layers = []
            for t in tags:
                i, o = shapes[f"mu_w{t}"]
                layers.append((rs.standard_normal(i).astype(np.float32), rs.standard_normal(o).astype(np.float32)))
            noise.append(layers)
    return dict(state=state, next_state=next_state, action=action, reward=reward, done=done, weights=weights,
                indices=indices, noise=noise)


_QBASE = dict(D=4, A=3, H=64, B=16, lr=1e-3, gamma=0.99, buffer_size=64)
Q_CASES = {
    "dqn_small": dict(_QBASE, seed=21, agent="dqn", net="discrete_q_network"),
    "double_small": dict(_QBASE, seed=22, agent="double", net="discrete_q_network"),
    "dueling_small": dict(_QBASE, seed=23, agent="dueling", net="dueling"),
    "multistep_small": dict(_QBASE, seed=24, agent="multistep", net="discrete_q_network", n_step=4),
    "per_small": dict(_QBASE, seed=25, agent="per", net="discrete_q_network", alpha=0.6),
    "noisy_small": dict(_QBASE, seed=26, agent="noisy", net="noisy"),
    "c51_small": dict(_QBASE, seed=27, agent="c51", net="discrete_q_network", K=51, v_min=-1, v_max=10),
    "rainbow_small": dict(_QBASE, seed=28, agent="rainbow", net="rainbow", K=51, v_min=-1, v_max=10, n_step=3, alpha=0.5),
    "ape_x_small": dict(_QBASE, seed=29, agent="ape_x", net="dueling", n_step=3, alpha=0.6, clip=40.0,
                        optim={"name": "rmsprop", "eps": 1.5e-7, "lr": 1e-3, "centered": True}),
    "dqn_cnn": dict(_QBASE, seed=32, agent="dqn", net="discrete_q_network", head="cnn", D=[4, 84, 84], A=4, B=8, H=64),
    "rainbow_cnn": dict(_QBASE, seed=33, agent="rainbow", net="rainbow", head="cnn", D=[4, 84, 84], A=4, B=8, H=64, K=51,
                        v_min=-1, v_max=10, n_step=3, alpha=0.5, lr=6.25e-5),
    "ape_x_cnn": dict(_QBASE, seed=34, agent="ape_x", net="dueling", head="cnn", D=[4, 84, 84], A=4, B=8, H=64, n_step=3,
                      alpha=0.6, clip=40.0, optim={"name": "rmsprop", "eps": 1.5e-7, "lr": 6.25e-5, "centered": True}),
    "dqn_h512": dict(_QBASE, seed=30, agent="dqn", net="discrete_q_network", H=512, B=32, A=2),
    "rainbow_h512": dict(_QBASE, seed=31, agent="rainbow", net="rainbow", H=512, B=32, A=2, K=51, v_min=-1, v_max=10,
                         n_step=3, alpha=0.5, lr=6.25e-5),
}


# ---- continuous off-policy family (DDPG / TD3 / SAC) -------------------------------------------------------------------
def ac_shapes(case, which):
    """which: "actor" | "critic" — state_dict key order of policy.py:8-56 / q_network.py:23-40."""
    from collections import OrderedDict
    H, D, A = case["H"], case["D"], case["A"]
    s = OrderedDict()
    s["head.l.weight"] = (H, D); s["head.l.bias"] = (H,)
    if which == "critic":
        s["e.weight"] = (H, A); s["e.bias"] = (H,)
        s["l.weight"] = (H, 2 * H); s["l.bias"] = (H,)
        s["q.weight"] = (1, H); s["q.bias"] = (1,)
    else:
        s["l.weight"] = (H, H); s["l.bias"] = (H,)
        if case["agent"] == "sac":
            s["mu.weight"] = (A, H); s["mu.bias"] = (A,)
            s["log_std.weight"] = (A, H); s["log_std.bias"] = (A,)
        else:
            s["pi.weight"] = (A, H); s["pi.bias"] = (A,)
    return s


AC_NETS = ("actor", "critic1", "critic2", "target_actor", "target_critic1", "target_critic2")


def ac_params(case, net):
    """Seeded parameters of one of AC_NETS (targets get their own seeds: the learners must not assume target == online)."""
    which = "actor" if "actor" in net else "critic"
    return make_params(ac_shapes(case, which), case["seed"] + 100 * (1 + AC_NETS.index(net)))


def ac_case_inputs(case):
    rs = np.random.RandomState(case["seed"] + 5)
    B, D, A, n = case["B"], case["D"], case["A"], case.get("n_learns", 1)
    out = dict(state=(0.7 * rs.standard_normal((B, D))).astype(np.float32),
               next_state=(0.7 * rs.standard_normal((B, D))).astype(np.float32),
               action=np.tanh(rs.standard_normal((B, A))).astype(np.float32),
               reward=rs.choice([0.1, -1.0, 1.0, 2.5], size=(B, 1)).astype(np.float64),
               done=rs.uniform(size=(B, 1)) < 0.2)
    # one set of normal draws per learn(): TD3 target smoothing / SAC next-state and actor reparameterisation noise
    out["noise"] = [{"target": rs.standard_normal((B, A)).astype(np.float32),
                     "next": rs.standard_normal((B, A)).astype(np.float32),
                     "actor": rs.standard_normal((B, A)).astype(np.float32)} for _ in range(n)]
    return out


_ACBASE = dict(D=3, A=1, H=64, B=16, gamma=0.99, actor_lr=5e-4, critic_lr=1e-3, alpha_lr=3e-4, tau=5e-3)
AC_CASES = {
    "ddpg_small": dict(_ACBASE, seed=41, agent="ddpg"),
    "ddpg_h512": dict(_ACBASE, seed=42, agent="ddpg", D=11, A=3, H=512, B=128),
    "td3_first": dict(_ACBASE, seed=43, agent="td3", A=2, num_learn=0),       # actor update, no soft update (num_learn == 0)
    "td3_skip": dict(_ACBASE, seed=44, agent="td3", A=2, num_learn=1),        # critics only
    "td3_delayed": dict(_ACBASE, seed=45, agent="td3", D=11, A=3, num_learn=2),   # actor + soft update of the three targets
    "sac_static": dict(_ACBASE, seed=46, agent="sac", A=2, dynamic_alpha=False),
    "sac_dynamic": dict(_ACBASE, seed=47, agent="sac", D=11, A=3, dynamic_alpha=True, n_learns=2),   # two learns: alpha lags
    "sac_h512": dict(_ACBASE, seed=48, agent="sac", D=11, A=3, H=512, B=64, dynamic_alpha=True),
}
