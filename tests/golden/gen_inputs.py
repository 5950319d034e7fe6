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
