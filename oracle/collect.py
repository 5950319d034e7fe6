"""CPU oracle of the COLLECT side of the hot path — TEST INFRASTRUCTURE (see oracle/__init__.py).

act_ppo()        jorldy/core/agent/ppo.py:54-69 (torch.multinomial -> inverse CDF of an injected uniform, same law;
                 torch.normal(mu, std) -> mu + std * injected normal, its definition)
act_q()          jorldy/core/agent/dqn.py:99-115 and ape_x.py:63-77, ONE reference act() call per row: a row of the batched
                 GPU collect is one reference actor (batch 1), so the single epsilon draw per call (dqn.py:104) becomes
                 one draw per row; np.random.randint(0, A) -> floor(u1 * A)
act_rainbow()    jorldy/core/agent/rainbow.py:139-152 (noisy-greedy on sum_k z_k p_k; logits2Q without max-subtraction,
                 rainbow.py:285-292)
NStepWindow      jorldy/core/agent/multistep.py:90-104 = rainbow.py:294-308, and ape_x.py:174-199 (deque of n+1, actor-side
                 priority |G_n - q_0| bootstrapped from the (n+1)-th step's behaviour q); one deque per actor, never cleared
rollout_loop()   jorldy/run_mode.py:68-91 / manager/distributed_manager.py:76-92 for N actors: act -> env.step -> record
                 -> `state = next_state if not done else env.reset()`, with the Philox draws of the CUDA path
                 (csrc/ppo.cu, csrc/env_classic.cu) so both sides see the same randomness

Pinning: act_* and NStepWindow are checked against fixtures minted from the unmodified reference classes
(tests/golden/make_golden_collect.py -> act_*.npz, nstep_*.npz; tests/test_collect_oracle.py).
"""
from collections import deque

import numpy as np
import torch

from . import nets, philox


def sample_discrete(pi, u):
    """Inverse CDF, float32, ascending action order (the rule of ppo_act_discrete_kernel): the first a with
    u * sum(pi) < cumsum(pi)[a]; the last action if rounding leaves none."""
    pi = np.asarray(pi, dtype=np.float32)
    M, A = pi.shape
    out = np.zeros((M, 1), dtype=np.int64)
    for m in range(M):
        tot = np.float32(0.0)
        for a in range(A):
            tot = np.float32(tot + pi[m, a])
        target = np.float32(np.float32(u[m]) * tot)
        c, pick = np.float32(0.0), A - 1
        for a in range(A):
            c = np.float32(c + pi[m, a])
            if target < c:
                pick = a
                break
        out[m, 0] = pick
    return out


def act_ppo(params, state, continuous, training=True, u=None, eps=None):
    """Returns action: int64 [M,1] (discrete) or float32 [M,A] (continuous)."""
    x = torch.as_tensor(state, dtype=torch.float32)
    with torch.no_grad():
        if continuous:
            mu, std, _ = nets.continuous_policy_value(params, x)
            z = mu + std * torch.as_tensor(eps, dtype=torch.float32) if training else mu
            return torch.tanh(z).numpy()
        pi, _ = nets.discrete_policy_value(params, x)
        if training:
            return sample_discrete(pi.numpy(), u)
        return torch.argmax(pi, dim=-1, keepdim=True).numpy()


def act_q(params, state, eps_rows, u2, net="dqn"):
    """Per-row epsilon-greedy.  u2 [M,2]: (epsilon draw, randint draw).  Returns (action int64 [M,1], q_sel f32 [M])."""
    x = torch.as_tensor(state, dtype=torch.float32)
    with torch.no_grad():
        q = (nets.dueling(params, x) if net == "dueling" else nets.discrete_q_network(params, x)).numpy()
    M, A = q.shape
    eps_rows = np.broadcast_to(np.asarray(eps_rows, dtype=np.float32), (M,))
    action = np.zeros((M, 1), dtype=np.int64)
    for m in range(M):
        if np.float32(u2[m, 0]) < eps_rows[m]:
            action[m, 0] = min(int(np.float32(u2[m, 1]) * np.float32(A)), A - 1)
        else:
            action[m, 0] = int(np.argmax(q[m]))
    return action, q[np.arange(M), action[:, 0]]


def act_rainbow(params, state, n_action, n_atom, v_min, v_max, noise):
    """noise: [(e_i, e_j)] x 4 in call order a1, v1, a2, v2, or None (eval).  Returns action int64 [M,1]."""
    x = torch.as_tensor(state, dtype=torch.float32)
    with torch.no_grad():
        logits = nets.rainbow_network(params, x, noise, n_action, n_atom)
        z = torch.linspace(v_min, v_max, n_atom).view(1, n_atom)
        e = torch.exp(logits)
        p = e / e.sum(dim=-1, keepdim=True)
        q = (z * p).sum(dim=-1)
        return torch.argmax(q, -1, keepdim=True).numpy()


class NStepWindow:
    """One actor's n-step assembler.  push(transition dict with leading dim 1) -> dict or {}."""

    def __init__(self, n_step, apex=False, gamma=0.99):
        self.n, self.apex, self.gamma = n_step, apex, gamma
        self.buf = deque(maxlen=n_step + 1 if apex else n_step)

    def push(self, tr):
        out = {}
        self.buf.append(tr)
        if len(self.buf) < self.buf.maxlen:
            return out
        first, last = self.buf[0], self.buf[-1]
        items = list(self.buf)[:-1] if self.apex else list(self.buf)
        out["state"], out["action"] = first["state"], first["action"]
        out["next_state"] = last["state"] if self.apex else last["next_state"]
        for key in first:
            if key not in ("state", "action", "next_state"):
                out[key] = np.stack([np.asarray(t[key]) for t in items], axis=1)
        if self.apex:
            g = np.asarray(last["q"])
            for i in reversed(range(self.n)):
                g = np.asarray(self.buf[i]["reward"]) + (1 - np.asarray(self.buf[i]["done"])) * self.gamma * g
            out["priority"] = abs(g - np.asarray(first["q"]))
            del out["q"]
        return out


def act_uniform(seed, stream_base, n_rows, ctr):
    """The uniform ppo_act_discrete_kernel / q_act_kernel draw for rows 0..n-1 at per-row counter `ctr`."""
    r = philox.philox4x32(seed, np.uint64(stream_base) + np.arange(n_rows, dtype=np.uint64), np.uint64(ctr))
    return philox.u01_float(r[0]), philox.u01_float(r[1])


def rollout_loop(params, env, T, seed, stream_base=0):
    """PPO discrete collect for env.n actors over T steps.  Returns actor-major arrays
    state [N,T,D], action [N,T], reward [N,T], done [N,T], last_next_state [N,D]."""
    N = env.n
    obs = env.reset()
    D = obs.shape[1]
    S = np.zeros((N, T, D), np.float32); A = np.zeros((N, T), np.int64)
    R = np.zeros((N, T), np.float32); Dn = np.zeros((N, T), np.float32)
    last = None
    for t in range(T):
        S[:, t] = obs
        u, _ = act_uniform(seed, stream_base, N, t)
        a = act_ppo(params, obs, False, True, u=u)
        next_obs, reward, done = env.step(a)
        A[:, t], R[:, t], Dn[:, t] = a[:, 0], reward, done.astype(np.float32)
        last = next_obs
        obs = env.obs          # post auto-reset observation: `state = next_state if not done else env.reset()`
    return dict(state=S, action=A, reward=R, done=Dn, last_next_state=last)
