"""CPU oracle of the continuous off-policy learners — TEST INFRASTRUCTURE, never imported by the product.

Functional torch-CPU restatement of jorldy/core/agent/{ddpg,td3,sac}.py learn() bodies (and the pieces of act()
that carry arithmetic), with the minibatch and every random draw injected.  Pinned against the UNMODIFIED reference
classes by tests/golden/{ddpg,td3,sac}_*.npz (minted by tests/golden/make_golden_ac.py).

nets                policy.py:8-20 (deterministic_policy), :38-56 (continuous_policy), q_network.py:23-40
ddpg_learn()        ddpg.py:120-158
td3_learn()         td3.py:145-188 (incl. the delayed actor update and the in-learn soft update)
sac_learn()         sac.py:162-260 (continuous branch), alpha lag of sac.py:238-246
soft_update()       ddpg.py:160-164
ou_step()           agent/utils.py:8-26 + ddpg.py:113-118
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import nets


def deterministic_policy(p, x):
    h = F.relu(F.linear(nets.mlp_head(p, x), p["l.weight"], p["l.bias"]))
    return torch.tanh(F.linear(h, p["pi.weight"], p["pi.bias"]))


def continuous_policy(p, x):
    h = F.relu(F.linear(nets.mlp_head(p, x), p["l.weight"], p["l.bias"]))
    mu = torch.clamp(F.linear(h, p["mu.weight"], p["mu.bias"]), min=-5.0, max=5.0)
    log_std = torch.tanh(F.linear(h, p["log_std.weight"], p["log_std.bias"]))
    return mu, log_std.exp()


def continuous_q_network(p, x1, x2):
    x1 = nets.mlp_head(p, x1)
    x2 = F.relu(F.linear(x2, p["e.weight"], p["e.bias"]))
    x = F.relu(F.linear(torch.cat([x1, x2], dim=-1), p["l.weight"], p["l.bias"]))
    return F.linear(x, p["q.weight"], p["q.bias"])


def soft_update(target, online, tau):
    """t := tau * p + (1 - tau) * t, two rounded products and one rounded sum in fp32."""
    return {k: tau * online[k] + (1 - tau) * target[k] for k in target}


def ou_step(X, mu, theta, sigma, normal):
    """One OU_Noise.sample(): X is (1, A); `normal` is the single randn(len(X)) draw (shared by all dims)."""
    dx = theta * (mu - X) + sigma * np.asarray(normal, dtype=np.float64).reshape(1)
    return X + dx


def _leaf(params):
    return {k: v.clone().requires_grad_(True) for k, v in params.items()}


def _adam(p, lr, state):
    opt = torch.optim.Adam(list(p.values()), lr=lr)
    if state is not None:
        opt.load_state_dict(state)
    return opt


def _step(opt, loss, p):
    opt.zero_grad(set_to_none=True)
    loss.backward()
    grads = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    opt.step()
    return grads


def _detach(p):
    return {k: v.detach().clone() for k, v in p.items()}


def ddpg_learn(actor, critic, t_actor, t_critic, batch, hp, opt_state=None):
    """hp: gamma, actor_lr, critic_lr.  Returns post-step params, grads and the result dict."""
    opt_state = opt_state or {}
    a_p, c_p = _leaf(actor), _leaf(critic)
    a_opt, c_opt = _adam(a_p, hp["actor_lr"], opt_state.get("actor")), _adam(c_p, hp["critic_lr"], opt_state.get("critic"))
    s, a, r, ns, d = (batch[k] for k in ("state", "action", "reward", "next_state", "done"))
    with torch.no_grad():
        next_q = continuous_q_network(t_critic, ns, deterministic_policy(t_actor, ns))
        target_q = r + (1 - d) * hp["gamma"] * next_q
    critic_loss = F.mse_loss(target_q, continuous_q_network(c_p, s, a))
    c_grads = _step(c_opt, critic_loss, c_p)
    max_Q = torch.max(target_q, axis=0).values.numpy()[0]
    actor_loss = -continuous_q_network(c_p, s, deterministic_policy(a_p, s)).mean()
    a_grads = _step(a_opt, actor_loss, a_p)
    return {"actor": _detach(a_p), "critic": _detach(c_p), "actor_grads": a_grads, "critic_grads": c_grads,
            "result": {"critic_loss": critic_loss.item(), "actor_loss": actor_loss.item(), "max_Q": float(max_Q)},
            "opt_state": {"actor": a_opt.state_dict(), "critic": c_opt.state_dict()}}


def td3_learn(actor, critic1, critic2, t_actor, t_critic1, t_critic2, batch, hp, noise, num_learn, opt_state=None):
    """hp: gamma, tau, actor_lr, critic_lr, update_delay, target_noise_std, target_noise_clip.  `noise` = the
    torch.randn_like(action) draw.  Returns post-step params of all six networks."""
    opt_state = opt_state or {}
    a_p, c1_p, c2_p = _leaf(actor), _leaf(critic1), _leaf(critic2)
    a_opt = _adam(a_p, hp["actor_lr"], opt_state.get("actor"))
    c1_opt, c2_opt = _adam(c1_p, hp["critic_lr"], opt_state.get("critic1")), _adam(c2_p, hp["critic_lr"], opt_state.get("critic2"))
    s, a, r, ns, d = (batch[k] for k in ("state", "action", "reward", "next_state", "done"))
    with torch.no_grad():
        nz = (noise * hp["target_noise_std"]).clamp(-hp["target_noise_clip"], hp["target_noise_clip"])
        next_action = (deterministic_policy(t_actor, ns) + nz).clamp(-1.0, 1.0)
        min_next_q = torch.min(continuous_q_network(t_critic1, ns, next_action), continuous_q_network(t_critic2, ns, next_action))
        target_q = r + (1 - d) * hp["gamma"] * min_next_q
    loss1 = F.mse_loss(target_q, continuous_q_network(c1_p, s, a))
    g1 = _step(c1_opt, loss1, c1_p)
    loss2 = F.mse_loss(target_q, continuous_q_network(c2_p, s, a))
    g2 = _step(c2_opt, loss2, c2_p)
    max_Q = torch.max(target_q, axis=0).values.numpy()[0]
    out = {"result": {"critic_loss1": loss1.item(), "critic_loss2": loss2.item(), "max_Q": float(max_Q)},
           "critic1_grads": g1, "critic2_grads": g2}
    ta, tc1, tc2 = t_actor, t_critic1, t_critic2
    if num_learn % hp["update_delay"] == 0:
        actor_loss = -continuous_q_network(c1_p, s, deterministic_policy(a_p, s)).mean()
        out["actor_grads"] = _step(a_opt, actor_loss, a_p)
        out["result"]["actor_loss"] = actor_loss.item()
        if num_learn > 0:
            tc1, tc2 = soft_update(tc1, _detach(c1_p), hp["tau"]), soft_update(tc2, _detach(c2_p), hp["tau"])
            ta = soft_update(ta, _detach(a_p), hp["tau"])
    out.update(actor=_detach(a_p), critic1=_detach(c1_p), critic2=_detach(c2_p), target_actor=ta, target_critic1=tc1,
               target_critic2=tc2, opt_state={"actor": a_opt.state_dict(), "critic1": c1_opt.state_dict(),
                                              "critic2": c2_opt.state_dict()})
    return out


def sac_sample_action(mu, std, eps):
    """sac.py:151-160 with Normal.rsample's eps injected."""
    z = mu + eps * std
    action = torch.tanh(z)
    var = std ** 2
    log_prob = -((z - mu) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))
    log_prob = log_prob - torch.log(1 - action.pow(2) + 1e-7)
    return action, log_prob.sum(1, keepdim=True)


def sac_learn(actor, critic1, critic2, t_critic1, t_critic2, log_alpha, alpha, batch, hp, eps_next, eps_actor, opt_state=None):
    """hp: gamma, actor_lr, critic_lr, alpha_lr, use_dynamic_alpha, target_entropy.  `alpha` is the value the agent
    carries INTO this learn() (exp of log_alpha one update ago, sac.py:241); returns the new (log_alpha, alpha)."""
    opt_state = opt_state or {}
    a_p, c1_p, c2_p = _leaf(actor), _leaf(critic1), _leaf(critic2)
    la = log_alpha.clone().requires_grad_(hp["use_dynamic_alpha"])
    a_opt = _adam(a_p, hp["actor_lr"], opt_state.get("actor"))
    c1_opt, c2_opt = _adam(c1_p, hp["critic_lr"], opt_state.get("critic1")), _adam(c2_p, hp["critic_lr"], opt_state.get("critic2"))
    s, a, r, ns, d = (batch[k] for k in ("state", "action", "reward", "next_state", "done"))
    q1, q2 = continuous_q_network(c1_p, s, a), continuous_q_network(c2_p, s, a)
    with torch.no_grad():
        next_action, next_log_prob = sac_sample_action(*continuous_policy(a_p, ns), eps_next)
        min_next_q = torch.min(continuous_q_network(t_critic1, ns, next_action), continuous_q_network(t_critic2, ns, next_action))
        target_q = r + (1 - d) * hp["gamma"] * (min_next_q + alpha * (-next_log_prob))
    max_Q = torch.max(target_q, axis=0).values.numpy()[0]
    loss1, loss2 = F.mse_loss(q1, target_q), F.mse_loss(q2, target_q)
    g1 = _step(c1_opt, loss1, c1_p)
    g2 = _step(c2_opt, loss2, c2_p)
    sample_action, log_prob = sac_sample_action(*continuous_policy(a_p, s), eps_actor)
    entropy = -log_prob
    min_q = torch.min(continuous_q_network(c1_p, s, sample_action), continuous_q_network(c2_p, s, sample_action))
    actor_loss = -((alpha.detach() * entropy) + min_q).mean()
    ga = _step(a_opt, actor_loss, a_p)
    alpha_loss = la * (entropy - hp["target_entropy"]).detach().mean()
    new_alpha = la.detach().exp()
    st = {"actor": a_opt.state_dict(), "critic1": c1_opt.state_dict(), "critic2": c2_opt.state_dict()}
    if hp["use_dynamic_alpha"]:
        al_opt = torch.optim.Adam([la], lr=hp["alpha_lr"])
        if opt_state.get("alpha") is not None:
            al_opt.load_state_dict(opt_state["alpha"])
        al_opt.zero_grad(set_to_none=True)
        alpha_loss.backward()
        al_opt.step()
        st["alpha"] = al_opt.state_dict()
    return {"actor": _detach(a_p), "critic1": _detach(c1_p), "critic2": _detach(c2_p), "actor_grads": ga,
            "critic1_grads": g1, "critic2_grads": g2, "log_alpha": la.detach().clone(), "alpha": new_alpha, "opt_state": st,
            "result": {"critic_loss1": loss1.item(), "critic_loss2": loss2.item(), "actor_loss": actor_loss.item(),
                       "alpha_loss": alpha_loss.item(), "max_Q": float(max_Q), "mean_Q": min_q.mean().item(),
                       "alpha": new_alpha.item(), "entropy": entropy.mean().item()}}


# ---- act() arithmetic (pinned by tests/golden/act_{ddpg,td3,sac}.npz) ---------------------------------------------------
def act_ddpg(actor, state, X, normal, mu, theta, sigma, training=True):
    """ddpg.py:113-118 for ONE actor (state (1, D), OU state X (1, A)): returns (action, new X).  The reference's action is
    float64 (float32 mu + float64 clipped OU state)."""
    with torch.no_grad():
        m = deterministic_policy(actor, torch.as_tensor(state, dtype=torch.float32)).numpy()
    if not training:
        return m, X
    X = ou_step(X, mu, theta, sigma, normal)
    return m + X.clip(-1.0, 1.0), X


def act_td3(actor, state, normal, action_noise_std, training=True):
    """td3.py:137-143: clip(actor(s) + N(0, std), -1, 1); `normal` = standard draws of shape (A,)."""
    with torch.no_grad():
        a = deterministic_policy(actor, torch.as_tensor(state, dtype=torch.float32)).numpy()
    if training:
        a = (a + action_noise_std * np.asarray(normal, dtype=np.float64)).clip(-1.0, 1.0)
    return a


def act_sac(actor, state, eps, training=True):
    """sac.py:139-142: tanh(Normal(mu, std).sample()) = tanh(mu + std * eps); tanh(mu) when not training."""
    with torch.no_grad():
        mu, std = continuous_policy(actor, torch.as_tensor(state, dtype=torch.float32))
        z = mu + std * torch.as_tensor(eps, dtype=torch.float32) if training else mu
        return torch.tanh(z).numpy()
