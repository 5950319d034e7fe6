"""CPU oracle of PPO (jorldy/core/agent/ppo.py), functional style.

gae()            ppo.py:95-110
prepass()        ppo.py:83-94
minibatch_loss() ppo.py:123-162
learn()          ppo.py:71-185 with the minibatch permutation injected (the reference uses an
                 unseeded np.random.shuffle, SURVEY.md hard part 3) and torch.optim.Adam +
                 clip_grad_norm_ exactly as the reference calls them.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch.distributions import Categorical, Normal

from . import nets


def gae(reward, done, value, next_value, n_step, gamma, lam, standardize):
    """All inputs [N*T, 1] float32 tensors, actor-major."""
    delta = reward + (1 - done) * gamma * next_value - value
    adv = delta.clone()
    adv, d = adv.view(-1, n_step), done.view(-1, n_step)
    for t in reversed(range(n_step - 1)):
        adv[:, t] += (1 - d[:, t]) * gamma * lam * adv[:, t + 1]
    ret = adv.view(-1, 1) + value
    if standardize:
        adv = (adv - adv.mean(dim=1, keepdim=True)) / (adv.std(dim=1, keepdim=True) + 1e-7)
    return adv.view(-1, 1), ret


def _dist(p, state, action, continuous):
    if continuous:
        mu, std, v = nets.continuous_policy_value(p, state)
        m = Normal(mu, std)
        z = torch.atanh(torch.clamp(action, -1 + 1e-7, 1 - 1e-7))
        return m, m.log_prob(z), v
    pi, v = nets.discrete_policy_value(p, state)
    return pi, None, v


def prepass(p, state, action, next_state, continuous):
    with torch.no_grad():
        if continuous:
            m, log_prob, value = _dist(p, state, action, True)
            next_value = nets.continuous_policy_value(p, next_state)[-1]
        else:
            pi, _, value = _dist(p, state, action, False)
            log_prob = pi.gather(1, action.long()).log()
            next_value = nets.discrete_policy_value(p, next_state)[-1]
    return value, next_value, log_prob


def minibatch_loss(p, state, action, value_old, ret, adv, log_prob_old, continuous, eps_clip, vf_coef, ent_coef):
    if continuous:
        m, log_prob, value_pred = _dist(p, state, action, True)
    else:
        pi, value_pred = nets.discrete_policy_value(p, state)
        m = Categorical(pi)
        log_prob = m.log_prob(action.squeeze(-1)).unsqueeze(-1)
    ratio = (log_prob - log_prob_old).sum(1, keepdim=True).exp()
    surr1 = ratio * adv
    surr2 = torch.clamp(ratio, min=1 - eps_clip, max=1 + eps_clip) * adv
    actor_loss = -torch.min(surr1, surr2).mean()
    value_pred_clipped = value_old + torch.clamp(value_pred - value_old, -eps_clip, eps_clip)
    critic_loss1 = F.mse_loss(value_pred, ret)
    critic_loss2 = F.mse_loss(value_pred_clipped, ret)
    critic_loss = torch.max(critic_loss1, critic_loss2).mean()
    entropy_loss = -m.entropy().mean()
    loss = actor_loss + vf_coef * critic_loss + ent_coef * entropy_loss
    aux = {"actor_loss": actor_loss, "critic_loss": critic_loss, "entropy_loss": entropy_loss,
           "max_ratio": ratio.max(), "min_prob": log_prob.exp().min(), "value_pred": value_pred}
    return loss, aux


def learn(params, batch, hp, perms, lr, opt_state=None, max_minibatches=None):
    """params: dict name -> tensor (cloned, leaf); batch: dict of [N*T, ...] float32 tensors;
    perms: list (per epoch) of index arrays.  Returns dict with post-step params, first-minibatch
    grads, per-learn stats and the pre-pass tensors."""
    continuous = hp["continuous"]
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    opt = torch.optim.Adam(list(p.values()), lr=lr)
    if opt_state is not None:
        opt.load_state_dict(opt_state)
    state, action, reward = batch["state"], batch["action"], batch["reward"]
    next_state, done = batch["next_state"], batch["done"]
    value, next_value, log_prob_old = prepass(p, state, action, next_state, continuous)
    adv, ret = gae(reward, done, value, next_value, hp["n_step"], hp["gamma"], hp["lambda"], hp["standardize"])
    out = {"value": value.clone(), "next_value": next_value.clone(), "log_prob_old": log_prob_old.clone(),
           "adv": adv.clone(), "ret": ret.clone(), "mean_ret": ret.mean().item()}
    stats = {k: [] for k in ["actor_loss", "critic_loss", "entropy_loss", "max_ratio", "min_prob"]}
    B = hp["batch_size"]
    n_done = 0
    first_grads = None
    for epoch in range(hp["n_epoch"]):
        idxs = np.asarray(perms[epoch])
        for offset in range(0, len(reward), B):
            idx = idxs[offset:offset + B]
            loss, aux = minibatch_loss(p, state[idx], action[idx], value[idx], ret[idx], adv[idx], log_prob_old[idx],
                                       continuous, hp["eps_clip"], hp["vf_coef"], hp["ent_coef"])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            if first_grads is None:
                first_grads = {k: v.grad.clone() for k, v in p.items()}
                out["first_value_pred"] = aux["value_pred"].detach().clone()
            torch.nn.utils.clip_grad_norm_(list(p.values()), hp["clip_grad_norm"])
            opt.step()
            for k in stats:
                stats[k].append(aux[k].item())
            n_done += 1
            if max_minibatches is not None and n_done >= max_minibatches:
                break
        if max_minibatches is not None and n_done >= max_minibatches:
            break
    out["params"] = {k: v.detach().clone() for k, v in p.items()}
    out["first_grads"] = first_grads
    out["result"] = {
        "actor_loss": float(np.mean(stats["actor_loss"])), "critic_loss": float(np.mean(stats["critic_loss"])),
        "entropy_loss": float(np.mean(stats["entropy_loss"])), "max_ratio": max(stats["max_ratio"]),
        "min_prob": min(stats["min_prob"]), "mean_ret": out["mean_ret"]}
    out["opt_state"] = opt.state_dict()
    return out
