"""CPU oracle of the value-based learners (jorldy/core/agent/{dqn,double,dueling,multistep,per,noisy,
c51,rainbow,ape_x}.py learn() bodies), functional style with the minibatch, PER weights and
NoisyNet draws injected.

td_learn()       dqn.py:117-151, double.py:13-52, multistep.py:25-64, per.py:37-88, ape_x.py:79-133,
                 noisy.py learn (Huber on a noisy net)
dist_learn()     c51.py:51-122 (variant "c51") and rainbow.py:154-253 (variant "rainbow")
"""
import torch
import torch.nn.functional as F

from . import nets


def _qnet(kind):
    return {"dqn": nets.discrete_q_network, "dueling": nets.dueling}[kind]


def _opt(params, optim):
    name = optim.get("name", "adam")
    kw = {k: v for k, v in optim.items() if k != "name"}
    if name == "adam":
        return torch.optim.Adam(params, **kw)
    if name == "rmsprop":
        return torch.optim.RMSprop(params, **kw)
    raise ValueError(name)


def td_learn(params, target_params, batch, hp, optim, opt_state=None):
    """hp: net ("dqn"|"dueling"|"noisy"), gamma, n_step, double (bool), loss ("huber"|"wmse"), order
    ("dqn"|"double"|"nstep"), alpha, clip (float|None), noise (None | [noise_s, noise_next, noise_target])."""
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    opt = _opt(list(p.values()), optim)
    if opt_state is not None:             # multi-step CPU baselines carry Adam / RMSprop state across learn() calls
        opt.load_state_dict(opt_state)
    state, action, reward = batch["state"], batch["action"], batch["reward"]
    next_state, done = batch["next_state"], batch["done"]
    A = hp["action_size"]
    noise = hp.get("noise") or [None, None, None]

    def fwd(pp, x, nz):
        if hp["net"] == "noisy":
            return nets.noisy_network(pp, x, nz)
        return _qnet(hp["net"])(pp, x)

    eye = torch.eye(A)
    one_hot = eye[action.view(-1).long()]
    q = (fwd(p, state, noise[0]) * one_hot).sum(1, keepdims=True)
    with torch.no_grad():
        max_Q = torch.max(q).item()
        if hp["double"]:
            next_q = fwd(p, next_state, noise[1])
            max_a = torch.argmax(next_q, axis=1)
            boot = (fwd(target_params, next_state, noise[2]) * eye[max_a.long()]).sum(1, keepdims=True)
        else:
            boot = fwd(target_params, next_state, noise[2]).max(1, keepdims=True).values
        if hp["order"] == "dqn":
            target_q = reward + (1 - done) * hp["gamma"] * boot
        elif hp["order"] == "double":
            target_q = reward + boot * (hp["gamma"] * (1 - done))
        else:
            target_q = boot
            for i in reversed(range(hp["n_step"])):
                target_q = reward[:, i] + (1 - done[:, i]) * hp["gamma"] * target_q
    out = {}
    if hp["loss"] == "huber":
        loss = F.smooth_l1_loss(q, target_q)
    else:
        td_error = abs(target_q - q)
        out["priority"] = torch.pow(td_error, hp["alpha"]).detach().view(-1).double()
        w = torch.unsqueeze(torch.FloatTensor(batch["weights"]), -1)
        loss = (w * (td_error ** 2)).mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    out["grads"] = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    if hp.get("clip"):
        torch.nn.utils.clip_grad_norm_(list(p.values()), hp["clip"])
    opt.step()
    out["params"] = {k: v.detach().clone() for k, v in p.items()}
    out["loss"], out["max_Q"] = loss.item(), max_Q
    out["opt_state"] = opt.state_dict()
    return out


def _logits2Q(logits, A, K, z, subtract_max):
    _l = logits.view(logits.shape[0], A, K)
    if subtract_max:                       # c51.py:127-130
        _l = _l - torch.max(_l, -1, keepdim=True).values
    p = torch.exp(F.log_softmax(_l, dim=-1))
    q = torch.sum(z.expand(p.shape[0], A, K) * p, dim=-1)
    return p, q


def dist_learn(params, target_params, batch, hp, optim, opt_state=None):
    """hp: variant ("c51"|"rainbow"), action_size, num_support, v_min, v_max, gamma, n_step, alpha, noise."""
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    opt = _opt(list(p.values()), optim)
    if opt_state is not None:             # multi-step CPU baselines carry Adam / RMSprop state across learn() calls
        opt.load_state_dict(opt_state)
    A, K = hp["action_size"], hp["num_support"]
    v_min, v_max = hp["v_min"], hp["v_max"]
    delta_z = (v_max - v_min) / (K - 1)
    z = torch.linspace(v_min, v_max, K).view(1, -1)
    rainbow = hp["variant"] == "rainbow"
    noise = hp.get("noise") or [None, None, None]
    state, action, reward = batch["state"], batch["action"], batch["reward"]
    next_state, done = batch["next_state"], batch["done"]
    B = state.shape[0]

    def fwd(pp, x, nz):
        if rainbow:
            return nets.rainbow_network(pp, x, nz, A, K)
        return nets.discrete_q_network(pp, x)

    logit = fwd(p, state, noise[0])
    p_logit, q_action = _logits2Q(logit, A, K, z, not rainbow)
    action_eye = torch.eye(A)
    p_action = torch.squeeze(action_eye[action.long()] @ p_logit, 1)
    target_dist = torch.zeros(B, K)
    with torch.no_grad():
        if rainbow:
            _, next_q_action = _logits2Q(fwd(p, next_state, noise[1]), A, K, z, False)
            target_p_logit, _ = _logits2Q(fwd(target_params, next_state, noise[2]), A, K, z, False)
            target_action = torch.argmax(next_q_action, -1, keepdim=True)
        else:
            target_p_logit, target_q_action = _logits2Q(fwd(target_params, next_state, None), A, K, z, True)
            target_action = torch.argmax(target_q_action, -1, keepdim=True)
        target_p_action = torch.squeeze(action_eye[target_action.long()] @ target_p_logit, 1)
        if rainbow:
            Tz = z
            for i in reversed(range(hp["n_step"])):
                Tz = reward[:, i].expand(-1, K) + (1 - done[:, i]) * hp["gamma"] * Tz
            done0 = done[:, 0, :]
        else:
            Tz = reward.expand(-1, K) + (1 - done) * hp["gamma"] * z
            done0 = done
        b = torch.clamp(Tz - v_min, 0, v_max - v_min) / delta_z
        l = torch.floor(b).long()
        u = torch.ceil(b).long()
        support_eye = torch.eye(K)
        l_oh, u_oh = support_eye[l], support_eye[u]
        lluu = l_oh * torch.unsqueeze(u - b, -1) + u_oh * torch.unsqueeze(b - l, -1)
        target_dist += done0 * torch.mean(l_oh * u_oh + lluu, 1)
        target_dist += (1 - done0) * torch.sum(torch.unsqueeze(target_p_action, -1) * lluu, 1)
        target_dist /= torch.clamp(torch.sum(target_dist, 1, keepdim=True), min=1e-8)
    out = {"max_Q": torch.max(q_action).item(), "max_logit": torch.max(logit).item(), "min_logit": torch.min(logit).item()}
    KL = -(target_dist * torch.clamp(p_action, min=1e-8).log()).sum(-1)
    if rainbow:
        out["priority"] = torch.pow(KL, hp["alpha"]).detach().double()
        w = torch.unsqueeze(torch.FloatTensor(batch["weights"]), -1)
        loss = (w * KL).mean()
    else:
        loss = KL.mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    out["grads"] = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    opt.step()
    out["params"] = {k: v.detach().clone() for k, v in p.items()}
    out["loss"] = loss.item()
    out["KL"] = KL.detach().clone()
    out["target_dist"] = target_dist.clone()
    out["opt_state"] = opt.state_dict()
    return out
