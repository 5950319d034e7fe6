"""Functional torch-CPU restatement of the hot-path networks (jorldy/core/network/*).

Parameters are plain dicts keyed exactly like the reference modules' state_dict (SURVEY.md
Appendix B), so the same dict feeds the reference module, this oracle and jorldy_b200.
"""
import torch
import torch.nn.functional as F


def mlp_head(p, x):
    """head.py:6-18 — relu(Linear(D_in, H))."""
    return F.relu(F.linear(x, p["head.l.weight"], p["head.l.bias"]))


def cnn_head(p, x):
    """head.py:21-61 — x/255 then three relu(conv) and flatten."""
    x = x / 255.0
    x = F.relu(F.conv2d(x, p["head.conv1.weight"], p["head.conv1.bias"], stride=4))
    x = F.relu(F.conv2d(x, p["head.conv2.weight"], p["head.conv2.bias"], stride=2))
    x = F.relu(F.conv2d(x, p["head.conv3.weight"], p["head.conv3.bias"], stride=1))
    return x.reshape(x.size(0), -1)


def head(p, x):
    return cnn_head(p, x) if "head.conv1.weight" in p else mlp_head(p, x)


def discrete_policy_value(p, x):
    """policy_value.py:19-22 — returns (pi = exp(log_softmax), v)."""
    h = F.relu(F.linear(head(p, x), p["l.weight"], p["l.bias"]))
    pi = torch.exp(F.log_softmax(F.linear(h, p["pi.weight"], p["pi.bias"]), dim=-1))
    return pi, F.linear(h, p["v.weight"], p["v.bias"])


def continuous_policy_value(p, x):
    """policy_value.py:51-57 — mu clamp(+-5), std = exp(tanh(.)), v."""
    h = F.relu(F.linear(head(p, x), p["l.weight"], p["l.bias"]))
    mu = torch.clamp(F.linear(h, p["mu.weight"], p["mu.bias"]), min=-5.0, max=5.0)
    log_std = torch.tanh(F.linear(h, p["log_std.weight"], p["log_std.bias"]))
    return mu, log_std.exp(), F.linear(h, p["v.weight"], p["v.bias"])


def discrete_q_network(p, x):
    """q_network.py:17-20."""
    h = F.relu(F.linear(head(p, x), p["l.weight"], p["l.bias"]))
    return F.linear(h, p["q.weight"], p["q.bias"])


def dueling(p, x):
    """network/dueling.py:21-35 — Q = V + A - mean_a A."""
    f = head(p, x)
    xa = F.relu(F.linear(f, p["l1_a.weight"], p["l1_a.bias"]))
    xv = F.relu(F.linear(f, p["l1_v.weight"], p["l1_v.bias"]))
    a = F.linear(xa, p["l2_a.weight"], p["l2_a.bias"])
    a = a - a.mean(dim=1, keepdim=True)
    v = F.linear(xv, p["l2_v.weight"], p["l2_v.bias"])
    return a + v


def factorized_noise(eps_i, eps_j):
    """network/utils.py:59-68 — f(e) = sign(e) sqrt|e|; eps_w = f(e_i) f(e_j)^T, eps_b = f(e_j)."""
    f_i = torch.sign(eps_i) * torch.sqrt(torch.abs(eps_i))
    f_j = torch.sign(eps_j) * torch.sqrt(torch.abs(eps_j))
    return torch.matmul(f_i.unsqueeze(1), f_j.unsqueeze(0)), f_j


def noisy_l(x, mu_w, sig_w, mu_b, sig_b, noise):
    """network/utils.py:55-86 — weight layout (in, out), y = x @ W + b.  `noise` is None (eval:
    zeros) or (eps_i, eps_j) raw normal draws for the factorised scheme."""
    if noise is None:
        return torch.matmul(x, mu_w) + mu_b
    eps_w, eps_b = factorized_noise(*noise)
    return torch.matmul(x, mu_w + sig_w * eps_w) + (mu_b + sig_b * eps_b)


def noisy_network(p, x, noise):
    """network/noisy.py:24-50; noise = [(e_i, e_j) for layer 1, layer 2] or None."""
    n1, n2 = (None, None) if noise is None else noise
    h = F.relu(noisy_l(head(p, x), p["mu_w1"], p["sig_w1"], p["mu_b1"], p["sig_b1"], n1))
    return noisy_l(h, p["mu_w2"], p["sig_w2"], p["mu_b2"], p["sig_b2"], n2)


def rainbow_network(p, x, noise, n_action, n_atom):
    """network/rainbow.py:34-94; noise order a1, v1, a2, v2 (the order noisy_l is called in)."""
    na1, nv1, na2, nv2 = (None,) * 4 if noise is None else noise
    f = F.relu(F.linear(head(p, x), p["l.weight"], p["l.bias"]))
    xa = F.relu(noisy_l(f, p["mu_w_a1"], p["sig_w_a1"], p["mu_b_a1"], p["sig_b_a1"], na1))
    xv = F.relu(noisy_l(f, p["mu_w_v1"], p["sig_w_v1"], p["mu_b_v1"], p["sig_b_v1"], nv1))
    xa = noisy_l(xa, p["mu_w_a2"], p["sig_w_a2"], p["mu_b_a2"], p["sig_b_a2"], na2).reshape(-1, n_action, n_atom)
    xa = xa - xa.mean(dim=1, keepdim=True)
    xv = noisy_l(xv, p["mu_w_v2"], p["sig_w_v2"], p["mu_b_v2"], p["sig_b_v2"], nv2).reshape(-1, 1, n_atom)
    return xa + xv
