"""Optimizer factory (jorldy/core/optimizer/__init__.py:4-31 maps snake_case names onto
torch.optim classes).  Only the optimisers the hot-path configs use exist here: `adam`
(every dqn/ppo/rainbow config) and `rmsprop` (config/ape_x/*: centered, eps 1.5e-7).  Each is one
fused launch pair over the network's flat buffers (csrc/optim.cu), with global-norm clipping
(clip_grad_norm_) folded into the same pass.
"""
from collections import OrderedDict

import numpy as np
import torch

from ..dev import C, ptr, stream_ptr


class _FlatOptimizer:
    def __init__(self, params, lr):
        net = getattr(params, "network", None)
        if net is None:
            raise TypeError("jorldy_b200 optimizers take `params=network.parameters()` of a FlatNetwork")
        self.network = net
        self.defaults = {"lr": lr}
        self.param_groups = [{"lr": lr, "params": list(range(len(net.p)))}]
        dev = net.device
        self._lr_dev = torch.tensor([lr], dtype=torch.float32, device=dev)
        self._lr_host = float(lr)
        self._step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._n_partials = C.jb_grad_partials_count(net.num_flat)
        self._partials = torch.zeros(self._n_partials, dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)

    def _sync_lr(self):
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self._lr_dev.fill_(lr)
            self._lr_host = lr

    def zero_grad(self, set_to_none=True):
        pass   # backward kernels overwrite the flat gradient buffer

    def state_tensors(self):
        """Every device tensor one step() mutates besides the parameters (CUDA-graph warm-ups save/restore these)."""
        return [self._step_dev] + [getattr(self, n) for n in self._STATE]

    def _slot_views(self, flat):
        out = []
        for name, v in self.network.p.items():
            off = v.storage_offset()
            out.append(flat[off:off + v.numel()].view_as(v))
        return out


class Adam(_FlatOptimizer):
    _STATE = ("exp_avg", "exp_avg_sq")

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, **kwargs):
        super().__init__(params, lr)
        self.betas, self.eps = betas, eps
        self.exp_avg = torch.zeros_like(self.network.flat)
        self.exp_avg_sq = torch.zeros_like(self.network.flat)

    def step(self, max_norm=None):
        """One update from network.grad; max_norm = clip_grad_norm_ threshold (None: no clipping)."""
        self._sync_lr()
        net, s = self.network, stream_ptr()
        C.jb_grad_sumsq(ptr(net.grad), net.num_flat, ptr(self._partials), ptr(self._step_dev), s)
        C.jb_adam_step(ptr(net.flat), ptr(net.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), net.num_flat,
                       ptr(self._lr_dev), self.betas[0], self.betas[1], self.eps, ptr(self._step_dev),
                       ptr(self._partials), self._n_partials, float(max_norm) if max_norm else 0.0,
                       ptr(self.grad_norm), s)

    def state_dict(self):
        step = float(self._step_dev.item())
        m, v = self._slot_views(self.exp_avg), self._slot_views(self.exp_avg_sq)
        state = {i: {"step": torch.tensor(step), "exp_avg": m[i].clone(), "exp_avg_sq": v[i].clone()}
                 for i in range(len(m))} if step > 0 else {}
        group = {"lr": float(self.param_groups[0]["lr"]), "betas": self.betas, "eps": self.eps, "weight_decay": 0,
                 "amsgrad": False, "params": list(range(len(m)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        self.param_groups[0]["lr"] = float(g["lr"])
        st = sd.get("state", {})
        if st:
            m, v = self._slot_views(self.exp_avg), self._slot_views(self.exp_avg_sq)
            for i in range(len(m)):
                e = st[i] if i in st else st[str(i)]
                m[i].copy_(torch.as_tensor(e["exp_avg"]).to(m[i].device))
                v[i].copy_(torch.as_tensor(e["exp_avg_sq"]).to(v[i].device))
            first = st[0] if 0 in st else st["0"]
            self._step_dev.fill_(int(float(first["step"])))


class RMSprop(_FlatOptimizer):
    _STATE = ("square_avg", "grad_avg")

    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, centered=False, **kwargs):
        super().__init__(params, lr)
        if not centered:
            raise NotImplementedError("only RMSprop(centered=True) is on the hot path (config/ape_x/*.py)")
        self.alpha, self.eps, self.centered = alpha, eps, centered
        self.square_avg = torch.zeros_like(self.network.flat)
        self.grad_avg = torch.zeros_like(self.network.flat)

    def step(self, max_norm=None):
        self._sync_lr()
        net, s = self.network, stream_ptr()
        C.jb_grad_sumsq(ptr(net.grad), net.num_flat, ptr(self._partials), ptr(self._step_dev), s)
        C.jb_rmsprop_centered_step(ptr(net.flat), ptr(net.grad), ptr(self.square_avg), ptr(self.grad_avg),
                                   net.num_flat, ptr(self._lr_dev), self.alpha, self.eps, ptr(self._partials),
                                   self._n_partials, float(max_norm) if max_norm else 0.0, ptr(self.grad_norm), s)

    def state_dict(self):
        step = float(self._step_dev.item())
        sq, ga = self._slot_views(self.square_avg), self._slot_views(self.grad_avg)
        state = {i: {"step": torch.tensor(step), "square_avg": sq[i].clone(), "grad_avg": ga[i].clone()}
                 for i in range(len(sq))} if step > 0 else {}
        group = {"lr": float(self.param_groups[0]["lr"]), "alpha": self.alpha, "eps": self.eps, "centered": True,
                 "momentum": 0, "weight_decay": 0, "params": list(range(len(sq)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        self.param_groups[0]["lr"] = float(sd["param_groups"][0]["lr"])
        st = sd.get("state", {})
        if st:
            sq, ga = self._slot_views(self.square_avg), self._slot_views(self.grad_avg)
            for i in range(len(sq)):
                e = st[i] if i in st else st[str(i)]
                sq[i].copy_(torch.as_tensor(e["square_avg"]).to(sq[i].device))
                ga[i].copy_(torch.as_tensor(e["grad_avg"]).to(ga[i].device))
            first = st[0] if 0 in st else st["0"]
            self._step_dev.fill_(int(float(first["step"])))


optimizer_dict = OrderedDict(adam=Adam, rmsprop=RMSprop)


class Optimizer:
    def __new__(cls, name, *args, **kwargs):
        if type(name) != str:
            print("### name variable must be string! ###")
            raise Exception
        name = name.lower()
        if name not in optimizer_dict.keys():
            print(f"### can use only follows {[opt for opt in optimizer_dict.keys()]}")
            raise Exception
        return optimizer_dict[name](*args, **kwargs)
