"""Flat-parameter network base: every network keeps ONE fp32 parameter buffer and ONE gradient
buffer in HBM (so the optimiser / all-reduce / target copy are single streaming launches) and
exposes torch-style named views for checkpoints.

state_dict keys and shapes follow the reference modules exactly so a
checkpoint written here loads into jorldy/core/network/* and vice versa.
"""
from collections import OrderedDict

import torch

from ..dev import C, ptr, require_cuda, stream_ptr

MAX_ROWS_PER_PASS = 16384   # inference chunk: activations (2 x 32 MB at H=512) stay L2-resident


def orthogonal_(shape, gain, generator=None):
    """Same initialiser family as jorldy/core/network/utils.py:110-125 (torch.nn.init.orthogonal_)."""
    w = torch.empty(shape, dtype=torch.float32)
    torch.nn.init.orthogonal_(w, gain, generator=generator)
    return w


def init_gain(nonlinearity):
    if isinstance(nonlinearity, str):
        if nonlinearity == "policy":
            return 0.01
        return torch.nn.init.calculate_gain(nonlinearity)
    return float(nonlinearity)


class _ParamList(list):
    """network.parameters() result; carries a back-reference so Optimizer(params=...) can reach the
    flat parameter / gradient buffers (reference call shape: reinforce.py:57, dqn.py:78)."""
    network = None


class FlatNetwork:
    """Owns flat params/grads; subclasses declare `self._specs = [(name, shape), ...]` in
    state_dict order before calling `_allocate`."""

    def __init__(self, device=None):
        self.device = require_cuda(device)
        self._specs = []
        self.training = True
        self._ws = {}

    # -- parameter storage ---------------------------------------------------------------------
    def _allocate(self):
        total, offs = 0, []
        for _, shape in self._specs:
            n = 1
            for s in shape:
                n *= s
            # keep every tensor 16-byte aligned inside the flat buffer (float4 paths)
            offs.append(total)
            total += (n + 3) // 4 * 4
        self.num_flat = total
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(total, dtype=torch.float32, device=self.device)
        self._offs = offs
        self.p = OrderedDict()
        self.g = OrderedDict()
        for (name, shape), off in zip(self._specs, offs):
            n = 1
            for s in shape:
                n *= s
            self.p[name] = self.flat[off:off + n].view(*shape)
            self.g[name] = self.grad[off:off + n].view(*shape)

    def rebind_grad(self, storage):
        """Moves the flat gradient buffer into `storage` (a 1-D fp32 CUDA tensor of at least num_flat elements, e.g.
        a peer-mapped symmetric-memory allocation for the in-kernel gradient exchange) and rebuilds the views."""
        assert storage.dtype == torch.float32 and storage.numel() >= self.num_flat and storage.device == self.grad.device
        new = storage[:self.num_flat]
        new.copy_(self.grad)
        self.grad = new
        for (name, shape), off in zip(self._specs, self._offs):
            n = 1
            for s in shape:
                n *= s
            self.g[name] = self.grad[off:off + n].view(*shape)

    def num_params(self):
        return sum(v.numel() for v in self.p.values())

    def parameters(self):
        params = _ParamList(self.p.values())
        params.network = self
        return params

    def state_dict(self):
        return OrderedDict((k, v.detach().clone()) for k, v in self.p.items())

    def load_state_dict(self, sd):
        missing = [k for k in self.p if k not in sd]
        extra = [k for k in sd if k not in self.p]
        if missing or extra:
            raise KeyError(f"state_dict mismatch: missing {missing}, unexpected {extra}")
        with torch.no_grad():
            for k, v in self.p.items():
                v.copy_(torch.as_tensor(sd[k], dtype=torch.float32).to(self.device).view_as(v))

    def copy_from(self, other):
        """Hard copy of another network's flat buffer (target-network update, dqn.py:153-154)."""
        assert other.num_flat == self.num_flat
        C.jb_copy_f32(ptr(self.flat), ptr(other.flat), self.num_flat, stream_ptr())

    def train(self, mode=True):
        self.training = mode
        return self

    def to(self, device):
        return self

    # -- workspaces ----------------------------------------------------------------------------
    def _buf(self, key, shape, dtype=torch.float32):
        k = (key, tuple(shape), dtype)
        t = self._ws.get(k)
        if t is None:
            t = torch.empty(*shape, dtype=dtype, device=self.device)
            self._ws[k] = t
        return t
