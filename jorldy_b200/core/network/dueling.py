"""Dueling Q network (jorldy/core/network/dueling.py:8-35): head -> {l1_a, l1_v} (Linear+ReLU) ->
{l2_a [A], l2_v [1]} -> Q = V + A - mean_a A."""
import torch

from ..dev import C, ptr, stream_ptr
from .base import FlatNetwork, init_gain, orthogonal_
from .head import make_head
from . import layers as L


class Dueling(FlatNetwork):
    def __init__(self, D_in, D_out, D_hidden=512, head="mlp", device=None, seed=None):
        super().__init__(device)
        self.D_in, self.D_out, self.D_hidden = D_in, D_out, D_hidden
        self.head = make_head(head, D_in, D_hidden)
        F = self.head.D_head_out
        self._specs = self.head.specs() + [
            ("l1_a.weight", (D_hidden, F)), ("l1_a.bias", (D_hidden,)),
            ("l1_v.weight", (D_hidden, F)), ("l1_v.bias", (D_hidden,)),
            ("l2_a.weight", (D_out, D_hidden)), ("l2_a.bias", (D_out,)),
            ("l2_v.weight", (1, D_hidden)), ("l2_v.bias", (1,))]
        self._allocate()
        self.nout = D_out
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        with torch.no_grad():
            self.head.init(self.p, gen)
            for n in ("l1_a", "l1_v"):
                self.p[f"{n}.weight"].copy_(orthogonal_((D_hidden, F), init_gain("relu"), gen))
            self.p["l2_a.weight"].copy_(orthogonal_((D_out, D_hidden), init_gain("linear"), gen))
            self.p["l2_v.weight"].copy_(orthogonal_((1, D_hidden), init_gain("linear"), gen))

    def forward(self, x, idx=None, M=None, out=None, tag="t.", save=True):
        M = M if M is not None else (idx.shape[0] if idx is not None else x.shape[0])
        p, H, A = self.p, self.D_hidden, self.D_out
        feat = self.head.forward(self, x, idx, M, tag, save)
        xa = self._buf(tag + "xa", (M, H)); xv = self._buf(tag + "xv", (M, H))
        L.linear_fwd(feat, p["l1_a.weight"], p["l1_a.bias"], xa, relu=True)
        L.linear_fwd(feat, p["l1_v.weight"], p["l1_v.bias"], xv, relu=True)
        a = self._buf(tag + "a", (M, A)); v = self._buf(tag + "v", (M, 1))
        L.heads_fwd(xa, [(p["l2_a.weight"], p["l2_a.bias"])], a)
        L.heads_fwd(xv, [(p["l2_v.weight"], p["l2_v.bias"])], v)
        if out is None:
            out = self._buf(tag + "q", (M, A))
        C.jb_dueling_fwd(ptr(a), ptr(v), M, A, 1, ptr(out), stream_ptr())
        return out

    def forward_rows(self, x, out):
        M = x.shape[0]
        for s in range(0, M, self.head.max_rows):
            e = min(M, s + self.head.max_rows)
            self.forward(x[s:e], None, e - s, out[s:e], tag=f"inf{e - s}.", save=False)
        return out

    def backward(self, dq, M, tag="t."):
        p, g, H, A = self.p, self.g, self.D_hidden, self.D_out
        F = self.head.D_head_out
        feat = self._buf(tag + "head.h", (M, F))
        xa = self._buf(tag + "xa", (M, H)); xv = self._buf(tag + "xv", (M, H))
        da = self._buf(tag + "da", (M, A)); dv = self._buf(tag + "dv", (M, 1))
        C.jb_dueling_bwd(ptr(dq), M, A, 1, ptr(da), ptr(dv), stream_ptr())
        dxa = self._buf(tag + "dxa", (M, H)); dxv = self._buf(tag + "dxv", (M, H))
        L.heads_bwd_dw(da, xa, [(g["l2_a.weight"], g["l2_a.bias"])])
        L.heads_bwd_dx(da, xa, [(p["l2_a.weight"], None)], dxa)
        L.heads_bwd_dw(dv, xv, [(g["l2_v.weight"], g["l2_v.bias"])])
        L.heads_bwd_dx(dv, xv, [(p["l2_v.weight"], None)], dxv)
        L.linear_bwd_dw(dxa, feat, g["l1_a.weight"], g["l1_a.bias"])
        L.linear_bwd_dw(dxv, feat, g["l1_v.weight"], g["l1_v.bias"])
        dfeat = self._buf(tag + "dfeat", (M, F))
        L.linear_bwd_dx(dxa, p["l1_a.weight"], dfeat, relu_act=feat)
        L.linear_bwd_dx(dxv, p["l1_v.weight"], dfeat, relu_act=feat, accumulate=True)
        self.head.backward(self, dfeat, M, tag)
