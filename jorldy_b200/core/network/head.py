"""Input heads (jorldy/core/network/head.py): `mlp` = relu(Linear(D_in, H)) (head.py:6-18) and
`cnn` = /255 -> conv 8x8s4 -> 4x4s2 -> 3x3s1 -> flatten (head.py:21-61).

A head owns no storage: it registers its parameter specs on the owning FlatNetwork and runs
forward/backward on that network's flat views.
"""
import torch

from ..dev import C, ptr, stream_ptr
from .base import init_gain, orthogonal_


class MLPHead:
    kind = "mlp"

    def __init__(self, D_in, D_hidden=512):
        if not isinstance(D_in, int):
            raise ValueError("mlp head expects an integer state_size")
        self.D_in, self.D_head_out = D_in, D_hidden

    def specs(self):
        return [("head.l.weight", (self.D_head_out, self.D_in)), ("head.l.bias", (self.D_head_out,))]

    def init(self, p, gen=None):
        p["head.l.weight"].copy_(orthogonal_((self.D_head_out, self.D_in), init_gain("relu"), gen))
        p["head.l.bias"].zero_()

    def forward(self, net, x, idx, M, tag, save):
        """x: [rows, D_in] f32 (all rows if idx is None else gathered by idx[M] int32)."""
        h = net._buf(tag + "head.h", (M, self.D_head_out))
        xg = net._buf(tag + "head.xg", (M, self.D_in)) if save else None
        C.jb_mlp_in_fwd(ptr(x), ptr(idx), ptr(net.p["head.l.weight"]), ptr(net.p["head.l.bias"]), M, self.D_in,
                        self.D_head_out, ptr(h), ptr(xg), stream_ptr())
        return h

    def backward(self, net, dh_pre, M, tag):
        """dh_pre: gradient w.r.t. the head's pre-activation (already ReLU-masked)."""
        xg = net._buf(tag + "head.xg", (M, self.D_in))
        C.jb_linear_bwd_dw(ptr(dh_pre), ptr(xg), ptr(net.g["head.l.weight"]), ptr(net.g["head.l.bias"]), M,
                           self.D_in, self.D_head_out, stream_ptr())


head_dict = {"mlp": MLPHead}


def make_head(name, D_in, D_hidden):
    if name not in head_dict:
        print(f"### can use only follows {list(head_dict.keys())}")
        raise Exception
    return head_dict[name](D_in, D_hidden)
