"""State-action critic of DDPG / TD3 / SAC (jorldy/core/network/q_network.py:23-40):

    x1 = head(state)                 (mlp head: relu(Linear(D_in1, H)))
    x2 = relu(e(action))             (Linear(D_in2, H))
    q  = q(relu(l(cat[x1, x2])))     (Linear(2H, H), Linear(H, 1))

The concatenation is never materialised by a copy: both first-layer products write their half of one [M, 2H]
activation buffer (jb_gemm with ldc = 2H).  backward() optionally returns d q / d action, which is what the
actor losses of the three agents differentiate through.
"""
import torch

from ..dev import C, ptr, stream_ptr
from .base import FlatNetwork, init_gain, orthogonal_
from .head import make_head
from . import layers as L


class ContinuousQ_Network(FlatNetwork):
    def __init__(self, D_in1, D_in2, head="mlp", D_hidden=512, device=None, seed=None):
        super().__init__(device)
        if head != "mlp":
            raise NotImplementedError("the state-action critic is built for the mlp head (config/{ddpg,td3,sac}/*.py)")
        self.D_in1, self.D_in2, self.D_hidden = D_in1, D_in2, D_hidden
        self.head = make_head(head, D_in1, D_hidden)
        F, H = self.head.D_head_out, D_hidden
        self._specs = self.head.specs() + [("e.weight", (H, D_in2)), ("e.bias", (H,)),
                                           ("l.weight", (H, H + F)), ("l.bias", (H,)),
                                           ("q.weight", (1, H)), ("q.bias", (1,))]
        self._allocate()
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        with torch.no_grad():
            self.head.init(self.p, gen)
            self.p["e.weight"].copy_(orthogonal_((H, D_in2), init_gain("relu"), gen))
            self.p["l.weight"].copy_(orthogonal_((H, H + F), init_gain("relu"), gen))
            self.p["q.weight"].copy_(orthogonal_((1, H), init_gain("linear"), gen))
        self._saved = {}

    def forward(self, x1, x2, tag="t."):
        """x1 [M, D_in1], x2 [M, D_in2] f32 device tensors -> q [M, 1]."""
        M = x1.shape[0]
        p, H, F = self.p, self.D_hidden, self.head.D_head_out
        x1 = x1.contiguous()
        x2 = x2.contiguous()
        cat = self._buf(tag + "cat", (M, F + H))
        s = stream_ptr()
        C.jb_gemm(ptr(x1), self.D_in1, 1, ptr(p["head.l.weight"]), self.D_in1, 1, ptr(cat), F + H, M, F, self.D_in1,
                  ptr(p["head.l.bias"]), 1, 0, 0, 0, 0, s)
        C.jb_gemm(ptr(x2), self.D_in2, 1, ptr(p["e.weight"]), self.D_in2, 1, cat.data_ptr() + 4 * F, F + H, M, H,
                  self.D_in2, ptr(p["e.bias"]), 1, 0, 0, 0, 0, s)
        h2 = self._buf(tag + "h2", (M, H))
        L.linear_fwd(cat, p["l.weight"], p["l.bias"], h2, relu=True)
        q = self._buf(tag + "q", (M, 1))
        L.heads_fwd(h2, [(p["q.weight"], p["q.bias"])], q)
        self._saved[tag] = (x1, x2)
        return q

    def backward(self, dq, M, tag="t.", params=True, want_dx2=False, dx2=None, accumulate=False):
        """dq [M, 1] = d loss / d q.  params: fill self.grad (overwrites); want_dx2: return d loss / d x2 [M, D_in2]
        (written to `dx2` if given, added to it if `accumulate`)."""
        p, g, H, F = self.p, self.g, self.D_hidden, self.head.D_head_out
        x1, x2 = self._saved[tag]
        cat = self._buf(tag + "cat", (M, F + H))
        h2 = self._buf(tag + "h2", (M, H))
        dh2 = self._buf(tag + "dh2", (M, H))
        dcat = self._buf(tag + "dcat", (M, F + H))
        s = stream_ptr()
        if params:
            L.heads_bwd_dw(dq, h2, [(g["q.weight"], g["q.bias"])])
        L.heads_bwd_dx(dq, h2, [(p["q.weight"], None)], dh2)                      # masked by relu(h2)
        if params:
            L.linear_bwd_dw(dh2, cat, g["l.weight"], g["l.bias"])
        L.linear_bwd_dx(dh2, p["l.weight"], dcat, relu_act=cat)                    # masked by relu(cat) (both halves)
        if params:
            # dW[out, in] = dcat_half^T x ; db = column sums (the jb_linear_bwd_dw product with lda = 2H)
            C.jb_gemm(ptr(dcat), F + H, 0, ptr(x1), self.D_in1, 0, ptr(g["head.l.weight"]), self.D_in1, F, self.D_in1, M,
                      0, 0, 0, 0, ptr(g["head.l.bias"]), 0, s)
            C.jb_gemm(dcat.data_ptr() + 4 * F, F + H, 0, ptr(x2), self.D_in2, 0, ptr(g["e.weight"]), self.D_in2, H,
                      self.D_in2, M, 0, 0, 0, 0, ptr(g["e.bias"]), 0, s)
        if not want_dx2:
            return None
        if dx2 is None:
            dx2 = self._buf(tag + "dx2", (M, self.D_in2))
        C.jb_gemm(dcat.data_ptr() + 4 * F, F + H, 1, ptr(p["e.weight"]), self.D_in2, 0, ptr(dx2), self.D_in2, M,
                  self.D_in2, H, 0, 0, 0, 0, 0, int(accumulate), s)
        return dx2
