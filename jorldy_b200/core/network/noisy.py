"""NoisyNet networks: `Noisy` (jorldy/core/network/noisy.py:8-50) and the Rainbow network
(jorldy/core/network/rainbow.py:8-94: head -> l -> noisy dueling streams over N_atom atoms).

Noisy tensors keep the reference layout (in, out) and parameter order (the direct nn.Parameters
precede the sub-modules in state_dict()).  Every forward draws fresh factor
noise (utils.py:59-68) through jb_noisy_make — Philox stream = layer index, device-side draw
counter — or takes injected normals `noise=[(eps_i, eps_j), ...]` for parity tests.
"""
import torch

from ..dev import C, ptr, stream_ptr
from .base import FlatNetwork, init_gain, orthogonal_
from .head import make_head
from . import layers as L


def _noisy_specs(tag, in_f, out_f):
    return [(f"mu_w{tag}", (in_f, out_f)), (f"sig_w{tag}", (in_f, out_f)), (f"mu_b{tag}", (out_f,)),
            (f"sig_b{tag}", (out_f,))]


def _noisy_init(p, tag, in_f, out_f, noise_type, gen):
    """utils.py:89-105 init_weights."""
    if noise_type == "factorized":
        mu_init, sig_init = 1.0 / (in_f ** 0.5), 0.5 / (in_f ** 0.5)
    else:
        mu_init, sig_init = (3.0 / in_f) ** 0.5, 0.017
    p[f"mu_w{tag}"].copy_((torch.rand((in_f, out_f), generator=gen) * 2 - 1) * mu_init)
    p[f"mu_b{tag}"].copy_((torch.rand((out_f,), generator=gen) * 2 - 1) * mu_init)
    p[f"sig_w{tag}"].fill_(sig_init)
    p[f"sig_b{tag}"].fill_(sig_init)


class _NoisyMixin:
    def _noisy_setup(self, noise_type, seed):
        if noise_type != "factorized":
            raise NotImplementedError("hot-path configs use factorized noise (config/noisy, config/rainbow)")
        self.noise_type = noise_type
        self.noise_seed = int(seed) if seed is not None else 0
        self._draw_ctr = torch.zeros(1, dtype=torch.int64, device=self.device)

    def _noisy_fwd(self, x, tag, lt, layer_id, in_f, out_f, y, relu, is_train, noise):
        """y = act(x @ (mu + sig*eps_w) + (mu_b + sig_b*eps_b)); keeps W/f vectors under `tag+lt`."""
        p = self.p
        fi = self._buf(tag + lt + ".fi", (in_f,)); fj = self._buf(tag + lt + ".fj", (out_f,))
        w = self._buf(tag + lt + ".w", (in_f, out_f)); b = self._buf(tag + lt + ".b", (out_f,))
        ei, ej = (noise if noise is not None else (None, None))
        C.jb_noisy_make(ptr(p[f"mu_w{lt}"]), ptr(p[f"sig_w{lt}"]), ptr(p[f"mu_b{lt}"]), ptr(p[f"sig_b{lt}"]), in_f, out_f,
                        ptr(ei), ptr(ej), self.noise_seed, layer_id, ptr(self._draw_ctr), int(is_train), ptr(fi), ptr(fj),
                        ptr(w), ptr(b), stream_ptr())
        L.linear_io_fwd(x, w, b, y, relu=relu)

    def _noisy_bwd(self, dy, x, tag, lt, in_f, out_f, dx, relu_act):
        """Gradients of one noisy layer: fills g[mu/sig], returns dx (masked by relu_act>0) if dx given."""
        g = self.g
        w = self._buf(tag + lt + ".w", (in_f, out_f))
        fi = self._buf(tag + lt + ".fi", (in_f,)); fj = self._buf(tag + lt + ".fj", (out_f,))
        dw = self._buf(tag + lt + ".dw", (in_f, out_f)); db = self._buf(tag + lt + ".db", (out_f,))
        L.linear_io_bwd_dw(dy, x, dw, db)
        C.jb_noisy_grad(ptr(dw), ptr(db), ptr(fi), ptr(fj), in_f, out_f, ptr(g[f"mu_w{lt}"]), ptr(g[f"sig_w{lt}"]),
                        ptr(g[f"mu_b{lt}"]), ptr(g[f"sig_b{lt}"]), stream_ptr())
        if dx is not None:
            L.linear_io_bwd_dx(dy, w, dx, relu_act=relu_act)


class Noisy(FlatNetwork, _NoisyMixin):
    def __init__(self, D_in, D_out, noise_type="factorized", D_hidden=512, head="mlp", device=None, seed=None):
        super().__init__(device)
        assert noise_type in ["independent", "factorized"]
        self._noisy_setup(noise_type, seed)
        self.D_in, self.D_out, self.D_hidden = D_in, D_out, D_hidden
        self.head = make_head(head, D_in, D_hidden)
        F = self.head.D_head_out
        self._specs = _noisy_specs("1", F, D_hidden) + _noisy_specs("2", D_hidden, D_out) + self.head.specs()
        self._allocate()
        self.nout = D_out
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        with torch.no_grad():
            self.head.init(self.p, gen)
            _noisy_init(self.p, "1", F, D_hidden, noise_type, gen)
            _noisy_init(self.p, "2", D_hidden, D_out, noise_type, gen)

    def forward(self, x, is_train=True, idx=None, M=None, out=None, tag="t.", save=True, noise=None):
        M = M if M is not None else (idx.shape[0] if idx is not None else x.shape[0])
        F, H, A = self.head.D_head_out, self.D_hidden, self.D_out
        n1, n2 = noise if noise is not None else (None, None)
        feat = self.head.forward(self, x, idx, M, tag, save)
        h = self._buf(tag + "h", (M, H))
        self._noisy_fwd(feat, tag, "1", 1, F, H, h, True, is_train, n1)
        if out is None:
            out = self._buf(tag + "q", (M, A))
        self._noisy_fwd(h, tag, "2", 2, H, A, out, False, is_train, n2)
        return out

    def forward_rows(self, x, out, is_train=True, noise=None):
        M = x.shape[0]
        for s in range(0, M, self.head.max_rows):
            e = min(M, s + self.head.max_rows)
            self.forward(x[s:e], is_train, None, e - s, out[s:e], tag=f"inf{e - s}.", save=False, noise=noise)
        return out

    def backward(self, dq, M, tag="t."):
        F, H, A = self.head.D_head_out, self.D_hidden, self.D_out
        feat = self._buf(tag + "head.h", (M, F)); h = self._buf(tag + "h", (M, H))
        dh = self._buf(tag + "dh", (M, H)); dfeat = self._buf(tag + "dfeat", (M, F))
        self._noisy_bwd(dq, h, tag, "2", H, A, dh, h)
        self._noisy_bwd(dh, feat, tag, "1", F, H, dfeat, feat)
        self.head.backward(self, dfeat, M, tag)

    def get_sig_w_mean(self):
        return torch.abs(self.p["sig_w1"]).mean(), torch.abs(self.p["sig_w2"]).mean()


class Rainbow(FlatNetwork, _NoisyMixin):
    def __init__(self, D_in, D_out, N_atom, noise_type="factorized", D_hidden=512, head="mlp", device=None, seed=None):
        super().__init__(device)
        self._noisy_setup(noise_type, seed)
        self.D_in, self.D_out, self.N_atom, self.D_hidden = D_in, D_out, N_atom, D_hidden
        self.head = make_head(head, D_in, D_hidden)
        F, H = self.head.D_head_out, D_hidden
        self._specs = (_noisy_specs("_a1", H, H) + _noisy_specs("_v1", H, H) + _noisy_specs("_a2", H, N_atom * D_out)
                       + _noisy_specs("_v2", H, N_atom) + self.head.specs() + [("l.weight", (H, F)), ("l.bias", (H,))])
        self._allocate()
        self.nout = D_out * N_atom
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        with torch.no_grad():
            self.head.init(self.p, gen)
            self.p["l.weight"].copy_(orthogonal_((H, F), init_gain("relu"), gen))
            _noisy_init(self.p, "_a1", H, H, noise_type, gen)
            _noisy_init(self.p, "_v1", H, H, noise_type, gen)
            _noisy_init(self.p, "_a2", H, N_atom * D_out, noise_type, gen)
            _noisy_init(self.p, "_v2", H, N_atom, noise_type, gen)

    def forward(self, x, is_train=True, idx=None, M=None, out=None, tag="t.", save=True, noise=None):
        """Returns logits [M, A, K].  noise order = the reference's call order: a1, v1, a2, v2."""
        M = M if M is not None else (idx.shape[0] if idx is not None else x.shape[0])
        H, A, K = self.D_hidden, self.D_out, self.N_atom
        na1, nv1, na2, nv2 = noise if noise is not None else (None,) * 4
        feat = self.head.forward(self, x, idx, M, tag, save)
        f = self._buf(tag + "f", (M, H))
        L.linear_fwd(feat, self.p["l.weight"], self.p["l.bias"], f, relu=True)
        xa = self._buf(tag + "xa", (M, H)); xv = self._buf(tag + "xv", (M, H))
        self._noisy_fwd(f, tag, "_a1", 1, H, H, xa, True, is_train, na1)
        self._noisy_fwd(f, tag, "_v1", 2, H, H, xv, True, is_train, nv1)
        a = self._buf(tag + "a", (M, A * K)); v = self._buf(tag + "v", (M, K))
        self._noisy_fwd(xa, tag, "_a2", 3, H, A * K, a, False, is_train, na2)
        self._noisy_fwd(xv, tag, "_v2", 4, H, K, v, False, is_train, nv2)
        if out is None:
            out = self._buf(tag + "logits", (M, A, K))
        C.jb_dueling_fwd(ptr(a), ptr(v), M, A, K, ptr(out), stream_ptr())
        return out

    def forward_rows(self, x, out, is_train=True, noise=None):
        """Chunked inference (act() over many env rows); out [M, A, K].  noise: injected draws (parity tests)."""
        M = x.shape[0]
        for s in range(0, M, self.head.max_rows):
            e = min(M, s + self.head.max_rows)
            self.forward(x[s:e], is_train, None, e - s, out[s:e], tag=f"inf{e - s}.", save=False, noise=noise)
        return out

    def backward(self, dlogits, M, tag="t."):
        H, A, K = self.D_hidden, self.D_out, self.N_atom
        F = self.head.D_head_out
        feat = self._buf(tag + "head.h", (M, F)); f = self._buf(tag + "f", (M, H))
        xa = self._buf(tag + "xa", (M, H)); xv = self._buf(tag + "xv", (M, H))
        da = self._buf(tag + "da", (M, A * K)); dv = self._buf(tag + "dv", (M, K))
        C.jb_dueling_bwd(ptr(dlogits), M, A, K, ptr(da), ptr(dv), stream_ptr())
        dxa = self._buf(tag + "dxa", (M, H)); dxv = self._buf(tag + "dxv", (M, H))
        self._noisy_bwd(da, xa, tag, "_a2", H, A * K, dxa, xa)
        self._noisy_bwd(dv, xv, tag, "_v2", H, K, dxv, xv)
        df = self._buf(tag + "df", (M, H)); df2 = self._buf(tag + "df2", (M, H))
        self._noisy_bwd(dxa, f, tag, "_a1", H, H, df, f)
        self._noisy_bwd(dxv, f, tag, "_v1", H, H, df2, f)
        df.add_(df2)
        L.linear_bwd_dw(df, feat, self.g["l.weight"], self.g["l.bias"])
        dfeat = self._buf(tag + "dfeat", (M, F))
        L.linear_bwd_dx(df, self.p["l.weight"], dfeat, relu_act=feat)
        self.head.backward(self, dfeat, M, tag)
