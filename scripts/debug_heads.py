import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jorldy_b200.core.network import layers as L
torch.manual_seed(0)
for (M, H, ns) in [(16, 64, [3]), (256, 512, [2, 1]), (33, 512, [3, 3, 1]), (16, 64, [3, 1])]:
    h = torch.relu(torch.randn(M, H)).cuda()
    ws = [torch.randn(n, H).cuda() for n in ns]; bs = [torch.randn(n).cuda() for n in ns]
    nout = sum(ns)
    out = torch.empty(M, nout, device="cuda")
    L.heads_fwd(h, list(zip(ws, bs)), out)
    ref = torch.cat([h @ w.t() + b for w, b in zip(ws, bs)], 1)
    print("fwd", M, H, ns, (out - ref).abs().max().item())
    dout = torch.randn(M, nout).cuda()
    dws = [torch.zeros_like(w) for w in ws]; dbs = [torch.zeros_like(b) for b in bs]
    L.heads_bwd_dw(dout, h, list(zip(dws, dbs)))
    o = 0
    for n, dw, db in zip(ns, dws, dbs):
        rdw = dout[:, o:o + n].t() @ h; rdb = dout[:, o:o + n].sum(0)
        print("  dw", (dw - rdw).abs().max().item(), "db", (db - rdb).abs().max().item(), db.tolist()[:3], rdb.tolist()[:3])
        o += n
    dh = torch.empty(M, H, device="cuda")
    L.heads_bwd_dx(dout, h, list(zip(ws, bs)), dh)
    rdh = (dout @ torch.cat(ws, 0)) * (h > 0)
    print("  dx", (dh - rdh).abs().max().item())
