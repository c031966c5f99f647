"""GPU-box diagnostic: full-size guided NAFNet fixture -- where do the match mismatches come from (features or search)?"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import restormer as O, nafnet as ON, weights as Wt
from oracle.make_golden_fullsize import fullsize_inputs
from textualdegremoval_b200.archs import define_network

name = sys.argv[1] if len(sys.argv) > 1 else "full_guided_nafnet_512"
z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
meta = json.loads(str(z["meta"]))
lq, rf, _ = fullsize_inputs(meta)
typ = dict(guided_restormer="RestormerRefFusion", guided_nafnet="NAFNetRefFusion")[meta["kind"]]
net = define_network(dict(type=typ, **meta["cfg"]))
sd = Wt.load_seeded(net, meta["seed"])
net = net.cuda().eval()
torch.set_grad_enabled(False)
y, aux = net(lq.cuda(), rf.cuda(), return_aux=True)
t0 = time.time()
fl, fr = O.masa_encoder(sd, O.pad_to(lq, 128)), O.masa_encoder(sd, O.pad_to(rf, 128))
print("oracle encoder", time.time() - t0, "s")
nchw = lambda t: t.float().permute(0, 3, 1, 2).contiguous().cpu()
for i in range(len(fl)):
    a, b = nchw(aux["feat_lq"][i]), fl[i]
    print(f"level {i}: |f| max {b.abs().max():.3e} rms {b.pow(2).mean().sqrt():.3e}  rel-L2 err {((a - b).norm() / b.norm()):.3e} max err {(a - b).abs().max():.3e}")
ds = float(aux["deep_scale"][0]); print("deep level scale", ds)
d32 = nchw(aux["deep32_lq"]) * ds; r32 = nchw(aux["deep32_ref"]) * ds
print(f"deep32 lq rel-L2 {((d32 - fl[-1]).norm() / fl[-1].norm()):.3e}  ref rel-L2 {((r32 - fr[-1]).norm() / fr[-1].norm()):.3e}")
ps = 16 if meta["kind"] == "guided_nafnet" else 8
h, w = O.pad_to(lq, ps * 8).shape[2:]
hr, wr = O.pad_to(rf, ps * 8).shape[2:]
# (1) oracle search on the ORACLE features, (2) oracle search on OUR deep features
_, ax_o = O.masa_warp(fl[-1], fr, ps, 8, 1.5, (1, 2, 3), h, w, hr, wr, return_aux=True)
fr2 = list(fr[:-1]) + [r32]
_, ax_m = O.masa_warp(d32, fr2, ps, 8, 1.5, (1, 2, 3), h, w, hr, wr, return_aux=True)
nq = 64
mine = aux["index"].cpu().long().view(-1, nq); idx_c = aux["idx"].cpu().long()
print("coarse: ours vs oracle(oracle feats)", (idx_c == ax_o["idx"]).float().mean().item(), " ours vs oracle(our feats)", (idx_c == ax_m["idx"]).float().mean().item())
print("fine  : ours vs oracle(oracle feats)", (mine == ax_o["index"].view(-1, nq)).float().mean().item(), " ours vs oracle(our feats)", (mine == ax_m["index"].view(-1, nq)).float().mean().item(),
      " oracle(our feats) vs oracle(oracle feats)", (ax_m["index"] == ax_o["index"]).float().mean().item())
top = ax_o["corr"].topk(2, -1).values
gap = top[..., 0] - top[..., 1]
print("oracle fine top-2 gap quantiles 1/10/50 %:", [float(np.quantile(gap.numpy(), q)) for q in (0.01, 0.1, 0.5)])
sc = aux["corr"]          # ours [nwin, dy, dx, nq]
oc = ax_m["corr"]         # [M, nq, d*d]
ours_c = sc.reshape(sc.shape[0], -1, sc.shape[-1]).permute(0, 2, 1).cpu()
print("corr (ours vs oracle on our feats) max abs diff", (ours_c - oc).abs().max().item(), " att diff", (aux["att"].cpu().view(-1) - ax_m["att"].reshape(-1)).abs().max().item())
