import sys, json, torch
sys.path.insert(0, "/root/repo")
from textualdegremoval_b200 import define_network, ops
sys.path.insert(0, "/root/repo/tools")
from bench_configs import randomise_gates
torch.manual_seed(0)
opt = dict(type="NAFNetRefFusion", img_channel=3, width=64, middle_blk_num=1, enc_blk_nums=[1, 1, 1, 28],
           dec_blk_nums=[1, 1, 1, 1], nf=64, ext_n_blocks=[4, 4, 4, 4], reffusion_n_blocks=[2, 2, 2, 2, 2])
net = define_network(opt); randomise_gates(net); net = net.cuda().eval()
lq = torch.rand(4, 3, 512, 512).cuda(); ref = torch.rand(4, 3, 512, 512).cuda()
with torch.no_grad():
    for _ in range(3): net(lq, ref)
    torch.cuda.synchronize()
    ops.PROF.start(); net(lq, ref); prof = ops.PROF.stop()
tags = {}
for name, tag, nb, fl, t in prof:
    d = tags.setdefault(f"{name}:{tag}", dict(ms=0.0, n=0, bytes=0, flops=0)); d["ms"] += t; d["n"] += 1; d["bytes"] += nb; d["flops"] += fl
tot = sum(v["ms"] for v in tags.values()); print("total", tot)
fam = {}
for k, v in tags.items():
    f = fam.setdefault(k.split(":")[0], [0.0, 0]); f[0] += v["ms"]; f[1] += v["n"]
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])[:10]: print(f"{k:28s} {v[0]:7.2f} ms n={v[1]}")
for k, v in sorted(tags.items(), key=lambda kv: -kv[1]["ms"])[:22]:
    print(f"{k:52s} {v['ms']:7.3f} ms n={v['n']:3d} {v['bytes']/v['ms']/1e6:8.0f} GB/s {v['flops']/v['ms']/1e9:7.0f} TF")

# ---- training step breakdown
from textualdegremoval_b200.ddp import RefGuidedTrainer
from bench_configs import TRAIN_OPT
net.train()
tr = RefGuidedTrainer(net, TRAIN_OPT)
gt = torch.rand(4, 3, 512, 512).cuda()
tr.feed_train_data(dict(lq=lq, gt=gt, ref_in=ref))
for _ in range(2): tr.optimize_parameters()
torch.cuda.synchronize()
ops.PROF.start(); tr.optimize_parameters(); prof = ops.PROF.stop()
tags = {}
for name, tag, nb, fl, t in prof:
    d = tags.setdefault(f"{name}:{tag}", dict(ms=0.0, n=0, bytes=0, flops=0)); d["ms"] += t; d["n"] += 1; d["bytes"] += nb; d["flops"] += fl
tot = sum(v["ms"] for v in tags.values()); print("train total", tot)
fam = {}
for k, v in tags.items():
    f = fam.setdefault(k.split(":")[0], [0.0, 0]); f[0] += v["ms"]; f[1] += v["n"]
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])[:14]: print(f"{k:28s} {v[0]:7.2f} ms n={v[1]}")
for k, v in sorted(tags.items(), key=lambda kv: -kv[1]["ms"])[:14]:
    print(f"{k:52s} {v['ms']:7.3f} ms n={v['n']:3d} {v['bytes']/v['ms']/1e6:8.0f} GB/s {v['flops']/v['ms']/1e9:7.0f} TF")
