import sys, json, torch, collections
sys.path.insert(0, '/root/repo')
from textualdegremoval_b200 import define_network, ops
from tools.bench_n3 import NETS
opt = NETS["DRSformer200L_SPA_RefFusion (option 007)"]
net = define_network(dict(opt)).cuda().eval()
g = torch.Generator().manual_seed(0)
lq = torch.rand(4, 3, 512, 512, generator=g).cuda(); ref = torch.rand(4, 3, 512, 512, generator=g).cuda()
with torch.no_grad():
    net(lq, ref); net(lq, ref)
    torch.cuda.synchronize()
    ops.PROF.start()
    net(lq, ref)
    recs = ops.PROF.stop()
fam = collections.defaultdict(lambda: [0, 0.0])
for name, tag, nb, fl, t in recs:
    k = name + ":" + tag
    fam[k][0] += 1; fam[k][1] += t
tot = sum(v[1] for v in fam.values())
print("total ms", tot)
for k, (n, ms) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{k:50s} n={n:4d} ms={ms:8.3f} us={ms/n*1e3:8.1f}")
