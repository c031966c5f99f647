"""Forward timing of the SURVEY 8(f) N3 families at their option-file sizes (512x512, batch per argument), CUDA events.

    python tools/bench_n3.py [batch]
Prints one JSON line per network.  These families are parity-first widenings: inference only, no per-kernel tuning.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from textualdegremoval_b200 import define_network  # noqa: E402

COMMON = dict(inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], heads=[1, 2, 4, 8], ffn_expansion_factor=2.66,
              bias=False, LayerNorm_type="WithBias", nf=48, ext_n_blocks=[4, 4, 4, 4], reffusion_n_blocks=[2, 2, 2, 2],
              lr_block_size=8, ref_down_block_size=1.5, dilations=[1, 2, 3])
NETS = {
    "PromptIRRefFusion (option 001 kwargs, decoder=True)": dict(type="PromptIRRefFusion", num_refinement_blocks=4, decoder=True, **COMMON),
    "DRSformer200L_SPA_RefFusion (option 007)": dict(type="DRSformer200L_SPA_RefFusion", **COMMON),
    "DRSformerRefFusion (options 008-010)": dict(type="DRSformerRefFusion", **COMMON),
}


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    g = torch.Generator().manual_seed(0)
    lq = torch.rand(batch, 3, 512, 512, generator=g).cuda()
    ref = torch.rand(batch, 3, 512, 512, generator=g).cuda()
    for name, opt in NETS.items():
        torch.manual_seed(0)
        net = define_network(dict(opt)).cuda().eval()
        with torch.no_grad():
            for p in net.parameters():
                if p.numel() == 1 and "alpha" in [n for n, q in net.named_parameters() if q is p][0]:
                    p.fill_(0.5)
            for _ in range(2):
                net(lq, ref)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            n = 3
            for _ in range(n):
                y = net(lq, ref)
            e.record()
            torch.cuda.synchronize()
        ms = s.elapsed_time(e) / n
        print(json.dumps(dict(net=name, params=sum(p.numel() for p in net.parameters()), batch=batch, size=512,
                              ms_per_step=round(ms, 2), img_per_s=round(batch / ms * 1e3, 2), finite=bool(torch.isfinite(y).all()),
                              peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 1))), flush=True)
        del net
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
