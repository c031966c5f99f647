"""CPU study (oracle = test infrastructure): end-to-end output error of the plain Restormer when every GEMM / depthwise
operand and every stored 16-bit intermediate is rounded to bf16 vs IEEE fp16 (fp32 accumulation, fp32 residual stream),
against the fp32 reference fixture.  python -m tools.operand_format_study"""
import json
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import restormer as O, weights as W
from oracle.make_golden_fullsize import fullsize_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def psnr_u8(a, b):
    a8, b8 = (a.clamp(0, 1) * 255).round().double(), (b.clamp(0, 1) * 255).round().double()
    return 20 * np.log10(255.0 / np.sqrt(((a8 - b8) ** 2).mean().item()))


def main():
    torch.set_grad_enabled(False)
    z = np.load(os.path.join(ROOT, "tests", "golden", "full_restormer_256.npz"))
    meta, ref = json.loads(str(z["meta"])), torch.from_numpy(z["out"])
    from textualdegremoval_b200.archs import define_network
    shapes = {k: v.shape for k, v in define_network(dict(type="Restormer", **meta["cfg"])).state_dict().items()}
    sd = W.seeded_state_dict(shapes, meta["seed"])
    x, _, _ = fullsize_inputs(meta)
    conv0, ln0 = O._conv, O._ln
    gt = (ref + 0.05 * (W.seeded_image("gt_noise", ref.shape, meta["seed"]) - 0.5)).clamp(0, 1)
    for fmt in (torch.bfloat16, torch.float16):
        r = lambda t: t.to(fmt).float()

        def conv(sd_, p, x_, **kw):
            y = F.conv2d(r(x_), r(sd_[p + ".weight"]), sd_.get(p + ".bias"), **kw)
            keep32 = p.endswith("project_out") or p in ("output", "patch_embed.proj") or "reduce_chan" in p
            return y if keep32 else r(y)

        O._conv = conv
        O._ln = lambda sd_, p, x_: r(ln0(sd_, p, x_))
        y = O.restormer_forward(sd, x, meta["cfg"]["heads"])
        O._conv, O._ln = conv0, ln0
        d = (y - ref).abs()
        print(f"{fmt}: max {d.max().item():.2e} mean {d.mean().item():.2e} psnr_u8 {psnr_u8(y, ref):.2f} dB  "
              f"psnr-delta {abs(psnr_u8(y, gt) - psnr_u8(ref, gt)):.4f} dB", flush=True)


if __name__ == "__main__":
    main()
