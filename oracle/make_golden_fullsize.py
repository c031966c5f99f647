"""TEST INFRASTRUCTURE ONLY -- full-size fixtures from the UNMODIFIED reference, one image each, at the sizes
BASELINE.json's configs name (this container only: needs /root/reference; ~3 minutes of CPU):

  full_guided_restormer_512   RestormerRefFusion, options/train_restoration/003*.yml network_g, 512x512 lq + ref
  full_guided_nafnet_512      NAFNetRefFusion, options/.../002*.yml network_g (5-entry fusion list), 512x512
  full_restormer_256          Restormer colour denoising (options/.../017*.yml network_g: BiasFree), 256x256
  full_dino_vitb_518          models/dino vit_base(img_size=518, patch 14, init_values 1, mlp), 518x518
  full_clip_vith_224          transformers.CLIPVisionModel ViT-H/14 (third party; version recorded), 224x224

Each .npz holds the reference module's fp32 output.  The guided ones also hold the oracle's match bookkeeping (which is
asserted here to reproduce the reference output at full size): the top-3 candidates of every coarse / fine arg-max with
their fp32 scores, so the GPU tests can grade index mismatches by the score gap they jump (ties vs errors).
Weights / inputs are regenerated from seeds (oracle.weights), never stored.

    python -m oracle.make_golden_fullsize [name ...]
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import ref_loader as R
from . import weights as W

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

OPTION_003 = dict(inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4,
                  heads=[1, 2, 4, 8], ffn_expansion_factor=2.66, bias=False, LayerNorm_type="WithBias",
                  dual_pixel_task=False, nf=48, ext_n_blocks=[4, 4, 4, 4], reffusion_n_blocks=[2, 2, 2, 2],
                  reffusion_n_blocks_middle=1, scale=1, num_nbr=1, psize=3, lr_block_size=8, ref_down_block_size=1.5,
                  dilations=[1, 2, 3])
OPTION_002 = dict(img_channel=3, width=64, middle_blk_num=1, enc_blk_nums=[1, 1, 1, 28], dec_blk_nums=[1, 1, 1, 1],
                  nf=64, ext_n_blocks=[4, 4, 4, 4], reffusion_n_blocks=[2, 2, 2, 2, 2], reffusion_n_blocks_middle=1, scale=1,
                  num_nbr=1, psize=3, lr_block_size=8, ref_down_block_size=1.5, dilations=[1, 2, 3])
OPTION_017 = dict(inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4,
                  heads=[1, 2, 4, 8], ffn_expansion_factor=2.66, bias=False, LayerNorm_type="BiasFree",
                  dual_pixel_task=False)
DINO_VITB = dict(img_size=518, patch_size=14, init_values=1.0, ffn_layer="mlp", block_chunks=0)
CLIP_VITH = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16, patch_size=14,
                 image_size=224, hidden_act="gelu")

FULL_CASES = {
    "full_guided_restormer_512": dict(kind="guided_restormer", cfg=OPTION_003, seed=61, size=512),
    "full_guided_nafnet_512": dict(kind="guided_nafnet", cfg=OPTION_002, seed=62, size=512),
    "full_restormer_256": dict(kind="restormer", cfg=OPTION_017, seed=63, size=256, sigma=25),
    "full_dino_vitb_518": dict(kind="dino", cfg=DINO_VITB, seed=64, size=518),
    "full_clip_vith_224": dict(kind="clip", cfg=CLIP_VITH, seed=65, size=224),
}


def fullsize_inputs(case):
    """Deterministic image-like inputs: gt = smoothed seeded noise; guided: lq = 9x9 box blur of gt,
    ref = gt shifted by (5, -7) + 2 % noise; denoising: lq = gt + N(0, sigma/255) from the reference's test-noise recipe
    (np.random.seed(0), data/restoration_dataset.py:479-480); ViTs: the smoothed image, CLIP-normalised for CLIP."""
    n, seed = case["size"], case["seed"]
    raw = W.seeded_image("full", (1, 3, n + 8, n + 8), seed)
    gt = F.avg_pool2d(raw, 5, 1, 2)[..., 4:-4, 4:-4].contiguous()
    gt = ((gt - gt.mean()) * 3 + 0.5).clamp(0, 1)
    kind = case["kind"]
    if kind in ("guided_restormer", "guided_nafnet"):
        lq = F.avg_pool2d(F.pad(gt, (4, 4, 4, 4), mode="reflect"), 9, 1, 0).contiguous()
        ref = (torch.roll(gt, (5, -7), (2, 3)) + 0.02 * (W.seeded_image("n", gt.shape, seed) - 0.5)).clamp(0, 1)
        return lq, ref.contiguous(), gt
    if kind == "restormer":
        np.random.seed(0)
        noise = np.random.normal(0, case["sigma"] / 255.0, tuple(gt.shape)).astype(np.float32)
        return gt + torch.from_numpy(noise), None, gt
    if kind == "clip":
        mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)     # guidance_generation_dataset.py:168-169
        std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
        return (gt - mean) / std, None, gt
    return gt, None, gt


def _top3(t):
    v, i = t.topk(3, -1)
    return v.numpy().astype(np.float32), i.numpy().astype(np.int32)


def _match_book(aux):
    cv, ci = _top3(aux["score"])
    fv, fi = _top3(aux["corr"])
    return dict(coarse_top_val=cv, coarse_top_idx=ci, fine_top_val=fv, fine_top_idx=fi,
                y1=aux["y1"].numpy().astype(np.int32), x1=aux["x1"].numpy().astype(np.int32))


def build(name):
    case = FULL_CASES[name]
    kind = case["kind"]
    t0 = time.time()
    extra = {}
    meta = dict(case)
    lq, ref, _ = fullsize_inputs(case)
    if kind == "guided_restormer":
        from . import restormer as O
        net = R.restormer_ref_fusion(**case["cfg"])
        sd = W.load_seeded(net, case["seed"])
        y = net(lq, ref)
        yo, aux = O.restormer_ref_fusion_forward(sd, lq, ref, case["cfg"]["heads"], return_aux=True)
        extra = _match_book(aux)
        meta["oracle_vs_reference_max"] = float((y - yo).abs().max())
    elif kind == "guided_nafnet":
        from . import nafnet as ON
        net = R.nafnet_ref_fusion(**case["cfg"])
        sd = W.load_seeded(net, case["seed"])
        y = net(lq, ref)
        yo, aux = ON.nafnet_ref_fusion_forward(sd, lq, ref, return_aux=True)
        extra = _match_book(aux)
        meta["oracle_vs_reference_max"] = float((y - yo).abs().max())
    elif kind == "restormer":
        from . import restormer as O
        net = R.restormer(**case["cfg"])
        sd = W.load_seeded(net, case["seed"])
        y = net(lq)
        meta["oracle_vs_reference_max"] = float((y - O.restormer_forward(sd, lq, case["cfg"]["heads"])).abs().max())
    elif kind == "dino":
        from . import vit as OV
        R._stub_packages()
        from models.dino.vision_transformers import vit_base
        net = vit_base(**case["cfg"]).eval()
        sd = W.load_seeded(net, case["seed"])
        y = net(lq)
        meta["oracle_vs_reference_max"] = float((y - OV.dino_vit_forward(sd, lq)).abs().max())
    else:
        import transformers
        from . import vit as OV
        net = transformers.CLIPVisionModel(transformers.CLIPVisionConfig(**case["cfg"])).eval()
        sd = W.load_seeded(net, case["seed"])
        y = net(lq, output_hidden_states=True)[0]
        meta["transformers"] = transformers.__version__
        meta["oracle_vs_reference_max"] = float((y - OV.clip_vision_forward(sd, lq, case["cfg"]["num_attention_heads"], case["cfg"]["patch_size"])).abs().max())
    assert meta["oracle_vs_reference_max"] < 2e-4, (name, meta["oracle_vs_reference_max"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), out=y.numpy().astype(np.float32), **extra)
    print(f"{name}: out {tuple(y.shape)} |out|max {float(y.abs().max()):.3f}  oracle vs reference max "
          f"{meta['oracle_vs_reference_max']:.2e}  ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    os.makedirs(OUT, exist_ok=True)
    for n in (sys.argv[1:] or list(FULL_CASES)):
        build(n)
