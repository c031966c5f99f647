"""Secondary measurements of the other BASELINE.json configs; the headline bench line stays bench.py.

    python tools/bench_configs.py [cfg2] [cfg4] [cfg5]   -> one JSON object per config on stdout

cfg2: Restormer color denoise, 256x256, batch 8 (forward + training step)
cfg4: CLIP ViT-H/14 + I2T Mapper + TR CleanMapper forward, batch 32 (embedding path only)
cfg5: NAFNetRefFusion (width 64, enc [1,1,1,28]) 512x512, batch 4 (forward + training step)
Timing: CUDA events on the launch stream, 3 warm-ups, 256 MB L2 flush between iterations, random-init weights.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from textualdegremoval_b200 import define_network  # noqa: E402
from textualdegremoval_b200.ddp import RefGuidedTrainer  # noqa: E402

DEV = "cuda"
TRAIN_OPT = dict(optim_g=dict(type="AdamW", lr=3e-4, ref_lr=1e-4, weight_decay=1e-4, betas=[0.9, 0.999]),
                 use_grad_clip=True, pixel_opt=dict(type="L1Loss", loss_weight=1.0))


def timed(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


def randomise_gates(net):
    with torch.no_grad():
        for n, p in net.named_parameters():
            leaf = n.rsplit(".", 1)[-1]
            if leaf in ("alpha", "beta", "gamma"):
                p.uniform_(0.1, 1.0)
            elif leaf == "temperature":
                p.uniform_(0.5, 1.5)


def restoration(name, net_opt, B, S, guided, gflop_per_img, flush, iters=5):
    torch.manual_seed(0)
    net = define_network(net_opt)
    randomise_gates(net)
    net = net.to(DEV).eval()
    g = torch.Generator().manual_seed(1)
    lq = torch.rand(B, 3, S, S, generator=g).to(DEV)
    gt = torch.rand(B, 3, S, S, generator=g).to(DEV)
    ref = (gt + 0.02 * torch.randn(B, 3, S, S, generator=g).to(DEV)).clamp(0, 1)
    args = (lq, ref) if guided else (lq,)
    with torch.no_grad():
        ms_f = timed(lambda: net(*args), iters, flush)
    torch.cuda.reset_peak_memory_stats()
    net.train()
    tr = RefGuidedTrainer(net, TRAIN_OPT)
    tr.feed_train_data(dict(lq=lq, gt=gt, ref_in=ref) if guided else dict(lq=lq, gt=gt))
    ms_t = timed(lambda: tr.optimize_parameters(), max(2, iters // 2), flush)
    return dict(config=name, batch=B, size=S, params=sum(p.numel() for p in net.parameters()),
                forward=dict(ms_per_step=ms_f, img_per_s=B / ms_f * 1e3, model_tflops=B * gflop_per_img / ms_f),
                train_step=dict(ms_per_step=ms_t, img_per_s=B / ms_t * 1e3, model_tflops=3 * B * gflop_per_img / ms_t,
                                loss=tr.current_loss(), peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)))


def cfg2(flush):
    opt = dict(type="Restormer", inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4,
               heads=[1, 2, 4, 8], ffn_expansion_factor=2.66, bias=False, LayerNorm_type="BiasFree")
    return restoration("cfg2 Restormer color denoise 256x256 b8", opt, 8, 256, False, 309.76, flush)


def cfg5(flush):
    opt = dict(type="NAFNetRefFusion", img_channel=3, width=64, middle_blk_num=1, enc_blk_nums=[1, 1, 1, 28],
               dec_blk_nums=[1, 1, 1, 1], nf=64, ext_n_blocks=[4, 4, 4, 4], reffusion_n_blocks=[2, 2, 2, 2, 2])
    return restoration("cfg5 NAFNetRefFusion derain 512x512 b4", opt, 4, 512, True, 2653.9, flush, iters=3)


def cfg4(flush):
    from textualdegremoval_b200.archs import vit_b200 as VB
    torch.manual_seed(0)
    clip = VB.CLIPVisionTower().to(DEV).eval()
    mapper = VB.Mapper(1280, 1024, 20).to(DEV).eval()
    clean = VB.CleanMapper(1024, 1024, 20).to(DEV).eval()
    x = torch.randn(32, 3, 224, 224, device=DEV)

    def step():
        h = clip(x, output_hidden_states=True)[0]
        return clean(mapper([h]))

    with torch.no_grad():
        ms = timed(step, 5, flush)
    return dict(config="cfg4 CLIP ViT-H/14 + Mapper + CleanMapper forward b32", batch=32,
                forward=dict(ms_per_step=ms, img_per_s=32 / ms * 1e3, model_tflops=32 * (323.8 + 64.2) / ms))


def main():
    names = sys.argv[1:] or ["cfg2", "cfg4", "cfg5"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for n in names:
        try:
            r = dict(cfg2=cfg2, cfg4=cfg4, cfg5=cfg5)[n](flush)
        except Exception as e:  # noqa: BLE001
            r = dict(config=n, error=f"{type(e).__name__}: {e}"[:500])
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
