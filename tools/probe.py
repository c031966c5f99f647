"""Run single kernels at the benchmark's shapes (for ncu captures and quick CUDA-event timing).

    python tools/probe.py [name ...]        names: see PROBES
Prints per-probe average time (CUDA events, 20 iterations after 3 warm-ups, 256 MB L2 flush between iterations),
algorithmic GB/s and TFLOP/s.  Under ncu use:  ncu --set full -k regex:<kernel> -s 3 -c 2 python tools/probe.py <name>
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from textualdegremoval_b200 import ops  # noqa: E402

DEV = "cuda"
BF16, F32 = torch.bfloat16, torch.float32


def r16(*s):
    return torch.randn(*s, device=DEV, dtype=F32).to(BF16)


def conv(B, H, W, Ci, Co, k=1, want="bf16", res=False, ld_out=None, ld_in=None, **kw):
    x = r16(B, H, W, ld_in or Ci)[..., :Ci]
    w = r16(k * k, Co, ops.round_up(Ci, 8))
    r = torch.randn(B, H, W, Co, device=DEV) if res else None
    o32 = torch.empty(B, H, W, ld_out or Co, device=DEV)[..., :Co] if want == "f32" else None
    o16 = torch.empty(B, H, W, ld_out or Co, device=DEV, dtype=BF16)[..., :Co] if want == "bf16" else None
    nbytes = B * H * W * (Ci * 2 + Co * (2 if want == "bf16" else 4) + (Co * 4 if res else 0))
    flops = 2 * B * H * W * Ci * Co * k * k
    return (lambda: ops.conv_gemm(x, w, Co, k=k, pad=k // 2, res2=r, out_f32=o32, out_bf16=o16, want=want, **kw)), nbytes, flops


def conv_ln(B, H, W, Ci, Co, batched=False):
    """1x1 conv + fp32 residual in place + fused LayerNorm of the finished rows (the block schedule's attn.v.proj / GDFN
    project_out with norm2 / next norm1 folded in)."""
    x = r16(B, H, W, Ci)
    w = r16(B if batched else 1, Co, ops.round_up(Ci, 8))
    res = torch.randn(B, H, W, Co, device=DEV)
    lw, lb = torch.rand(Co, device=DEV) + 0.5, torch.randn(Co, device=DEV)
    xn = ops.rows16(B, H, W, Co, DEV)
    nbytes = B * H * W * (Ci * 2 + Co * 4 * 2 + Co * 2)
    return (lambda: ops.conv_gemm(x, w, Co, res2=res, out_f32=res, w_batched=batched, ln=(1, lw, lb, 1e-5, xn))), nbytes, \
        2 * B * H * W * Ci * Co


def dw(B, H, W, C_, gate, pitched=False):
    x = r16(B, H, W, ops.round_up(C_, 64) if pitched else C_)[..., :C_]
    w = torch.randn(9, C_, device=DEV)
    co = C_ // 2 if gate else C_
    o = ops.rows16(B, H, W, co, DEV) if pitched else torch.empty(B, H, W, co, device=DEV, dtype=BF16)
    return (lambda: ops.dwconv3x3(x, w, None, gate, out=o)), B * H * W * (C_ + co) * 2, 18 * B * H * W * C_


def ln(B, H, W, C_):
    x = torch.randn(B, H, W, C_, device=DEV)
    w = torch.ones(C_, device=DEV)
    b = torch.zeros(C_, device=DEV)
    o = torch.empty(B, H, W, C_, device=DEV, dtype=BF16)
    return (lambda: ops.rownorm(x, 1, w, b, 1e-5, out=o)), B * H * W * C_ * 6, 0


def gram(B, H, W, C_, heads):
    qkv = r16(B, H, W, 3 * C_)
    t = torch.ones(heads, device=DEV)
    wo = torch.randn(C_, C_, device=DEV)
    return (lambda: ops.mdta_weff(qkv, C_, heads, t, wo)), B * H * W * 2 * C_ * 2, 6 * B * H * W * C_ * (C_ // heads)


def wg(B, H, W, Ci, Co, k=1):
    dy, x = r16(B, H, W, Co), r16(B, H, W, Ci)
    out = torch.zeros(Co, Ci, k, k, device=DEV)
    return (lambda: ops.wgrad(dy, x, out, k=k, pad=k // 2)), B * H * W * (Ci + Co) * 2, 2 * B * H * W * Ci * Co * k * k


def dwwg(B, H, W, C_):
    dy, x = r16(B, H, W, C_), r16(B, H, W, C_)
    dw_, db = torch.zeros(C_, 1, 3, 3, device=DEV), torch.zeros(C_, device=DEV)
    return (lambda: ops.dwconv3x3_wgrad(dy, x, dw_, db)), B * H * W * C_ * 4, 18 * B * H * W * C_


def lnb(B, H, W, C_):
    x, add = torch.randn(B, H, W, C_, device=DEV), torch.randn(B, H, W, C_, device=DEV)
    dy = r16(B, H, W, C_)
    w = torch.ones(C_, device=DEV)
    dwt, dbt = torch.zeros(C_, device=DEV), torch.zeros(C_, device=DEV)
    return (lambda: ops.rownorm_bwd(x, dy, 1, w, 1e-5, add=add, out=add, dweight=dwt, dbias=dbt)), B * H * W * C_ * 14, 0


def gateb(B, H, W, C2):
    y, dg = r16(B, H, W, C2), r16(B, H, W, C2 // 2)
    return (lambda: ops.gate_bwd(y, dg, 1)), B * H * W * C2 * 5, 0


PROBES = {
    "wg_pin96": lambda: wg(4, 512, 512, 96, 512),
    "wg_qkv96": lambda: wg(4, 512, 512, 96, 288),
    "wg_pout256": lambda: wg(4, 512, 512, 256, 96),
    "wg_3x3_48": lambda: wg(8, 512, 512, 48, 48, k=3),
    "wg_3x3_96": lambda: wg(8, 256, 256, 96, 96, k=3),
    "wg_3x3_384": lambda: wg(8, 64, 64, 384, 384, k=3),
    "dwwg512": lambda: dwwg(4, 512, 512, 512),
    "dwwg288": lambda: dwwg(4, 512, 512, 288),
    "dwwg1024": lambda: dwwg(4, 128, 128, 1024),
    "lnb96": lambda: lnb(4, 512, 512, 96),
    "gateb512": lambda: gateb(4, 512, 512, 512),
    "qkv96_ld320": lambda: conv(4, 512, 512, 96, 288, ld_out=320),
    "qkv96_in128": lambda: conv(4, 512, 512, 96, 288, ld_in=128),
    "qkv96_both": lambda: conv(4, 512, 512, 96, 288, ld_out=320, ld_in=128),
    "qkv48_ld192": lambda: conv(4, 512, 512, 48, 144, ld_out=192, ld_in=64),
    "co64": lambda: conv(4, 512, 512, 96, 64),
    "co128": lambda: conv(4, 512, 512, 96, 128),
    "co192": lambda: conv(4, 512, 512, 96, 192),
    "co256": lambda: conv(4, 512, 512, 96, 256),
    "co384": lambda: conv(4, 512, 512, 96, 384),
    "co64k64": lambda: conv(4, 512, 512, 64, 64),
    "co256k64": lambda: conv(4, 512, 512, 64, 256),
    "pin96": lambda: conv(4, 512, 512, 96, 512),
    "qkv96": lambda: conv(4, 512, 512, 96, 288),
    "pout256": lambda: conv(4, 512, 512, 256, 96, want="f32", res=True),
    "qkv48": lambda: conv(4, 512, 512, 48, 144),
    "pout256_ln": lambda: conv_ln(4, 512, 512, 256, 96),
    "pout96wb_ln": lambda: conv_ln(4, 512, 512, 96, 96, batched=True),
    "pout128_ln48": lambda: conv_ln(4, 512, 512, 128, 48),
    "pout96wb": lambda: conv(4, 512, 512, 96, 96, want="f32", res=True),
    "pout128": lambda: conv(4, 512, 512, 128, 48, want="f32", res=True),
    "pin192": lambda: conv(4, 128, 128, 192, 1024),
    "vit_fc1": lambda: conv(1, 1, 8224, 1280, 5120),          # CLIP ViT-H/14 fc1 over 32 x 257 tokens (flat GEMM view)
    "vit_fc2": lambda: conv(1, 1, 8224, 5120, 1280, want="f32", res=True),   # CLIP fc2 + residual (fp32 stream)
    "vit_proj": lambda: conv(1, 1, 8224, 1280, 1280, want="f32", res=True),
    "vit_qkv": lambda: conv(1, 1, 10960, 768, 2304),          # DINOv2 ViT-B/14 qkv over 8 x 1370 tokens
    "c3x3_48": lambda: conv(8, 512, 512, 48, 48, k=3, relu=True),
    "c3x3_96": lambda: conv(8, 256, 256, 96, 96, k=3, relu=True),
    "c3x3_384": lambda: conv(8, 64, 64, 384, 384, k=3, relu=True),
    "c3x3_192": lambda: conv(8, 128, 128, 192, 192, k=3, relu=True),
    "c3x3_256": lambda: conv(8, 128, 128, 256, 256, k=3, relu=True),
    "c3x3_512": lambda: conv(8, 64, 64, 512, 512, k=3, relu=True),
    "c3x3_96s2": lambda: conv(8, 256, 256, 96, 192, k=3, relu=True, stride=2),
    "dwg512": lambda: dw(4, 512, 512, 512, 1),
    "dw288": lambda: dw(4, 512, 512, 288, 0),
    "dw288_ld320": lambda: dw(4, 512, 512, 288, 0, pitched=True),
    "dwg2048": lambda: dw(4, 64, 64, 2048, 1),
    "ln96": lambda: ln(4, 512, 512, 96),
    "ln1280": lambda: ln(1, 1, 8224, 1280),
    "ln1280_32": lambda: ln(1, 1, 32, 1280),
    "ln768": lambda: ln(1, 1, 10960, 768),
    "gram96": lambda: gram(4, 512, 512, 96, 1),
    "gram96_h2": lambda: gram(4, 256, 256, 96, 2),
    "gram192_h4": lambda: gram(4, 128, 128, 192, 4),
    "gram384_h8": lambda: gram(4, 64, 64, 384, 8),
    "gram48": lambda: gram(4, 512, 512, 48, 1),
}


def small_ci(B, H, W, Co, f32=True, f16=True):
    x = torch.rand(B, H, W, 3, device=DEV)
    w = torch.randn(Co, 3, 3, 3, device=DEV) * 0.1
    b = torch.randn(Co, device=DEV) * 0.1
    o32 = torch.empty(B, H, W, Co, device=DEV) if f32 else None
    o16 = torch.empty(B, H, W, Co, device=DEV, dtype=torch.float16) if f16 else None
    nbytes = B * H * W * (12 + Co * ((4 if f32 else 0) + (2 if f16 else 0)))
    return (lambda: ops.conv3x3_small_ci(x, w, b, relu=True, out_f32=o32, out_bf16=o16)), nbytes, 2 * 27 * B * H * W * Co


PROBES["sci_b4"] = lambda: small_ci(4, 512, 512, 48)
PROBES["sci_b8"] = lambda: small_ci(8, 512, 512, 48)
PROBES["sci_b4_f32"] = lambda: small_ci(4, 512, 512, 48, f16=False)


def main():
    names = sys.argv[1:] or list(PROBES)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    # box calibration: plain device copy bandwidth (read + write bytes)
    a = torch.empty(1 << 29, dtype=torch.bfloat16, device=DEV); b_ = torch.empty_like(a)
    for _ in range(2):
        b_.copy_(a)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); b_.copy_(a); e.record(); torch.cuda.synchronize()
    print(f"[box] copy bandwidth {2 * a.numel() * 2 / s.elapsed_time(e) / 1e6:.0f} GB/s", flush=True)
    del a, b_
    for n in names:
        fn, nbytes, flops = PROBES[n]()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        iters = 20
        for _ in range(iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        ms = tot / iters
        print(f"{n:10s} {ms * 1e3:9.1f} us  {nbytes / ms / 1e6:8.1f} GB/s  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
