"""Per-kernel and per-model parity checks of the CUDA path (through the C ABI) against the fp32 CPU oracle.

Used two ways:
  * ``tests/test_kernels_gpu.py`` / ``tests/test_models_gpu.py`` parametrise over ``CHECKS`` (pytest -m gpu);
  * ``python -m tests.gpu_checks [--isolate] [--only GROUP]`` runs everything, never stops at the first failure, and
    writes ``gpurun_out/gpu_checks.json`` -- one GPU call gives the whole picture.  ``--isolate`` runs each group in
    its own process so that a trapping kernel cannot poison the checks after it.

Reference values are computed on CPU in fp32 from the SAME bf16-rounded operands the kernels see, so the tolerance
only has to cover accumulation order and the rounding of the stored result.
"""
import json
import os
import subprocess
import sys
import time
import traceback

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BF16, F32 = torch.bfloat16, torch.float32
DEV = "cuda"


def _ops():
    from textualdegremoval_b200 import ops
    return ops


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).float()


def q(t):
    """round to bf16 and back (what a bf16 tensor holds)."""
    return t.to(BF16).float()


def nhwc(t):   # NCHW cpu -> NHWC cuda (same dtype)
    return t.permute(0, 2, 3, 1).contiguous().to(DEV)


def nchw(t):   # NHWC cuda -> NCHW cpu fp32
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


def result(name, got, ref, rtol, note=""):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    err = (got - ref).abs().max().item() if got.numel() else 0.0
    scale = max(ref.abs().max().item(), 1e-6)
    finite = bool(torch.isfinite(got).all())
    ok = finite and err <= rtol * scale
    return dict(name=name, max_err=err, ref_scale=scale, tol=rtol * scale, ok=ok, note=note)


# =============================================================================================== pointwise
def check_layout():
    ops = _ops()
    out = []
    x = rnd(2, 3, 20, 28, seed=1)
    y = ops.nchw_to_nhwc(x.to(DEV), 24, 32)
    ref = F.pad(x, (0, 4, 0, 4)).permute(0, 2, 3, 1)
    out.append(result("nchw_to_nhwc_pad", y, ref, 0.0))
    z = ops.nhwc_to_nchw(y, 20, 28)
    out.append(result("nhwc_to_nchw_crop", z, x, 0.0))
    wide = torch.zeros(2, 24, 32, 8, device=DEV)
    wide[..., :3] = y
    z = ops.nhwc_to_nchw(wide[..., :3], 20, 28, res=y)
    out.append(result("nhwc_to_nchw_slice_res", z, 2 * x, 0.0))
    src = rnd(1, 6, 10, 48, seed=2).to(DEV)                      # NHWC
    dst = torch.zeros(1, 6, 10, 96, device=DEV)
    d16 = torch.zeros(1, 6, 10, 96, device=DEV, dtype=BF16)
    ops.copy_rows(src, dst32=dst[..., 48:], dst16=d16[..., :48])
    out.append(result("copy_rows_f32_slice", dst[..., 48:], src, 0.0))
    out.append(result("copy_rows_bf16_slice", d16[..., :48], q(src.cpu()), 0.0))
    out.append(result("copy_rows_untouched", dst[..., :48], torch.zeros(1, 6, 10, 48), 0.0))
    return out


def check_rownorm():
    ops = _ops()
    out = []
    for C_ in (16, 48, 96, 192, 384, 768, 1536):
        x = rnd(2, 5, 7, C_, seed=C_) * 2 + 0.3
        w = 1 + 0.2 * rnd(C_, seed=C_ + 1)
        b = 0.1 * rnd(C_, seed=C_ + 2)
        mu = x.mean(-1, keepdim=True)
        var = ((x - mu) ** 2).mean(-1, keepdim=True)
        refs = {0: x, 1: (x - mu) / torch.sqrt(var + 1e-5) * w + b, 2: x / torch.sqrt(var + 1e-5) * w}
        for mode in (0, 1, 2):
            y = ops.rownorm(x.to(DEV), mode, w.to(DEV), b.to(DEV) if mode == 1 else None, 1e-5)
            out.append(result(f"rownorm_m{mode}_C{C_}", y, refs[mode], 6e-3))
    # strided view in and out
    buf = (rnd(1, 4, 4, 96, seed=9)).to(DEV)
    o = torch.zeros(1, 4, 4, 64, device=DEV, dtype=BF16)
    ops.rownorm(buf[..., 48:], 0, out=o[..., 16:])
    out.append(result("rownorm_strided", o[..., 16:], buf[..., 48:].cpu(), 6e-3))
    return out


def _dw_ref(x, w, b, gate):
    y = F.conv2d(x, w, b, padding=1, groups=x.shape[1])
    if gate:
        a, c = y.chunk(2, 1)
        y = (F.gelu(a) if gate == 1 else a) * c
    return y


def check_dwconv():
    ops = _ops()
    out = []
    for (C_, H, W, gate, bias) in ((48, 9, 13, 0, False), (144, 16, 16, 0, True), (256, 8, 21, 1, False),
                                   (32, 5, 4, 2, True), (1024, 6, 8, 1, True)):
        x = q(rnd(2, C_, H, W, seed=C_ + H))
        w = rnd(C_, 1, 3, 3, seed=3) * 0.3
        b = rnd(C_, seed=4) * 0.1 if bias else None
        ref = _dw_ref(x, w, b, gate)
        y = ops.dwconv3x3(nhwc(x.to(BF16)), ops.pack_dw_weight(w.to(DEV)), b.to(DEV) if bias else None, gate)
        out.append(result(f"dwconv_C{C_}_{H}x{W}_g{gate}", nchw(y), ref, 1e-2))
    # GELU alone: identity stencil, x2 = 1, x1 sweeps every bf16 value in [-9, 9] -> the kernel's gate must equal the
    # fp64 erf GELU (F.gelu default, R:239) to within one bf16 ulp of the result (rounding ties), tails included
    bits = torch.arange(0, 1 << 16, dtype=torch.int32).to(torch.int16).view(torch.bfloat16).float()
    vals = bits[torch.isfinite(bits) & (bits.abs() <= 9.0)]
    n = vals.numel()
    Hh = (n + 63) // 64
    x1 = torch.zeros(Hh * 64)
    x1[:n] = vals
    x = torch.stack([x1.view(1, Hh, 64).expand(32, Hh, 64), torch.ones(32, Hh, 64)], 0).reshape(1, 64, Hh, 64)
    w = torch.zeros(64, 1, 3, 3)
    w[:, 0, 1, 1] = 1.0
    y = ops.dwconv3x3(nhwc(x.to(BF16)), ops.pack_dw_weight(w.to(DEV)), None, 1)
    got = nchw(y).float().cpu()[0, 0].reshape(-1)[:n].double()
    exact = vals.double() * 0.5 * torch.special.erfc(-vals.double() / 2 ** 0.5)
    ulp = torch.maximum(exact.abs(), torch.tensor(2.0 ** -126, dtype=torch.float64))
    ulp = 2.0 ** (torch.floor(torch.log2(ulp)) - 7)
    ulp = torch.clamp(ulp, min=2e-6)             # fp32 evaluations (torch's included) lose the far negative tail
    worst = float(((got - exact).abs() / ulp).max())
    abs_err = float((got - exact.to(BF16).double()).abs().max())
    out.append(dict(name="dwconv_gelu_all_bf16_inputs_ulp", ok=worst <= 1.0, max_err=worst, ref_scale=1.0, tol=1.0,
                    note=f"{n} inputs, worst error in units of max(bf16 ulp of the exact result, 2e-6); max |got - bf16(exact)| = {abs_err:.3e}"))
    return out


def check_small_convs():
    ops = _ops()
    out = []
    for Ci, Co in ((3, 48), (1, 16), (6, 24)):
        x = rnd(2, Ci, 11, 14, seed=Ci)
        w = rnd(Co, Ci, 3, 3, seed=5) * 0.2
        b = rnd(Co, seed=6) * 0.1
        ref = F.relu(F.conv2d(x, w, b, padding=1))
        o32 = torch.zeros(2, 11, 14, 2 * Co, device=DEV)
        o16 = torch.zeros(2, 11, 14, Co, device=DEV, dtype=BF16)
        ops.conv3x3_small_ci(nhwc(x), w.to(DEV), b.to(DEV), relu=True, out_f32=o32[..., Co:], out_bf16=o16)
        out.append(result(f"small_ci_{Ci}to{Co}_f32", nchw(o32[..., Co:]), ref, 1e-5))
        out.append(result(f"small_ci_{Ci}to{Co}_bf16", nchw(o16), ref, 6e-3))
    for Ci, Co in ((96, 3), (32, 1)):
        x = q(rnd(2, Ci, 9, 10, seed=Ci))
        w = rnd(Co, Ci, 3, 3, seed=7) * 0.1
        b = rnd(Co, seed=8) * 0.1
        res = rnd(2, Co, 9, 10, seed=9)
        ref = F.conv2d(x, w, b, padding=1) + res
        y = ops.conv3x3_small_co(nhwc(x.to(BF16)), w.to(DEV), b.to(DEV), nhwc(res))
        out.append(result(f"small_co_{Ci}to{Co}", nchw(y), ref, 1e-4))
    return out


# =============================================================================================== conv_gemm
def _conv_case(name, impl, B=2, H=16, W=16, Ci=48, Co=144, k=1, stride=1, pad=0, dil=1, bias=False, relu=False,
               res2=None, res1=False, alpha=1.0, use_scale_ptr=False, store_mode=0, want="f32", batched=False,
               rowscale=False, gelu=False, tol=2e-3):
    ops = _ops()
    x = q(rnd(B, Ci, H, W, seed=H * W + Ci))
    nb = B if batched else 1
    w = q(rnd(nb, Co, Ci, k, k, seed=Co + k) * (1.0 / (Ci * k * k) ** 0.5))
    b = rnd(Co, seed=11) * 0.2 if bias else None
    ys = []
    for i in range(B):
        ys.append(F.conv2d(x[i:i + 1], w[i if batched else 0], None, stride=stride, padding=pad, dilation=dil))
    y = torch.cat(ys, 0)
    OH, OW = y.shape[2:]
    rs = None
    if rowscale:
        rs = (rnd(B, OH, OW, seed=12).abs() + 0.5)
        y = y * rs.unsqueeze(1)
    if bias:
        y = y + b.view(1, -1, 1, 1)
    if relu:
        y = F.relu(y)
    if gelu:
        y = F.gelu(y)
    g = 0.7 if use_scale_ptr else 1.0
    y = y * alpha * g
    if store_mode == 1:
        y = F.pixel_unshuffle(y, 2)
    elif store_mode == 2:
        y = F.pixel_shuffle(y, 2)
    r1 = r2 = None
    if res1:
        r1 = rnd(*y.shape, seed=13)
        y = y + 0.5 * g * r1
    if res2 is not None:
        r2 = rnd(*y.shape, seed=14)
        if res2 == "bf16":
            r2 = q(r2)
        y = y + r2
    wp = torch.cat([ops.pack_conv_weight(w[i].to(DEV)) for i in range(nb)], 0)
    o32, o16 = ops.conv_gemm(
        nhwc(x.to(BF16)), wp, Co, k=k, stride=stride, pad=pad, dil=dil, bias=b.to(DEV) if bias else None, relu=relu,
        gelu=gelu, rowscale=rs.to(DEV) if rowscale else None, alpha=alpha,
        scale_ptr=torch.tensor([g], device=DEV) if use_scale_ptr else None,
        res1=nhwc(r1) if res1 else None, res1_scale=0.5,
        res2=(nhwc(r2.to(BF16)) if res2 == "bf16" else nhwc(r2)) if res2 is not None else None,
        want=want, store_mode=store_mode, w_batched=batched, impl=impl)
    out = []
    tag = "simt" if impl else "tc"
    if o32 is not None:
        out.append(result(f"conv_{tag}_{name}_f32", nchw(o32), y, tol))
    if o16 is not None:
        out.append(result(f"conv_{tag}_{name}_bf16", nchw(o16), y, 8e-3))
    return out


CONV_CASES = [
    dict(name="1x1_48to144"),
    dict(name="1x1_16to48_bias", Ci=16, Co=48, bias=True),
    dict(name="1x1_96to512_2ntiles", Ci=96, Co=512, H=8, W=24),
    dict(name="1x1_384to384_res2", Ci=384, Co=384, H=8, W=8, res2="f32", want="both"),
    dict(name="1x1_fusion_epilogue", Ci=128, Co=96, bias=True, res1=True, res2="f32", use_scale_ptr=True),
    dict(name="1x1_batched_w", Ci=96, Co=96, batched=True, res2="f32"),
    dict(name="1x1_odd_spatial", Ci=48, Co=48, H=20, W=24, B=1),
    dict(name="1x1_w8", Ci=64, Co=64, H=16, W=8),
    dict(name="1x1_big_k", Ci=1024, Co=256, H=8, W=16, B=1),
    dict(name="flat_gemm_gelu", Ci=768, Co=3072, H=1, W=300, B=2, bias=True, gelu=True, want="bf16"),
    dict(name="flat_gemm_res", Ci=3072, Co=768, H=1, W=1370, B=1, bias=True, res2="f32"),
    dict(name="3x3_48to48_bias_relu", Ci=48, Co=48, k=3, pad=1, bias=True, relu=True, want="bf16"),
    dict(name="3x3_res_bf16", Ci=32, Co=32, k=3, pad=1, bias=True, res2="bf16", want="bf16"),
    dict(name="3x3_stride2", Ci=48, Co=96, k=3, pad=1, stride=2, bias=True, relu=True, H=32, W=32),
    dict(name="3x3_stride2_odd", Ci=16, Co=32, k=3, pad=1, stride=2, H=20, W=40, B=1),
    dict(name="3x3_unshuffle", Ci=48, Co=24, k=3, pad=1, store_mode=1, H=16, W=32),
    dict(name="3x3_unshuffle_co12", Ci=24, Co=12, k=3, pad=1, store_mode=1, H=16, W=16, want="both"),
    dict(name="3x3_shuffle", Ci=96, Co=192, k=3, pad=1, store_mode=2, want="both", H=8, W=16),
    dict(name="1x1_shuffle_skip", Ci=256, Co=512, store_mode=2, res2="f32", H=4, W=4, B=1),
    dict(name="3x3_dil2_rowscale", Ci=128, Co=8, k=3, pad=2, dil=2, rowscale=True, batched=True, res2="f32"),
    dict(name="3x3_dil3", Ci=64, Co=64, k=3, pad=3, dil=3, H=16, W=16),
    dict(name="3x3_valid", Ci=64, Co=64, k=3, pad=0, H=15, W=15, B=3),
    dict(name="2x2_stride2", Ci=64, Co=128, k=2, pad=0, stride=2, bias=True, H=16, W=32),
    # >= 2 pixel tiles per SM: the resident-weight mode of the 1x1 convs (A-only ring, pixel-tile-major items)
    dict(name="1x1_res_96to288", Ci=96, Co=288, H=160, W=176, B=2, want="bf16"),
    dict(name="1x1_res_48to144_bias", Ci=48, Co=144, H=192, W=200, B=1, bias=True, want="bf16"),
    dict(name="1x1_res_96to512", Ci=96, Co=512, H=160, W=160, B=2, want="bf16"),
    dict(name="1x1_res_256to96_res2", Ci=256, Co=96, H=160, W=160, B=2, bias=True, res2="f32"),
    dict(name="1x1_res_192to96_fusion", Ci=192, Co=96, H=176, W=160, B=2, bias=True, res1=True, res2="f32",
         use_scale_ptr=True),
    dict(name="1x1_res_ragged_40to72", Ci=40, Co=72, H=250, W=170, B=1, want="both"),
    # pair mode (two pixel tiles per streamed weight tile: dense 3x3 at C >= 96 with >= 2 items per SM); 323 = odd number
    # of pixel tiles -> the last pair has a ghost tile whose loads are TMA zero fill and whose stores are clipped
    dict(name="3x3_pair_96_odd_tiles", Ci=96, Co=96, k=3, pad=1, H=152, W=272, B=1, bias=True, relu=True, want="bf16"),
    dict(name="3x3_pair_128_res2", Ci=128, Co=128, k=3, pad=1, H=160, W=144, B=2, bias=True, res2="f32"),
    dict(name="3x3_pair_stride2_96to192", Ci=96, Co=192, k=3, pad=1, stride=2, H=304, W=272, B=1, relu=True, want="bf16"),
    dict(name="3x3_pair_384_2ntiles", Ci=384, Co=384, k=3, pad=1, H=64, W=80, B=3, want="bf16"),
]


def check_fp16_path():
    """The fp16-operand instantiations used by the MASA feature encoder (R:100-134): tdr_conv_gemm with in_fp16 / out_fp16
    (TMA and generic epilogues), tdr_cast_rows, tdr_pack_conv_weight_f16, tdr_masa_split3, tdr_wgrad on an fp16 tape
    (converted to bf16 first),
    tdr_relu_mask on fp16 activations -- each against fp32 CPU math on the same fp16-rounded operands."""
    ops = _ops()
    F16 = torch.float16
    qh = lambda t: t.to(F16).float()
    out = []
    for (name, B, H, W, Ci, Co, k, st, relu, res, o16) in (
            ("3x3_48_relu_h", 2, 16, 16, 48, 48, 3, 1, True, False, True),
            ("3x3_96_res32", 1, 16, 24, 96, 96, 3, 1, False, True, False),
            ("3x3_s2_f32", 2, 32, 32, 48, 96, 3, 2, True, False, False),
            ("3x3_192_relu_h", 1, 16, 16, 192, 192, 3, 1, True, False, True),
            ("3x3_s2_h", 1, 16, 16, 16, 32, 3, 2, True, False, True)):
        x = qh(rnd(B, Ci, H, W, seed=Ci + H))
        w = rnd(Co, Ci, k, k, seed=Co) * (1.0 / (Ci * k * k) ** 0.5)
        b = rnd(Co, seed=3) * 0.2
        wp = ops.pack_conv_weight_f16(w.to(DEV))
        y = F.conv2d(x, qh(w), b, stride=st, padding=1)
        if relu:
            y = F.relu(y)
        r2 = None
        if res:
            r2 = rnd(*y.shape, seed=14)
            y = y + r2
        o32, oh = ops.conv_gemm(nhwc(x.to(F16)), wp, Co, k=k, stride=st, pad=1, bias=b.to(DEV), relu=relu,
                                res2=nhwc(r2) if res else None, want="bf16" if o16 else "f32", out_fp16=o16)
        if o16:
            assert oh.dtype == F16
            out.append(result(f"conv_fp16_{name}", nchw(oh), y, 1.5e-3))
        else:
            out.append(result(f"conv_fp16_{name}", nchw(o32), y, 2e-4))
    # saturation instead of inf
    x = torch.full((1, 16, 8, 8), 200.0)
    w = torch.full((8, 16, 1, 1), 100.0)
    _, oh = ops.conv_gemm(nhwc(x.to(F16)), ops.pack_conv_weight_f16(w.to(DEV)), 8, out_fp16=True)
    out.append(dict(name="conv_fp16_saturates", max_err=float((oh.float() - 65504.0).abs().max()), ref_scale=65504.0,
                    tol=0.0, ok=bool((oh.float() == 65504.0).all())))
    # cast_rows / split3
    xr = rnd(3, 5, 7, 48, seed=2) * 3
    xb, xh = ops.cast_rows(xr.to(DEV), want_bf16=True, want_fp16=True)
    out.append(result("cast_rows_bf16", xb.float().cpu(), q(xr), 0.0))
    out.append(result("cast_rows_fp16", xh.float().cpu(), qh(xr), 0.0))
    s3 = ops.masa_split3(xr.to(DEV)).float().cpu()
    hi = q(xr)
    out.append(result("split3_hi", s3[..., :48], hi, 0.0))
    out.append(result("split3_lo", s3[..., 48:96], q(xr - hi), 0.0))
    out.append(result("split3_hi2", s3[..., 96:], hi, 0.0))
    out.append(result("split3_sum", s3[..., :48] + s3[..., 48:96], xr, 2.0 ** -16))
    # wgrad with fp16 activations (ops.wgrad converts them to bf16: tolerance = bf16 rounding of x), bf16 dy
    for (B, H, W, Ci, Co, k, st) in ((2, 16, 16, 48, 48, 3, 1), (1, 16, 16, 96, 192, 3, 2), (1, 8, 24, 96, 96, 1, 1)):
        x = qh(rnd(B, Ci, H, W, seed=Ci + H))
        wt = torch.zeros(Co, Ci, k, k, requires_grad=True)
        y = F.conv2d(x, wt, None, stride=st, padding=k // 2)
        dy = q(rnd(*y.shape, seed=Co + W))
        (ref,) = torch.autograd.grad(y, wt, dy)
        dw = torch.zeros((Co, Ci, k, k), device=DEV)
        ops.wgrad(nhwc(dy.to(BF16)), nhwc(x.to(F16)), dw, k=k, stride=st, pad=k // 2, accumulate=False)
        out.append(grad_result(f"wgrad_fp16x_Ci{Ci}_Co{Co}_k{k}s{st}", dw, ref, 5e-3))
    # relu_mask reads only sign / zero bits: identical for bf16 and fp16 activations
    yv = rnd(2, 4, 4, 32, seed=9)
    yv[0, 0, 0, :4] = torch.tensor([0.0, -0.0, 1e-7, -1e-7])
    dyv = q(rnd(2, 4, 4, 32, seed=10))
    for dt, nm in ((F16, "fp16"), (BF16, "bf16")):
        y16 = yv.to(dt)
        got = ops.relu_mask(y16.to(DEV), dyv.to(BF16).to(DEV)).float().cpu()
        out.append(result(f"relu_mask_{nm}", got, torch.where(y16.float() > 0, dyv, torch.zeros_like(dyv)), 0.0))
    # ---- the inference forward's fp16 instantiations of the block kernels (ops.operand_dtype) --------------------------
    for (C_, H, W, gate, bias) in ((48, 9, 13, 0, False), (144, 16, 16, 0, True), (256, 8, 21, 1, False),
                                   (32, 5, 4, 2, True), (1024, 6, 8, 1, True)):
        x = qh(rnd(2, C_, H, W, seed=C_ + H))
        w = rnd(C_, 1, 3, 3, seed=3) * 0.3
        b = rnd(C_, seed=4) * 0.1 if bias else None
        y = ops.dwconv3x3(nhwc(x.to(F16)), ops.pack_dw_weight(w.to(DEV)), b.to(DEV) if bias else None, gate)
        assert y.dtype == F16
        out.append(result(f"dwconv_fp16_C{C_}_{H}x{W}_g{gate}", nchw(y), _dw_ref(x, w, b, gate), 1.5e-3))
    from oracle import restormer as O
    xr = rnd(2, 96, 7, 9, seed=21) * 2 + 0.5
    lw, lb = rnd(96, seed=22) * 0.5 + 1.0, rnd(96, seed=23) * 0.3
    for mode in (0, 1, 2):
        got = ops.rownorm(nhwc(xr), mode, lw.to(DEV), lb.to(DEV) if mode == 1 else None, 1e-5, dt=F16)
        ref = xr if mode == 0 else O.layernorm_c(xr, lw, lb if mode == 1 else None)
        out.append(result(f"rownorm_fp16_m{mode}", nchw(got), ref, 6e-4))
    xg = qh(rnd(2, 64, 9, 11, seed=5))
    out.append(result("gate_mul_fp16", nchw(ops.gate_mul(nhwc(xg.to(F16)))), xg[:, :32] * xg[:, 32:], 1e-3))
    # MDTA Gram + fold on fp16 q, k
    for (C_, heads, H, W, B) in ((96, 2, 16, 16, 2), (48, 1, 16, 24, 1)):
        c = C_ // heads
        qkv = qh(rnd(B, 3 * C_, H, W, seed=C_ + heads))
        temp = torch.rand(heads, generator=torch.Generator().manual_seed(1)) + 0.5
        wpo = rnd(C_, C_, seed=2) / C_ ** 0.5
        qq, kk, vv = qkv.view(B, 3, heads, c, H * W).unbind(1)
        qn = qq / qq.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        kn = kk / kk.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        attn = torch.softmax(qn @ kn.transpose(-1, -2) * temp.view(1, heads, 1, 1), -1)
        weff_ref = torch.zeros(B, C_, C_)
        for h in range(heads):
            weff_ref[:, :, h * c:(h + 1) * c] = wpo[:, h * c:(h + 1) * c] @ attn[:, h]
        weff, attn_g = ops.mdta_weff(nhwc(qkv.to(F16)), C_, heads, temp.to(DEV), wpo.to(DEV), want_attn=True)
        assert weff.dtype == F16
        out.append(result(f"mdta_fp16_attn_C{C_}_h{heads}", attn_g, attn, 2e-4))
        out.append(result(f"mdta_fp16_weff_C{C_}_h{heads}", weff[..., :C_], weff_ref, 1e-3))
    # 1x1 conv + residual + fused LayerNorm with fp16 operands and fp16 LN output; PixelShuffle store of an fp16 output
    B, Ci, Co, H, W = 2, 256, 96, 20, 24
    x = qh(rnd(B, Ci, H, W, seed=H + Ci))
    w = rnd(Co, Ci, 1, 1, seed=Co) * (1.0 / Ci ** 0.5)
    r2 = rnd(B, Co, H, W, seed=14) * 2.0 + 0.7
    y = F.conv2d(x, qh(w)) + r2
    lw, lb = rnd(Co, seed=15) * 0.5 + 1.0, rnd(Co, seed=16) * 0.3
    ln_out = ops.rows16(B, H, W, Co, DEV, F16)
    o32, _ = ops.conv_gemm(nhwc(x.to(F16)), ops.pack_conv_weight(w.to(DEV), dt=F16), Co, res2=nhwc(r2), want="f32",
                           ln=(1, lw.to(DEV), lb.to(DEV), 1e-5, ln_out))
    out.append(result("conv_ln_fp16_f32", nchw(o32), y, 2e-4))
    out.append(result("conv_ln_fp16_ln", nchw(ln_out), O.layernorm_c(y, lw, lb), 1.5e-3))
    x = qh(rnd(1, 96, 8, 16, seed=31))
    w = rnd(192, 96, 3, 3, seed=32) * (1.0 / (96 * 9) ** 0.5)
    _, oh = ops.conv_gemm(nhwc(x.to(F16)), ops.pack_conv_weight(w.to(DEV), dt=F16), 192, k=3, pad=1, store_mode=2)
    assert oh.dtype == F16
    out.append(result("conv_fp16_pixel_shuffle", nchw(oh), F.pixel_shuffle(F.conv2d(x, qh(w), padding=1), 2), 1.5e-3))
    return out


def check_gdfn_tail():
    """tdr_gdfn_tail (depthwise 3x3 + GELU gate + project_out + residuals in one kernel) vs fp32 CPU math on the same
    16-bit-rounded operands; the gated intermediate is rounded to the operand format as the kernel does."""
    ops = _ops()
    F16 = torch.float16
    out = []
    cases = [  # B, H, W, hp, C, fusion, dw_bias, out_bias, dtype
        (2, 16, 32, 128, 48, False, False, False, BF16), (1, 24, 40, 256, 96, True, True, True, F16),
        (1, 8, 16, 512, 192, False, True, False, F16), (2, 20, 24, 48, 16, False, True, True, BF16),
        (1, 19, 37, 88, 32, True, False, True, F16), (3, 64, 48, 256, 96, False, False, False, F16),
    ]
    for (B, H, W, hp, Cc, fusion, dwb, ob, dt) in cases:
        rq = (lambda t: t.to(dt).float())
        hid = rq(rnd(B, 2 * hp, H, W, seed=hp + H))
        wdw = rnd(2 * hp, 1, 3, 3, seed=3) * 0.3
        bdw = rnd(2 * hp, seed=4) * 0.1 if dwb else None
        wo = rnd(Cc, hp, 1, 1, seed=5) * (1.0 / hp ** 0.5)
        bo = rnd(Cc, seed=6) * 0.2 if ob else None
        x0 = rnd(B, Cc, H, W, seed=7)
        x1 = rnd(B, Cc, H, W, seed=8) if fusion else None
        alpha = 0.6
        y = F.conv2d(hid, wdw, bdw, padding=1, groups=2 * hp)
        gte = rq(F.gelu(y[:, :hp]) * y[:, hp:])
        ffn = F.conv2d(gte, rq(wo), bo)
        ref = (x1 + ffn) * alpha + x0 if fusion else ffn + x0
        hid_d = nhwc(hid.to(dt))
        w9 = ops.pack_dw_weight(wdw.to(DEV))
        wp = ops.pack_conv_weight(wo.to(DEV), dt=dt)
        assert ops.gdfn_tail_supported(hid_d, wp, Cc)
        x32 = nhwc(x0)
        ops.gdfn_tail(hid_d, w9, bdw.to(DEV) if dwb else None, wp, Cc, bias=bo.to(DEV) if ob else None,
                      scale_ptr=torch.tensor([alpha], device=DEV) if fusion else None, res1=nhwc(x1) if fusion else None,
                      res2=x32, out=x32)
        out.append(result(f"gdfn_tail_B{B}_{H}x{W}_hp{hp}_C{Cc}_f{int(fusion)}_{'fp16' if dt == F16 else 'bf16'}", nchw(x32), ref,
                          2e-3 if dt == F16 else 6e-3))
    return out


def check_conv_ln():
    """1x1 conv + residual with the LayerNorm of the finished rows emitted by the same epilogue (norm1 / norm2 of
    R:318-331 folded into the producing conv): fp32 rows must equal the plain epilogue's, the bf16 LN output must match
    the fp32 LayerNorm of those rows, in place (out == res2) as the block schedule uses it, many tiles per CTA."""
    from oracle import restormer as O
    ops = _ops()
    out = []
    cases = [dict(Ci=96, Co=96, H=16, W=16, mode=1), dict(Ci=256, Co=96, H=20, W=24, mode=2, bias=True),
             dict(Ci=128, Co=48, H=16, W=32, mode=1, bias=True), dict(Ci=48, Co=48, H=9, W=13, mode=2, B=3),
             dict(Ci=96, Co=96, H=16, W=16, mode=1, batched=True, bias=True), dict(Ci=64, Co=72, H=8, W=40, mode=1), dict(Ci=64, Co=96, H=64, W=96, mode=2, B=4, inplace=True),
             dict(Ci=64, Co=64, H=16, W=16, mode=2), dict(Ci=32, Co=8, H=16, W=16, mode=1, B=1),
             dict(Ci=256, Co=96, H=160, W=192, mode=1, bias=True, B=2, inplace=True),
             dict(Ci=96, Co=48, H=128, W=256, mode=2, B=2, inplace=True, batched=True)]
    for c in cases:
        B, Ci, Co, H, W, mode = c.get("B", 2), c["Ci"], c["Co"], c["H"], c["W"], c["mode"]
        batched = c.get("batched", False)
        x = q(rnd(B, Ci, H, W, seed=H + Ci))
        nb = B if batched else 1
        w = q(rnd(nb, Co, Ci, 1, 1, seed=Co) * (1.0 / Ci ** 0.5))
        b = rnd(Co, seed=11) * 0.2 if c.get("bias") else None
        r2 = rnd(B, Co, H, W, seed=14) * 2.0 + 0.7                 # non-zero channel mean: WithBias must subtract it
        y = torch.cat([F.conv2d(x[i:i + 1], w[i if batched else 0], b) for i in range(B)], 0) + r2
        lw, lb = rnd(Co, seed=15) * 0.5 + 1.0, rnd(Co, seed=16) * 0.3
        ref_ln = O.layernorm_c(y, lw, lb if mode == 1 else None)
        wp = torch.cat([ops.pack_conv_weight(w[i].to(DEV)) for i in range(nb)], 0)
        res = nhwc(r2)
        ln_out = ops.rows16(B, H, W, Co, DEV)
        ln_out.fill_(float("nan"))
        o32 = res if c.get("inplace") else None
        o32, _ = ops.conv_gemm(nhwc(x.to(BF16)), wp, Co, bias=b.to(DEV) if b is not None else None, res2=res,
                               out_f32=o32, want="f32", w_batched=batched,
                               ln=(mode, lw.to(DEV), lb.to(DEV), 1e-5, ln_out))
        name = f"Ci{Ci}_Co{Co}_{H}x{W}_m{mode}" + ("_wb" if batched else "") + ("_inplace" if c.get("inplace") else "")
        out.append(result(f"conv_ln_{name}_f32", nchw(o32), y, 2e-3))
        out.append(result(f"conv_ln_{name}_ln", nchw(ln_out), ref_ln, 1e-2))
        # against the standalone norm kernel on the same fp32 rows: same bf16 values up to rounding of the statistics
        alone = ops.rownorm(o32, mode, lw.to(DEV), lb.to(DEV), 1e-5)
        out.append(result(f"conv_ln_{name}_vs_rownorm", nchw(ln_out), nchw(alone).float(), 8e-3))
    return out


def check_input_pipeline():
    """tdr_prepare_patches vs tensors made by the unmodified reference dataset chain (tests/golden/input_pipeline.npz):
    bit-exact (0 tolerance), all augmentation modes, reflect padding, normalisation, one launch for a whole batch."""
    import numpy as np
    from oracle import input_pipeline as IP
    ops = _ops()
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_pipeline.npz"))
    out = []
    groups = {}
    for k in range(int(z["n"])):
        top, left, mode, size, norm = [int(v) for v in z[f"s{k}_dec"]]
        groups.setdefault((size, norm), []).append((k, top, left, mode))
    for (size, norm), items in groups.items():
        frames, crops, refs = [], [], []
        for k, top, left, mode in items:
            for which in ("gt", "lq"):
                frames.append(torch.from_numpy(z[f"s{k}_{which}_frame"]).to(DEV))
                crops.append((top, left, mode))
                refs.append(torch.from_numpy(z[f"s{k}_{which}"]))
        mean, std = (z["mean"].tolist(), z["std"].tolist()) if norm else (None, None)
        got = ops.prepare_patches(frames, crops, size, mean=mean, std=std)
        out.append(result(f"prepare_patches_size{size}_norm{norm}_n{len(frames)}", got, torch.stack(refs), 0.0))
    # the reference image of a sample is not cropped or augmented (restoration_dataset.py:236): rectangular, mode 0 / flips
    rng = np.random.RandomState(3)
    fr = [rng.randint(0, 256, (37, 53, 3)).astype(np.uint8) for _ in range(3)]
    for mode in (0, 1, 4, 5):
        got = ops.prepare_patches([torch.from_numpy(f).to(DEV) for f in fr], [(0, 0, mode)] * 3, (37, 53))
        ref = torch.stack([torch.from_numpy(IP.prepare_patch(f, 0, 0, mode, (37, 53))) for f in fr])
        out.append(result(f"prepare_patches_full_frame_mode{mode}", got, ref, 0.0))
    gray = rng.randint(0, 256, (20, 20, 1)).astype(np.uint8)
    got = ops.prepare_patches([torch.from_numpy(gray).to(DEV)], [(2, 3, 6)], 16)
    out.append(result("prepare_patches_gray_rot270", got, torch.from_numpy(IP.prepare_patch(gray, 2, 3, 6, 16))[None], 0.0))
    # training noise (data/restoration_dataset.py:474-476): with the reference's own CPU draw the result is bit-identical
    g = torch.Generator().manual_seed(9)
    img = torch.rand(3, 3, 64, 80, generator=g)
    sig = [15.0, 37.5, 50.0]
    noise = torch.randn(img.shape, generator=g)
    want = torch.stack([img[b].clone().add_(noise[b].clone().mul_(torch.FloatTensor([sig[b]]) / 255.0).float()) for b in range(3)])
    got = ops.add_gaussian_noise(img.to(DEV), sig, noise=noise.to(DEV))
    out.append(result("gaussian_noise_given_draw", got, want, 0.0))
    # device-side draw (Philox + Box-Muller): moments, reproducibility from the seed, independence of samples / seeds
    big = torch.zeros(2, 3, 512, 512, device=DEV)
    z1 = ops.add_gaussian_noise(big, 255.0, seed=1234)
    z2 = ops.add_gaussian_noise(big, 255.0, seed=1234)
    z3 = ops.add_gaussian_noise(big, 255.0, seed=1235)
    n = z1.numel()
    zz = z1.double()
    mean, var = zz.mean().item(), zz.var().item()
    kurt = ((zz - mean) ** 4).mean().item() / var ** 2
    tail = (zz.abs() > 3).double().mean().item()                     # P(|z| > 3) = 2.6998e-3
    corr = lambda a, b: float((a.double().flatten() * b.double().flatten()).mean())
    stats = dict(mean=abs(mean) < 5 / n ** 0.5, var=abs(var - 1) < 5e-3, kurt=abs(kurt - 3) < 3e-2, tail=abs(tail - 2.6998e-3) < 3e-4,
                 same_seed=bool((z1 == z2).all()), other_seed=abs(corr(z1, z3)) < 5e-3, samples=abs(corr(z1[0], z1[1])) < 5e-3,
                 neighbours=abs(corr(z1.flatten()[:-1], z1.flatten()[1:])) < 5e-3)
    out.append(dict(name="gaussian_noise_device_draw", max_err=0.0 if all(stats.values()) else 1.0, tol=0.0, ok=all(stats.values()),
                    note=f"mean {mean:.2e} var {var:.5f} kurtosis {kurt:.4f} P(|z|>3) {tail:.3e}; {stats}"))
    return out


def check_psnr():
    """ops.psnr_u8 (device quantisation + exact integer sums, float64 tail on the host) vs the doubles the unmodified
    reference tensor2img + calculate_psnr returned (tests/golden/psnr.npz): identical float64; integer sums vs oracle."""
    import numpy as np
    from oracle import metrics as M
    from oracle.make_golden_metrics import CASES, make_pair
    ops = _ops()
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "psnr.npz"))["psnr"]
    out = []
    for i, case in enumerate(CASES):
        res, gt = make_pair(case, 500 + i)
        got = ops.psnr_u8(res[None].to(DEV), gt[None].to(DEV), case["crop"])[0]
        same = (got == ref[i]) or (np.isinf(got) and np.isinf(ref[i]))
        out.append(dict(name=f"psnr_{case['kind']}_c{case['c']}_crop{case['crop']}", ok=bool(same), max_err=0.0 if same else
                        abs(got - ref[i]), ref_scale=1.0, tol=0.0, note=f"{got!r} vs reference {ref[i]!r}"))
    # a batch at validation size: per-image sums equal the oracle's integers
    g = torch.Generator().manual_seed(9)
    a = torch.rand(3, 3, 128, 160, generator=g) * 1.2 - 0.1
    b = a + torch.randn(3, 3, 128, 160, generator=g) * 0.02
    got = ops.psnr_u8(a.to(DEV), b.to(DEV), 3)
    want = [M.psnr(a[k].numpy(), b[k].numpy(), 3) for k in range(3)]
    out.append(dict(name="psnr_batch3_128x160_crop3", ok=got == want, max_err=max(abs(x - y) for x, y in zip(got, want)),
                    ref_scale=1.0, tol=0.0, note=f"{got} vs {want}"))
    return out


def check_conv_simt():
    out = []
    for c in CONV_CASES:
        out += _conv_case(impl=1, **c)
    return out


def check_conv_tc_basic():
    return _conv_case(impl=0, **CONV_CASES[0]) + _conv_case(impl=0, **CONV_CASES[1])


def check_conv_tc():
    out = []
    for c in CONV_CASES[2:]:
        out += _conv_case(impl=0, **c)
    return out


def check_conv_origin():
    """per-sample windows (MASA fine search): sample b reads image origin[b,0] shifted by (y0,x0)."""
    ops = _ops()
    out = []
    B, Ci, Co, Hh, Ww = 2, 64, 64, 18, 20
    x = q(rnd(B, Ci, Hh, Ww, seed=3))
    org = torch.tensor([[0, 0, 0], [1, 1, 5], [0, 3, 2], [1, 0, 4], [0, 1, 1]], dtype=torch.int32)
    nw = org.shape[0]
    w = q(rnd(nw, Co, Ci, 3, 3, seed=4) * 0.05)
    wh = ww = 15
    ys = []
    for i in range(nw):
        im, y0, x0 = [int(v) for v in org[i]]
        ys.append(F.conv2d(x[im:im + 1, :, y0:y0 + wh, x0:x0 + ww], w[i]))
    ref = torch.cat(ys, 0)
    wp = torch.cat([ops.pack_conv_weight(w[i].to(DEV)) for i in range(nw)], 0)
    for impl in (1, 0):
        o32, _ = ops.conv_gemm(nhwc(x.to(BF16)), wp, Co, k=3, pad=0, want="f32", w_batched=True, origin=org.to(DEV),
                               window=(wh, ww), impl=impl)
        out.append(result(f"conv_{'simt' if impl else 'tc'}_origin", nchw(o32), ref, 2e-3))
    return out


# =============================================================================================== MDTA
def check_mdta():
    ops = _ops()
    out = []
    for (C_, heads, H, W, B) in ((48, 1, 16, 24, 2), (32, 2, 9, 13, 1), (96, 2, 16, 16, 2), (96, 1, 20, 20, 1),
                                 (192, 8, 8, 8, 2), (384, 8, 8, 8, 1), (64, 1, 32, 40, 1), (16, 1, 64, 64, 1)):
        c = C_ // heads
        qkv = q(rnd(B, 3 * C_, H, W, seed=C_ + heads))
        temp = torch.rand(heads, generator=torch.Generator().manual_seed(1)) + 0.5
        wpo = rnd(C_, C_, seed=2) / C_ ** 0.5
        qq, kk, vv = qkv.view(B, 3, heads, c, H * W).unbind(1)
        qn = qq / qq.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        kn = kk / kk.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        attn = torch.softmax(qn @ kn.transpose(-1, -2) * temp.view(1, heads, 1, 1), -1)
        weff_ref = torch.zeros(B, C_, C_)
        for h in range(heads):
            weff_ref[:, :, h * c:(h + 1) * c] = wpo[:, h * c:(h + 1) * c] @ attn[:, h]
        weff, attn_g = ops.mdta_weff(nhwc(qkv.to(BF16)), C_, heads, temp.to(DEV), wpo.to(DEV), want_attn=True)
        tag = f"C{C_}_h{heads}_{H}x{W}"
        out.append(result(f"mdta_attn_{tag}", attn_g, attn, 2e-3))
        out.append(result(f"mdta_weff_{tag}", weff[..., :C_], weff_ref, 8e-3))
    return out


# =============================================================================================== transformer block
def check_block():
    """One TransformerBlock / ResFusionBlock through the kernel schedule vs the oracle block."""
    from oracle import restormer as O, weights as Wt
    from textualdegremoval_b200.archs import restormer_b200_arch as A
    out = []
    for (dim, heads, ln, bias, fusion, H, W) in ((48, 1, "WithBias", False, False, 16, 16),
                                                (96, 2, "BiasFree", True, False, 8, 24),
                                                (96, 1, "WithBias", False, True, 16, 16),
                                                (32, 2, "WithBias", True, True, 8, 8)):
        cls = A.TransformerResFusionBlock if fusion else A.TransformerBlock
        blk = cls(dim, heads, 2.66, bias, ln)
        sd = Wt.load_seeded(blk, seed=dim + heads)
        x = rnd(2, dim, H, W, seed=dim)
        psd = {"b." + k_: v for k_, v in sd.items()}
        ref = (O.res_fusion_block if fusion else O.transformer_block)(psd, "b", x, heads)
        blk = blk.to(DEV)
        p = A._prep_block(blk)
        x32 = nhwc(x)
        A.run_block(x32, p)
        out.append(result(f"block_d{dim}_h{heads}_{ln}_b{int(bias)}_f{int(fusion)}", nchw(x32), ref, 2e-3,
                          note="inference schedule: fp16 operands, fp32 stream"))
    return out


# =============================================================================================== MASA
def _tile_mask(index_g, index_r, same_win, B, py, px):
    """[B, py, px] bool: windows whose coarse placement and all fine indices agree with the oracle."""
    return ((index_g == index_r).all(dim=1) & same_win).view(B, py, px)


def _warp_on_agreeing_tiles(name, got, ref, ok_tiles, rtol):
    """Compare a warped-feature map only inside the windows whose matches agree (the others gather from a different
    place by construction).  got / ref NCHW [B, C, py*t, px*t]."""
    B, py, px = ok_tiles.shape
    t_y, t_x = got.shape[2] // py, got.shape[3] // px
    m = ok_tiles.repeat_interleave(t_y, 1).repeat_interleave(t_x, 2).unsqueeze(1).float()
    r = result(name, got * m, ref * m, rtol)
    r["note"] = f"compared on {int(ok_tiles.sum())}/{ok_tiles.numel()} windows with identical matches"
    r["ok"] = r["ok"] and bool(ok_tiles.any())
    return r


def check_masa():
    """MASA search/transfer kernels vs the oracle on identical features: the searches see the SAME fp32 deepest-level
    features as the oracle (split-bf16 correlations, 2^-16 relative), the transfer gathers from bf16 copies."""
    from oracle import restormer as O
    from textualdegremoval_b200.archs import restormer_b200_arch as A
    ops = _ops()
    out = []
    for (B, nf, h, w, hr, wr, seed) in ((2, 16, 128, 128, 128, 128, 1), (1, 8, 128, 192, 192, 128, 2),
                                        (1, 16, 256, 256, 256, 256, 3)):
        net = A.RestormerRefFusion(dim=nf, nf=nf, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1,
                                   ext_n_blocks=[1, 1, 1, 1])
        g = torch.Generator().manual_seed(seed)
        smooth = lambda t: F.avg_pool2d(t, 3, 1, 1)
        f_lq_deep = smooth(torch.randn(B, nf * 8, h // 8, w // 8, generator=g))          # NOT bf16-representable
        f_ref = [q(smooth(torch.randn(B, nf * 2 ** i, hr >> i, wr >> i, generator=g))) for i in range(3)]
        f_ref.append(smooth(torch.randn(B, nf * 8, hr >> 3, wr >> 3, generator=g)))
        warps_ref, aux = O.masa_warp(f_lq_deep, f_ref, 8, 8, 1.5, [1, 2, 3], h, w, hr, wr, return_aux=True)
        d = [nf, nf * 2, nf * 4, nf * 8]
        targets = [torch.zeros(B, h >> i, w >> i, d[i], device=DEV) for i in range(4)]
        a = net._masa_warp(nhwc(f_lq_deep), nhwc(f_ref[-1]), [nhwc(t.to(BF16)) for t in f_ref], h, w, hr, wr, targets)
        tag = f"{h}x{w}_{hr}x{wr}"
        score = a["score"][..., : aux["score"].shape[1]].reshape(B, -1, aux["score"].shape[1]).permute(0, 2, 1)
        out.append(result(f"masa_coarse_score_{tag}", score, aux["score"], 2e-5, note="split-bf16 (hi+lo) correlation"))
        agree = (a["idx"].cpu().long() == aux["idx"]).float().mean().item()
        out.append(dict(name=f"masa_coarse_idx_{tag}", max_err=1 - agree, ref_scale=1, tol=0.0, ok=agree == 1.0,
                        note="fraction of blocks whose arg-max differs"))
        y1 = a["origin"][:, 1].view(B, -1).cpu().long()
        x1 = a["origin"][:, 2].view(B, -1).cpu().long()
        same_win = ((y1 == aux["y1"]) & (x1 == aux["x1"])).view(-1)
        nq = aux["index"].shape[1] * aux["index"].shape[2]
        idx_g = a["index"].cpu().long().view(-1, nq)
        idx_r = aux["index"].view(-1, nq)
        agree_f = (idx_g == idx_r)[same_win].float().mean().item() if same_win.any() else 0.0
        out.append(dict(name=f"masa_fine_idx_{tag}", max_err=1 - agree_f, ref_scale=1, tol=1e-3, ok=agree_f >= 0.999,
                        note="fraction of fine matches that differ on identical fp32 features (bar: <= 0.1 %)"))
        out.append(result(f"masa_att_{tag}", a["att"].view(-1, nq), aux["att"].view(-1, nq), 2e-5))
        g_ = a["geom"]
        ok_tiles = _tile_mask(idx_g, idx_r, same_win, B, g_["py"], g_["px"])
        for i in range(4):      # deepest level: the transfer reads the bf16 copy of fp32 features (2^-9 relative)
            out.append(_warp_on_agreeing_tiles(f"masa_warp_l{i}_{tag}", nchw(targets[i]), warps_ref[i], ok_tiles,
                                               4e-3 if i == 3 else 2e-3))
    return out


# =============================================================================================== models vs golden
def _golden(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return json.loads(str(z["meta"])), torch.from_numpy(z["out"])


def psnr_u8(a, b):
    """calculate_psnr on tensor2img outputs (metrics/psnr_ssim.py:9-63, utils_image.py:129-191): clamp, x255, round."""
    a8 = (a.clamp(0, 1) * 255).round().double()
    b8 = (b.clamp(0, 1) * 255).round().double()
    mse = ((a8 - b8) ** 2).mean().item()
    return float("inf") if mse == 0 else 20 * np.log10(255.0 / np.sqrt(mse))


# End-to-end forward parity against the fp32 reference output (outputs are O(1) images).  north_star: |delta| < 1e-3 and
# PSNR within 0.01 dB.  Asserted here:
#   mean |delta| <= 1e-3                    the north-star number, on the mean;
#   max  |delta| <= 9.6e-3                  documented bf16 max tolerance: the reference's OWN bf16 forward differs from
#                                           its fp32 forward by max 9.6e-3 / mean 1.7e-3 / 53.8 dB (SURVEY 8c calibration,
#                                           random-init Restormer 128x128) -- GEMM operands here are bf16 by the metric's
#                                           definition, everything else (residual stream, LN, softmax, Gram) is fp32;
#   PSNR_u8(ours, reference) >= 55 dB       uint8-rounded outputs (val.use_image semantics);
#   |PSNR(ours, gt) - PSNR(reference, gt)| <= 0.01 dB against a synthetic ground truth at ~38 dB.
E2E_MEAN_TOL = 1e-3
E2E_TOL = 9.6e-3
E2E_PSNR_MIN = 55.0
PSNR_DELTA_TOL = 0.01
GUIDED_MEAN_TOL, GUIDED_PSNR_MIN = E2E_MEAN_TOL, E2E_PSNR_MIN


def guided_result(name, y, ref):
    r = result(name, y, ref, 1.0)
    mx = r["max_err"]
    mean = (y - ref).abs().mean().item()
    ps = psnr_u8(y, ref)
    r["ok"] = bool(torch.isfinite(y).all() and mean <= E2E_MEAN_TOL and mx <= E2E_TOL and ps >= E2E_PSNR_MIN)
    r["tol"] = E2E_TOL
    r["note"] = (f"max|d|={mx:.2e} (<= {E2E_TOL:g}) mean|d|={mean:.2e} (<= {E2E_MEAN_TOL:g}) "
                 f"psnr_u8(ours,ref)={ps:.2f}dB (>= {E2E_PSNR_MIN:g})")
    return r


def psnr_delta_result(name, y, ref, seed):
    """PSNR against a ground truth vs the reference's PSNR against it (val.use_image semantics: uint8)."""
    from oracle import weights as Wt
    gt = (ref + 0.05 * (Wt.seeded_image("gt_noise", ref.shape, seed) - 0.5)).clamp(0, 1)
    dp = abs(psnr_u8(y, gt) - psnr_u8(ref, gt))
    return dict(name=f"psnr_delta_{name}", max_err=dp, tol=PSNR_DELTA_TOL, ok=bool(dp <= PSNR_DELTA_TOL),
                note=f"PSNR(ours, gt) {psnr_u8(y, gt):.4f} dB vs PSNR(reference, gt) {psnr_u8(ref, gt):.4f} dB (bar 0.01 dB)")


def check_restormer_golden():
    from oracle import weights as Wt
    from textualdegremoval_b200.archs import define_network
    out = []
    for name in ("restormer_withbias", "restormer_biasfree", "restormer_gray_bias"):
        meta, ref = _golden(name)
        net = define_network(dict(type="Restormer", **meta["cfg"]))
        Wt.load_seeded(net, meta["seed"])
        net = net.to(DEV).eval()
        x = Wt.seeded_image("x", meta["shape"], meta["seed"])
        with torch.no_grad():
            y = net(x.to(DEV)).cpu()
        out.append(guided_result(f"golden_{name}", y, ref))
        out.append(psnr_delta_result(name, y, ref, meta["seed"]))
    return out


def check_guided_golden():
    from oracle import weights as Wt
    from oracle.make_golden import guided_inputs
    from textualdegremoval_b200.archs import define_network
    out = []
    for name in ("guided_restormer_128", "guided_restormer_ragged", "guided_restormer_dual"):
        meta, ref = _golden(name)           # *_dual: dual_pixel_task (option 004_0: 6 input channels, skip_conv, R:957-961)
        net = define_network(dict(type="RestormerRefFusion", **meta["cfg"]))
        Wt.load_seeded(net, meta["seed"])
        net = net.to(DEV).eval()
        lq, rf = guided_inputs(meta)
        with torch.no_grad():
            y = net(lq.to(DEV), rf.to(DEV)).cpu()
        out.append(guided_result(f"golden_{name}", y, ref))
        out.append(psnr_delta_result(name, y, ref, meta["seed"]))
    return out


def check_promptir():
    """PromptIRRefFusion (decoder=True) -- the prompt kernels and the wide-head (c = 176) Gram against
    torch, then the whole net against the golden fixture made by the unmodified reference module."""
    from oracle import weights as Wt
    from oracle.make_golden import guided_inputs
    from textualdegremoval_b200.archs import define_network
    ops = _ops()
    out = []
    # PromptGenBlock pieces (network_promptir_guided_arch.py:424-440)
    B, C_, L, D, S, H, W = 3, 96, 5, 64, 16, 24, 40
    emb = rnd(B, C_, seed=1)
    lw, lb = rnd(L, C_, seed=2) * 0.3, rnd(L, seed=3)
    wts = ops.prompt_weights(emb.to(DEV), lw.to(DEV), lb.to(DEV))
    wts_ref = torch.softmax(F.linear(emb, lw, lb), 1)
    out.append(result("prompt_weights", wts, wts_ref, 1e-5))
    pp = torch.rand(L, D, S, S, generator=torch.Generator().manual_seed(4))
    want = F.interpolate((wts_ref.view(B, L, 1, 1, 1) * pp.unsqueeze(0)).sum(1), (H, W), mode="bilinear")
    for dt in (torch.float16, BF16):
        o16 = torch.empty(B, H, W, D, dtype=dt, device=DEV)
        ops.prompt_mix_resize(pp.to(DEV), wts_ref.to(DEV), H, W, o16)
        out.append(result(f"prompt_mix_resize_{str(dt)[6:]}", o16.permute(0, 3, 1, 2), want, 8e-3 if dt == BF16 else 1e-3))
    # wide heads: c = 176 (noise_level3: 704 channels, 4 heads) through the generic Gram + softmax + fold
    for (C2, heads, Hh, Ww, Bb) in ((704, 4, 16, 16, 2), (352, 2, 9, 13, 1)):
        c = C2 // heads
        qkv = q(rnd(Bb, 3 * C2, Hh, Ww, seed=C2 + heads))
        temp = torch.rand(heads, generator=torch.Generator().manual_seed(1)) + 0.5
        wpo = rnd(C2, C2, seed=2) / C2 ** 0.5
        qq, kk, vv = qkv.view(Bb, 3, heads, c, Hh * Ww).unbind(1)
        qn = qq / qq.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        kn = kk / kk.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        attn = torch.softmax(qn @ kn.transpose(-1, -2) * temp.view(1, heads, 1, 1), -1)
        weff_ref = torch.zeros(Bb, C2, C2)
        for h in range(heads):
            weff_ref[:, :, h * c:(h + 1) * c] = wpo[:, h * c:(h + 1) * c] @ attn[:, h]
        weff, attn_g = ops.mdta_weff(nhwc(qkv.to(BF16)), C2, heads, temp.to(DEV), wpo.to(DEV), want_attn=True)
        out.append(result(f"mdta_wide_attn_C{C2}_h{heads}", attn_g, attn, 2e-3))
        out.append(result(f"mdta_wide_weff_C{C2}_h{heads}", weff[..., :C2], weff_ref, 8e-3))
    # end to end
    meta, ref = _golden("guided_promptir_128")
    net = define_network(dict(type="PromptIRRefFusion", **meta["cfg"]))
    Wt.load_seeded(net, meta["seed"])
    net = net.to(DEV).eval()
    lq, rf = guided_inputs(meta)
    with torch.no_grad():
        y = net(lq.to(DEV), rf.to(DEV)).cpu()
    out.append(guided_result("golden_guided_promptir_128", y, ref))
    out.append(psnr_delta_result("guided_promptir_128", y, ref, meta["seed"]))
    return out


def check_drsformer():
    """the DRSformer family -- top-k sparse attention fold, generic grouped / depthwise stencils, MEFC gate
    and per-sample reduction weights against torch, then both networks against the golden fixtures made by the unmodified
    reference modules (options 007 and 008-010)."""
    from oracle import weights as Wt
    from oracle.make_golden import guided_inputs
    from textualdegremoval_b200.archs import define_network
    ops = _ops()
    out = []
    # TKSA: attn = sum_i w_i softmax(top-k_i masked scores), folded into project_out
    for (C2, heads, Hh, Ww, Bb) in ((48, 1, 16, 24, 2), (96, 2, 12, 12, 1), (192, 2, 8, 8, 1)):
        c = C2 // heads
        qkv = q(rnd(Bb, 3 * C2, Hh, Ww, seed=C2 + heads))
        temp = torch.rand(heads, generator=torch.Generator().manual_seed(1)) + 0.5
        wpo = rnd(C2, C2, seed=2) / C2 ** 0.5
        tw = torch.tensor([0.2, 0.35, 0.15, 0.3])
        qq, kk, vv = qkv.view(Bb, 3, heads, c, Hh * Ww).unbind(1)
        qn = qq / qq.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        kn = kk / kk.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        sc = qn @ kn.transpose(-1, -2) * temp.view(1, heads, 1, 1)
        attn = torch.zeros_like(sc)
        for wi, k_ in zip(tw, (int(c / 2), int(c * 2 / 3), int(c * 3 / 4), int(c * 4 / 5))):
            idx = torch.topk(sc, k=k_, dim=-1)[1]
            mask = torch.zeros_like(sc).scatter_(-1, idx, 1.)
            attn = attn + wi * torch.where(mask > 0, sc, torch.full_like(sc, float("-inf"))).softmax(-1)
        weff_ref = torch.zeros(Bb, C2, C2)
        for h in range(heads):
            weff_ref[:, :, h * c:(h + 1) * c] = wpo[:, h * c:(h + 1) * c] @ attn[:, h]
        weff, attn_g = ops.mdta_weff(nhwc(qkv.to(BF16)), C2, heads, temp.to(DEV), wpo.to(DEV), want_attn=True,
                                     topk_w=tw.to(DEV))
        out.append(result(f"tksa_attn_C{C2}_h{heads}", attn_g, attn, 2e-3))
        out.append(result(f"tksa_weff_C{C2}_h{heads}", weff[..., :C2], weff_ref, 8e-3))
    # generic stencils against F.conv2d on the same 16-bit-rounded inputs
    B, H, W, C_ = 2, 19, 45, 24
    x = q(rnd(B, C_, H, W, seed=5))
    x16 = nhwc(x.to(BF16))
    idx = torch.arange(C_, dtype=torch.int32, device=DEV)
    for K, dil in ((1, 1), (3, 1), (5, 1), (7, 1), (3, 2), (5, 2), (7, 2)):
        w = rnd(C_, 1, K, K, seed=K + dil) / K
        o = torch.empty(B, H, W, C_, dtype=BF16, device=DEV)
        ops.grouped_stencil(x16, idx, w.to(DEV), None, K, o, dil=dil, relu=(K == 5))
        want = F.conv2d(x, w, padding=dil * (K - 1) // 2, dilation=dil, groups=C_)
        out.append(result(f"stencil_dw_k{K}_d{dil}", o.permute(0, 3, 1, 2), F.relu(want) if K == 5 else want, 8e-3))
    o = torch.empty(B, H, W, C_, dtype=BF16, device=DEV)
    ops.grouped_stencil(x16, idx, None, None, 3, o, pool=True)
    out.append(result("stencil_avgpool3", o.permute(0, 3, 1, 2), F.avg_pool2d(x, 3, 1, 1, count_include_pad=False), 8e-3))
    for K in (3, 5):                                     # Conv2d(2h, h, K, groups=h) over a permuted channel table
        hch = 11
        perm = torch.randperm(2 * hch, generator=torch.Generator().manual_seed(K))
        w = rnd(hch, 2, K, K, seed=20 + K) / K
        bias = rnd(hch, seed=30 + K)
        xin = q(rnd(B, C_, H, W, seed=6))
        o = torch.zeros(B, H, W, 16, dtype=BF16, device=DEV)
        ops.grouped_stencil(nhwc(xin.to(BF16)), perm.to(torch.int32).to(DEV), w.to(DEV), bias.to(DEV), K, o[..., 3:3 + hch],
                            relu=True)
        want = F.relu(F.conv2d(xin[:, perm], w, bias, padding=K // 2, groups=hch))
        out.append(result(f"stencil_grouped_k{K}", o[..., 3:3 + hch].permute(0, 3, 1, 2), want, 8e-3))
    # MEFC gate and per-sample reduction weights
    emb = rnd(3, 32, seed=7)
    w1, b1, w2, b2 = rnd(64, 32, seed=8) * 0.2, rnd(64, seed=9), rnd(32, 64, seed=10) * 0.2, rnd(32, seed=11)
    g = ops.mefc_gate(emb.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV), 8)
    g_ref = torch.softmax(F.linear(F.relu(F.linear(emb, w1, b1)), w2, b2).view(3, 4, 8), -1)
    out.append(result("mefc_gate", g, g_ref, 1e-5))
    wo = rnd(16, 8 * 16, seed=12)
    wm = ops.mefc_mix_weights(wo.to(DEV), g[:, 2], 16, BF16)
    out.append(result("mefc_mix_weights", wm[..., :128], wo.unsqueeze(0) * g_ref[:, 2].repeat_interleave(16, -1).unsqueeze(1), 8e-3))
    # end to end
    for name, typ in (("guided_drsformer_spa_128", "DRSformer200L_SPA_RefFusion"),
                      ("guided_drsformer_spa_bias_ragged", "DRSformer200L_SPA_RefFusion"),
                      ("guided_drsformer_128", "DRSformerRefFusion")):
        meta, ref = _golden(name)
        net = define_network(dict(type=typ, **meta["cfg"]))
        Wt.load_seeded(net, meta["seed"])
        net = net.to(DEV).eval()
        lq, rf = guided_inputs(meta)
        with torch.no_grad():
            y = net(lq.to(DEV), rf.to(DEV)).cpu()
        out.append(guided_result(f"golden_{name}", y, ref))
        out.append(psnr_delta_result(name, y, ref, meta["seed"]))
    return out


class _ForcedMatches:
    """Test hook: run a guided net with the ORACLE's coarse / fine matches instead of its own arg-max results (the
    confidence is re-read from our own correlation at the forced index).  Around 1 % of the fine searches of a
    random-weight fixture are decided by score gaps below 3e-6 -- the oracle's own fp32 summation noise -- and every
    such tie that resolves the other way moves a 3s x 3s patch of warped reference features at every scale; the guided
    NAFNet amplifies one flipped match to |delta| 7e-2.  End-to-end parity is therefore asserted in two parts: the
    matches agree up to ties (_match_agreement / _book_agreement), and GIVEN identical matches the output meets the bar."""

    def __init__(self, y1, x1, index):
        self.y1, self.x1, self.index = y1.reshape(-1), x1.reshape(-1), index

    def __enter__(self):
        ops = _ops()
        self.ops, self.c, self.f = ops, ops.masa_coarse_argmax, ops.masa_fine_argmax

        def coarse(score, nblk, d_y, d_x):
            idx, origin = self.c(score, nblk, d_y, d_x)
            origin[:, 1] = self.y1.to(origin)
            origin[:, 2] = self.x1.to(origin)
            return idx, origin

        def fine(corr):
            index, att = self.f(corr)
            forced = self.index.reshape(index.shape).to(index)
            nwin, dy, dx, nq = corr.shape
            att_f = corr.view(nwin, dy * dx, nq).gather(1, forced.long().unsqueeze(1)).squeeze(1).contiguous()
            return forced.contiguous(), att_f

        ops.masa_coarse_argmax, ops.masa_fine_argmax = coarse, fine
        return self

    def __exit__(self, *exc):
        self.ops.masa_coarse_argmax, self.ops.masa_fine_argmax = self.c, self.f


# Unconditional bound for the guided NAFNet (informational beside the conditional check above): a handful of tied matches
# resolve differently and each one costs a visible patch.
NAF_GUIDED_PSNR_MIN = 45.0


def naf_unconditional_result(name, y, ref):
    r = guided_result(name, y, ref)
    ps = psnr_u8(y, ref)
    mean = (y - ref).abs().mean().item()
    r["ok"] = bool(torch.isfinite(y).all() and ps >= NAF_GUIDED_PSNR_MIN and mean <= 2e-3)
    r["tol"] = None
    r["note"] = ("own matches (tied searches may resolve differently, see _ForcedMatches): " + r["note"] +
                 f"; asserted here: psnr >= {NAF_GUIDED_PSNR_MIN:g} dB, mean <= 2e-3")
    return r


def check_nafnet():
    """NAFNet helper kernels, one NAFBlock, and both NAFNet classes against oracle / golden fixtures."""
    from oracle import nafnet as ON, weights as Wt
    from oracle.make_golden import denoise_inputs, guided_inputs
    from textualdegremoval_b200.archs import define_network, nafnet_b200_arch as A
    ops = _ops()
    out = []
    x = q(rnd(2, 64, 9, 11, seed=5))
    y = ops.gate_mul(nhwc(x.to(BF16)))
    out.append(result("gate_mul", nchw(y), x[:, :32] * x[:, 32:], 8e-3))
    for (c, H, W, B) in ((32, 16, 16, 2), (128, 40, 24, 1)):
        blk = A.NAFBlock(c)
        sd = Wt.load_seeded(blk, seed=c)
        xin = rnd(B, c, H, W, seed=c + 1)
        ref = ON.naf_block({"b." + k_: v for k_, v in sd.items()}, "b", xin)
        blk = blk.to(DEV)
        x32 = nhwc(xin)
        A.run_naf_block(x32, A._prep_naf(blk))
        out.append(result(f"naf_block_c{c}", nchw(x32), ref, 1.5e-2))
    for name in ("nafnet_tiny_gray64", "nafnet_rgb_ragged"):
        meta, ref = _golden(name)
        net = define_network(dict(type="NAFNet", **meta["cfg"]))
        Wt.load_seeded(net, meta["seed"])
        net = net.to(DEV).eval()
        lq, _ = denoise_inputs(meta)
        with torch.no_grad():
            yy = net(lq.to(DEV)).cpu()
        out.append(guided_result(f"golden_{name}", yy, ref))
        out.append(psnr_delta_result(name, yy, ref, meta["seed"]))
    meta, ref = _golden("guided_nafnet_256")
    net = define_network(dict(type="NAFNetRefFusion", **meta["cfg"]))
    Wt.load_seeded(net, meta["seed"])
    net = net.to(DEV).eval()
    lq, rf = guided_inputs(meta)
    with torch.no_grad():
        yy, aux = net(lq.to(DEV), rf.to(DEV), return_aux=True)
        yy = yy.cpu()
        _, aux_r = ON.nafnet_ref_fusion_forward(Wt.seeded_state_dict({k_: v.shape for k_, v in net.state_dict().items()},
                                                                     meta["seed"]), lq, rf, return_aux=True)
    out += _match_agreement("naf_stage", aux, aux_r)
    out.append(naf_unconditional_result("golden_guided_nafnet_256_own_matches", yy, ref))
    with _ForcedMatches(aux_r["y1"], aux_r["x1"], aux_r["index"]), torch.no_grad():
        yf = net(lq.to(DEV), rf.to(DEV)).cpu()
    out.append(guided_result("golden_guided_nafnet_256", yf, ref))
    out.append(psnr_delta_result("guided_nafnet_256", yf, ref, meta["seed"]))
    return out


def check_vit():
    """ViT glue kernels, DINOv2 / CLIP towers, mappers and reference-crop selection against oracle / golden."""
    from oracle import vit as OV, weights as Wt
    from textualdegremoval_b200.archs import vit_b200 as VB
    ops = _ops()
    out = []
    # softmax rows / crop-resize against torch
    s = rnd(3, 5, 37, seed=1) * 4
    sp = torch.zeros(3, 5, 40)
    sp[..., :37] = s
    o16 = torch.empty(3, 5, 40, dtype=BF16, device=DEV)
    ops.softmax_rows(sp.to(DEV), 37, 0.25, o16)
    out.append(result("softmax_rows", o16[..., :37], torch.softmax(s * 0.25, -1), 8e-3))
    out.append(result("softmax_rows_pad", o16[..., 37:], torch.zeros(3, 5, 3), 0.0))
    img = torch.rand(2, 3, 40, 52, generator=torch.Generator().manual_seed(2))
    org = torch.tensor([[0, 0, 0], [1, 8, 12], [0, 3, 20]], dtype=torch.int32)
    ref = torch.stack([F.interpolate(img[b:b + 1, :, y:y + 32, x:x + 32], size=(28, 42), mode="bilinear")[0]
                       for b, y, x in org.tolist()])
    out.append(result("crop_resize", ops.crop_resize(img.to(DEV), org.to(DEV), (32, 32), (28, 42)), ref, 1e-5))
    # fused attention (tdr_vit_attention) against fp32 torch on the same bf16-rounded q, k, v: the head dims / token counts
    # of the towers (hd 64, N 1370; hd 80, N 257), single-tile and ragged key / query tiles, strided (padded) rows
    for name, (B, N, heads, hd, pad) in dict(dino=(2, 1370, 12, 64, 0), clip=(3, 257, 16, 80, 0), one_tile=(2, 128, 2, 64, 0),
                                             ragged=(2, 129, 3, 32, 8), tiny=(3, 5, 4, 16, 0), two_tiles=(1, 256, 2, 80, 16)).items():
        D = heads * hd
        buf = torch.zeros(B, 1, N, 3 * D + pad, dtype=BF16)
        buf[..., :3 * D] = (rnd(B, 1, N, 3 * D, seed=N + hd) * 1.5).to(BF16)
        qkv = buf.to(DEV)[..., :3 * D]
        o = ops.vit_attention(qkv, heads, hd, hd ** -0.5)
        q, k, v = [t.reshape(B, N, heads, hd).permute(0, 2, 1, 3) for t in buf[:, 0, :, :3 * D].float().split(D, -1)]
        want = (torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, -1) @ v).permute(0, 2, 1, 3).reshape(B, 1, N, D)
        out.append(result(f"vit_attention_{name}", o, want, 1e-2))
    # towers
    meta, gold = _golden("dino_vit_tiny")
    net = VB.DinoVisionTransformer(**meta["cfg"])
    sd = Wt.load_seeded(net, meta["seed"])
    net = net.to(DEV).eval()
    x = Wt.seeded_image("x", meta["shape"], meta["seed"])
    with torch.no_grad():
        y = net(x.to(DEV)).cpu()
    out.append(result("golden_dino_vit_tiny", y, gold, 2e-2))
    meta, gold = _golden("clip_vit_tiny")
    net = VB.CLIPVisionTower(**meta["cfg"])
    Wt.load_seeded(net, meta["seed"])
    net = net.to(DEV).eval()
    with torch.no_grad():
        y = net(Wt.seeded_image("x", meta["shape"], meta["seed"]).to(DEV), output_hidden_states=True)[0].cpu()
    out.append(result("golden_clip_vit_tiny", y, gold, 2e-2))
    z = np.load(os.path.join(ROOT, "tests", "golden", "mappers_tiny.npz"))
    meta = json.loads(str(z["meta"]))
    c = meta["cfg"]
    m, cm = VB.Mapper(c["input_dim"], c["mid_dim"], c["num_words"]), VB.CleanMapper(c["mid_dim"], c["mid_dim"], c["num_words"])
    Wt.load_seeded(m, meta["seed"]); Wt.load_seeded(cm, meta["seed"] + 1)
    m, cm = m.to(DEV), cm.to(DEV)
    emb = Wt.seeded_image("emb", meta["shape"], meta["seed"]) * 2 - 1
    with torch.no_grad():
        w1 = m([emb.to(DEV)])
        w2 = cm(w1)
    out.append(result("golden_mapper", w1.cpu(), torch.from_numpy(z["out"]), 2e-2))
    out.append(result("golden_clean_mapper", w2.cpu(), torch.from_numpy(z["out2"]), 3e-2))
    # reference-crop selection: lq is a blurred copy of one specific crop of ref -> that crop must win
    net = VB.DinoVisionTransformer(img_size=70, patch_size=14, embed_dim=64, depth=2, num_heads=4)
    sd = Wt.load_seeded(net, 61)
    net = net.to(DEV).eval()
    g = torch.Generator().manual_seed(5)
    ref_img = F.avg_pool2d(torch.rand(2, 3, 84, 84, generator=g), 3, 1, 1)
    lq = torch.stack([ref_img[0, :, 14:70, 0:56], ref_img[1, :, 28:84, 28:84]]) + 0.01 * torch.randn(2, 3, 56, 56, generator=g)
    with torch.no_grad():
        sel, idx, cos = VB.select_reference_crop(net, lq.to(DEV), ref_img.to(DEV))
        sel_r, idx_r, cos_r = OV.dino_select_crop(sd, lq, ref_img, heads=4)
    out.append(result("dino_select_cosine", cos, cos_r[:, 0], 2e-2))
    same = bool((idx.cpu() == idx_r).all())
    out.append(dict(name="dino_select_index", max_err=0.0 if same else 1.0, ref_scale=1, tol=0.0, ok=same,
                    note=f"ours {idx.cpu().tolist()} oracle {idx_r.tolist()}"))
    if same:
        out.append(result("dino_select_crop", sel, sel_r, 1e-6))
    return out


def check_optim():
    """Fused clip + AdamW (+EMA) tail on flat buffers vs torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW."""
    from textualdegremoval_b200.ddp import DDPStep
    out = []
    torch.manual_seed(3)
    shapes = {"a.weight": (33, 17), "masa_x.weight": (65,), "b.bias": (129, 3), "masa_y.weight": (7, 7, 3)}
    ps = {k: torch.nn.Parameter(torch.randn(*s, device=DEV)) for k, s in shapes.items()}
    ref = {k: torch.nn.Parameter(p.detach().clone()) for k, p in ps.items()}
    opt = torch.optim.AdamW([dict(params=[ref[k] for k in shapes if "masa" not in k], lr=2e-4),
                             dict(params=[ref[k] for k in shapes if "masa" in k], lr=1e-4)],
                            lr=2e-4, weight_decay=1e-4, betas=(0.9, 0.999))
    eng = DDPStep(ps.items(), lr=2e-4, ref_lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), max_grad_norm=0.01,
                  ema_decay=0.999)
    ema_ref = {k: p.detach().clone() for k, p in ref.items()}
    for step in range(4):
        for k in shapes:
            g = torch.randn(*shapes[k], device=DEV) * (0.5 if step % 2 else 1e-4)     # clipped and unclipped steps
            ps[k].grad.copy_(g)
            ref[k].grad = g.clone()
        torch.nn.utils.clip_grad_norm_(list(ref.values()), 0.01)
        opt.step()
        eng.step()
        for k in shapes:
            ema_ref[k].mul_(0.999).add_(ref[k].detach(), alpha=0.001)
    for k in shapes:
        out.append(result(f"adamw_{k}", ps[k].detach(), ref[k].detach(), 2e-6))
    ema = torch.cat([v.reshape(-1) for e, g in zip(eng.ema, eng.groups) for v in g.views(e)])
    order = [k for k in shapes if "masa" not in k] + [k for k in shapes if "masa" in k]
    out.append(result("ema", ema, torch.cat([ema_ref[k].reshape(-1) for k in order]), 2e-6))
    return out


def check_guided_stages():
    """Guided net stage by stage against the oracle (features, match indices, warps, output)."""
    from oracle import restormer as O, weights as Wt
    from oracle.make_golden import GUIDED_CASES, guided_inputs
    from textualdegremoval_b200.archs import define_network
    out = []
    meta = GUIDED_CASES["guided_restormer_128"]
    net = define_network(dict(type="RestormerRefFusion", **meta["cfg"]))
    sd = Wt.load_seeded(net, meta["seed"])
    net = net.to(DEV).eval()
    lq, rf = guided_inputs(meta)
    with torch.no_grad():
        y_ref, aux_r = O.restormer_ref_fusion_forward(sd, lq, rf, return_aux=True)
        y, aux = net(lq.to(DEV), rf.to(DEV), return_aux=True)
    for i in range(4):
        out.append(result(f"stage_feat_lq_l{i}", nchw(aux["feat_lq"][i]), aux_r["feat_lq"][i], 2e-2))
        out.append(result(f"stage_feat_ref_l{i}", nchw(aux["feat_ref"][i]), aux_r["feat_ref"][i], 2e-2))
    ds = float(aux["deep_scale"][0])          # the deepest level runs in units of a power-of-two scale (masa._masa_encode)
    out.append(result("stage_deep32_lq", nchw(aux["deep32_lq"]) * ds, aux_r["feat_lq"][3], 2e-3,
                      note=f"fp32 stream of the deepest level (fp16 GEMM operands), level scale {ds:g}"))
    out.append(result("stage_deep32_ref", nchw(aux["deep32_ref"]) * ds, aux_r["feat_ref"][3], 2e-3))
    out += _match_agreement("stage", aux, aux_r)
    B = lq.shape[0]
    nq = aux_r["index"].shape[1] * aux_r["index"].shape[2]
    y1 = aux["origin"][:, 1].view(B, -1).cpu().long()
    x1 = aux["origin"][:, 2].view(B, -1).cpu().long()
    same_win = ((y1 == aux_r["y1"]) & (x1 == aux_r["x1"])).view(-1)
    ok_tiles = _tile_mask(aux["index"].cpu().long().view(-1, nq), aux_r["index"].view(-1, nq), same_win, B,
                          aux["geom"]["py"], aux["geom"]["px"])
    for i in range(4):
        out.append(_warp_on_agreeing_tiles(f"stage_warp_l{i}", nchw(aux["warps"][i]), aux_r["warps"][i], ok_tiles, 1e-2))
    out.append(guided_result("stage_output", y.cpu(), y_ref))
    return out


FINE_AGREE_MIN = 0.999
# Arg-max mismatches across a gap below TIE_TOL in the ORACLE's own fp32 score are ties, not errors: the oracle's fp32
# dot products over 9C = 1296..9216 terms carry ~1e-6 of summation-order noise themselves (the 128x128 fixture has a
# coarse top-2 gap of 1.19e-6 = one fp32 ulp, which the reference's own result depends on the BLAS build for).
TIE_TOL = 1e-5


def _match_agreement(tag, aux, aux_r):
    """Coarse / fine arg-max agreement of an end-to-end run with the oracle's (same inputs, our features vs fp32).
    A mismatch counts as agreement when the oracle's score of our candidate is within TIE_TOL of its maximum."""
    sc = aux_r["score"]                                          # [B, nblk, Hr*Wr]
    B, nblk, _ = sc.shape
    ours = aux["idx"].cpu().long().view(B, nblk)
    gap_c = sc.max(-1).values - sc.gather(2, ours.unsqueeze(-1)).squeeze(-1)
    exact_c = (ours == aux_r["idx"]).float().mean().item()
    agree_c = (gap_c <= TIE_TOL).float().mean().item()
    nq = aux_r["index"].shape[1] * aux_r["index"].shape[2]
    y1 = aux["origin"][:, 1].view(B, -1).cpu().long()
    x1 = aux["origin"][:, 2].view(B, -1).cpu().long()
    same_win = ((y1 == aux_r["y1"]) & (x1 == aux_r["x1"])).view(-1)
    corr = aux_r["corr"]                                         # [M, nq, d*d] in the oracle's windows
    mine = aux["index"].cpu().long().view(-1, nq)
    gap_f = corr.max(-1).values - corr.gather(2, mine.unsqueeze(-1)).squeeze(-1)
    exact_f = (mine == aux_r["index"].view(-1, nq))[same_win].float().mean().item() if same_win.any() else 0.0
    agree_f = (gap_f <= TIE_TOL)[same_win].float().mean().item() if same_win.any() else 0.0
    worst = gap_f[same_win].max().item() if same_win.any() else float("nan")
    return [dict(name=f"{tag}_coarse_idx", max_err=1 - agree_c, ref_scale=1, tol=0.0, ok=agree_c == 1.0,
                 note=f"coarse arg-max: identical {exact_c:.4f}, identical up to ties < {TIE_TOL:g} {agree_c:.4f}; "
                      f"largest oracle-score gap jumped {gap_c.max().item():.2e}"),
            dict(name=f"{tag}_fine_idx", max_err=1 - agree_f, ref_scale=1, tol=1 - FINE_AGREE_MIN,
                 ok=agree_f >= FINE_AGREE_MIN,
                 note=f"fine arg-max in {int(same_win.sum())}/{same_win.numel()} identically placed windows: identical "
                      f"{exact_f:.4f}, up to ties {agree_f:.4f} (bar 99.9 %); largest gap jumped {worst:.2e}")]



# =============================================================================================== backward kernels
def grad_result(name, got, ref, rel=3e-2):
    """Gradient comparison: relative L2 error of the whole tensor (bf16 operands, fp32 accumulation)."""
    got = got.detach().float().cpu().reshape(-1)
    ref = ref.detach().float().cpu().reshape(-1)
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    nr = ref.norm().item()
    err = (got - ref).norm().item() / max(nr, 1e-20)
    ok = bool(torch.isfinite(got).all()) and (err <= rel or (got - ref).abs().max().item() < 1e-9)
    return dict(name=name, max_err=err, ref_scale=nr, tol=rel, ok=ok, note="rel-L2")


def check_wgrad():
    """tdr_wgrad (tcgen05 pixel-contraction GEMM) vs autograd of F.conv2d on the same bf16-rounded operands."""
    ops = _ops()
    out = []
    cases = [  # B, H, W, Ci, Co, k, stride, pad, dil
        (2, 16, 16, 48, 144, 1, 1, 0, 1), (1, 8, 24, 96, 96, 1, 1, 0, 1), (2, 16, 16, 256, 96, 1, 1, 0, 1),
        (1, 16, 32, 96, 512, 1, 1, 0, 1), (2, 8, 8, 384, 40, 1, 1, 0, 1), (1, 8, 8, 520, 136, 1, 1, 0, 1),
        (2, 16, 16, 48, 48, 3, 1, 1, 1), (1, 12, 20, 96, 24, 3, 1, 1, 1), (2, 16, 16, 16, 32, 3, 2, 1, 1),
        (1, 9, 13, 32, 64, 3, 1, 1, 1), (1, 16, 16, 8, 48, 3, 1, 1, 1), (1, 16, 16, 96, 8, 3, 1, 1, 1),
        (1, 1, 300, 80, 160, 1, 1, 0, 1),
    ]
    for (B, H, W, Ci, Co, k, st, pad, dil) in cases:
        x = q(rnd(B, Ci, H, W, seed=Ci + H))
        wt = torch.zeros(Co, Ci, k, k, requires_grad=True)
        y = F.conv2d(x, wt, None, stride=st, padding=pad, dilation=dil)
        dy = q(rnd(*y.shape, seed=Co + W))
        (ref,) = torch.autograd.grad(y, wt, dy)
        dw = torch.full((Co, Ci, k, k), 0.5, device=DEV)
        ops.wgrad(nhwc(dy.to(BF16)), nhwc(x.to(BF16)), dw, k=k, stride=st, pad=pad, dil=dil, accumulate=True)
        out.append(grad_result(f"wgrad_B{B}_{H}x{W}_Ci{Ci}_Co{Co}_k{k}s{st}", dw - 0.5, ref, 1e-2))
    # per-sample (MDTA Weff gradient), strided channel-slice operands
    B, H, W, C_ = 2, 16, 16, 96
    buf = q(rnd(B, H, W, 3 * C_, seed=5))
    dyv = q(rnd(B, H, W, C_, seed=6))
    ref = torch.einsum("bhwo,bhwi->boi", dyv, buf[..., 2 * C_:])
    dwe = torch.empty(B, C_, C_, device=DEV)
    ops.wgrad(dyv.to(BF16).to(DEV), buf.to(BF16).to(DEV)[..., 2 * C_:], dwe, Co=C_, Ci=C_, per_sample=True,
              strides=(C_ * C_, C_, 1, 0), accumulate=False)
    out.append(grad_result("wgrad_per_sample_slice", dwe, ref, 1e-2))
    # channel maps (GDFN padded halves): padded 2*hp = 272 kernel rows -> 2*h = 254 parameter rows
    h, hp = 127, 136
    idx2 = torch.cat([torch.arange(h), hp + torch.arange(h)])
    m2 = torch.full((2 * hp,), -1, dtype=torch.int32)
    m2[idx2] = torch.arange(2 * h, dtype=torch.int32)
    dyp = torch.zeros(1, 8, 8, 2 * hp)
    dyl = q(rnd(1, 8, 8, 2 * h, seed=7))
    dyp[..., idx2] = dyl
    xx = q(rnd(1, 8, 8, 48, seed=8))
    ref = torch.einsum("bhwo,bhwi->oi", dyl, xx).reshape(2 * h, 48, 1, 1)
    dw = torch.zeros(2 * h, 48, 1, 1, device=DEV)
    ops.wgrad(dyp.to(BF16).to(DEV), xx.to(BF16).to(DEV), dw, co_map=m2.to(DEV))
    out.append(grad_result("wgrad_co_map", dw, ref, 1e-2))
    return out


def check_bwd_pointwise():
    ops = _ops()
    out = []
    # colsum
    for C_ in (48, 288, 1024, 2560):
        x = q(rnd(2, 7, 9, C_, seed=C_))
        o = torch.ones(C_, device=DEV)
        ops.colsum(x.to(BF16).to(DEV), o)
        out.append(grad_result(f"colsum_C{C_}", o - 1, x.sum((0, 1, 2)), 1e-4))
    # depthwise wgrad
    for (C_, H, W, bias) in ((48, 9, 13, True), (144, 16, 16, False), (288, 8, 12, True), (2048, 4, 4, True)):
        x = q(rnd(2, C_, H, W, seed=C_ + 1))
        wt = torch.zeros(C_, 1, 3, 3, requires_grad=True)
        bb = torch.zeros(C_, requires_grad=True)
        y = F.conv2d(x, wt, bb, padding=1, groups=C_)
        dy = q(rnd(*y.shape, seed=C_ + 2))
        rw, rb = torch.autograd.grad(y, (wt, bb), dy)
        dw = torch.zeros(C_, 1, 3, 3, device=DEV)
        db = torch.zeros(C_, device=DEV) if bias else None
        ops.dwconv3x3_wgrad(nhwc(dy.to(BF16)), nhwc(x.to(BF16)), dw, db)
        out.append(grad_result(f"dw_wgrad_C{C_}_{H}x{W}", dw, rw, 1e-4))
        if bias:
            out.append(grad_result(f"dw_bgrad_C{C_}", db, rb, 1e-4))
    # dwconv data gradient = dwconv with flipped taps
    x = q(rnd(1, 64, 10, 12, seed=3)).requires_grad_(True)
    wt = rnd(64, 1, 3, 3, seed=4) * 0.3
    y = F.conv2d(x, wt, None, padding=1, groups=64)
    dy = q(rnd(*y.shape, seed=5))
    (rx,) = torch.autograd.grad(y, x, dy)
    dx = ops.dwconv3x3(nhwc(dy.to(BF16)), ops.pack_dw_weight(wt.flip(2, 3).to(DEV)), None)
    out.append(result("dwconv_dgrad_flipped", nchw(dx), rx, 1e-2))
    # rownorm backward
    for C_ in (48, 96, 192, 768):
        for mode in (0, 1, 2):
            x = (rnd(2, 5, 7, C_, seed=C_) * 2 + 0.3).requires_grad_(True)
            w = (1 + 0.2 * rnd(C_, seed=C_ + 1)).requires_grad_(True)
            b = (0.1 * rnd(C_, seed=C_ + 2)).requires_grad_(True)
            mu = x.mean(-1, keepdim=True)
            var = ((x - mu) ** 2).mean(-1, keepdim=True)
            y = {0: x * 1.0, 1: (x - mu) / torch.sqrt(var + 1e-5) * w + b, 2: x / torch.sqrt(var + 1e-5) * w}[mode]
            dy = q(rnd(*y.shape, seed=C_ + 3))
            add = rnd(*y.shape, seed=C_ + 4)
            gs = torch.autograd.grad(y, (x, w, b)[: (1, 3, 2)[mode]], dy)
            dwt = torch.zeros(C_, device=DEV) if mode else None
            dbt = torch.zeros(C_, device=DEV) if mode == 1 else None
            dx = ops.rownorm_bwd(x.detach().to(DEV) if mode else None, dy.to(BF16).to(DEV), mode,
                                 w.detach().to(DEV) if mode else None, 1e-5, add=add.to(DEV), dweight=dwt, dbias=dbt)
            out.append(grad_result(f"rownorm_bwd_dx_m{mode}_C{C_}", dx - add.to(DEV), gs[0], 1e-4))
            if mode:
                out.append(grad_result(f"rownorm_bwd_dw_m{mode}_C{C_}", dwt, gs[1], 1e-4))
            if mode == 1:
                out.append(grad_result(f"rownorm_bwd_db_m{mode}_C{C_}", dbt, gs[2], 1e-4))
    # gate backward
    for gate in (1, 2):
        yv = q(rnd(2, 6, 5, 272, seed=9)).requires_grad_(True)
        a, b = yv.chunk(2, -1)
        gv = (F.gelu(a) if gate == 1 else a) * b
        dg = q(rnd(*gv.shape, seed=10))
        (ry,) = torch.autograd.grad(gv, yv, dg)
        dyo = torch.empty(2, 6, 5, 272, device=DEV, dtype=BF16)
        ops.gate_bwd(yv.detach().to(BF16).to(DEV), dg.to(BF16).to(DEV), gate, out=dyo)
        out.append(grad_result(f"gate_bwd_g{gate}", dyo, ry, 8e-3))
    # fused recompute + gate backward (tdr_dwconv3x3_gate_bwd) == dwconv(gate 0) followed by gate_bwd
    for gate in (1, 2):
        hx = q(rnd(2, 6, 9, 272, seed=21))
        w9 = (rnd(9, 272, seed=22) * 0.3).to(DEV)
        b9 = (rnd(272, seed=23) * 0.1).to(DEV)
        dgv = q(rnd(2, 6, 9, 136, seed=24))
        add = rnd(2, 136, seed=25).to(DEV) if gate == 2 else None
        hx_d, dg_d = hx.to(BF16).to(DEV), dgv.to(BF16).to(DEV)
        ysep = ops.dwconv3x3(hx_d, w9, b9, gate=0)
        ref_dy = ops.gate_bwd(ysep.clone(), dg_d, gate, dg_add=add)
        fused = ops.dwconv3x3_gate_bwd(hx_d, w9, b9, gate, dg_d, dg_add=add)
        out.append(grad_result(f"dwconv_gate_bwd_fused_g{gate}", fused, ref_dy, 1e-2))
    # scale_add / dot / pixel shuffle / relu mask
    xs, ys = rnd(2, 4, 4, 96, seed=11), rnd(2, 4, 4, 96, seed=12)
    al = torch.tensor([0.37], device=DEV)
    out.append(result("scale_add", ops.scale_add(xs.to(DEV), ys.to(DEV), scale_ptr=al), 0.37 * xs + ys, 1e-6))
    o = torch.ones(1, device=DEV)
    ops.dot_f32(xs.to(DEV), ys.to(DEV), o)
    out.append(result("dot_f32", o - 1, (xs * ys).sum().reshape(1), 1e-5))
    t = q(rnd(2, 16, 6, 10, seed=13))                                       # NCHW
    out.append(result("pixel_unshuffle", nchw(ops.pixel_shuffle(nhwc(t.to(BF16)), 1)), F.pixel_unshuffle(t, 2), 0.0))
    out.append(result("pixel_shuffle", nchw(ops.pixel_shuffle(nhwc(t.to(BF16)), 2)), F.pixel_shuffle(t, 2), 0.0))
    for cc, hh, ww in ((96, 6, 10), (64, 8, 4), (24, 6, 6)):                # 16-byte vector kernels (C % 32 / C % 8) + fallback
        t = q(rnd(3, cc, hh, ww, seed=16 + cc))
        out.append(result(f"pixel_unshuffle_C{cc}", nchw(ops.pixel_shuffle(nhwc(t.to(BF16)), 1)), F.pixel_unshuffle(t, 2), 0.0))
        out.append(result(f"pixel_shuffle_C{cc}", nchw(ops.pixel_shuffle(nhwc(t.to(BF16)), 2)), F.pixel_shuffle(t, 2), 0.0))
    yy, dd = q(rnd(1, 5, 5, 64, seed=14)), q(rnd(1, 5, 5, 64, seed=15))
    out.append(result("relu_mask", ops.relu_mask(yy.to(BF16).to(DEV), dd.to(BF16).to(DEV)), dd * (yy > 0), 0.0))
    return out


def check_block_bwd():
    """One TransformerBlock / ResFusionBlock: training forward + explicit backward vs autograd through the oracle."""
    from oracle import restormer as O, weights as Wt
    from textualdegremoval_b200.archs import restormer_b200_arch as A
    from textualdegremoval_b200.archs import restormer_train as TR
    out = []
    for (dim, heads, ln, bias, fusion, H, W) in ((48, 1, "WithBias", False, False, 16, 16),
                                                (96, 2, "BiasFree", True, False, 8, 24),
                                                (96, 1, "WithBias", False, True, 16, 16),
                                                (32, 2, "WithBias", True, True, 8, 8)):
        cls = A.TransformerResFusionBlock if fusion else A.TransformerBlock
        blk = cls(dim, heads, 2.66, bias, ln)
        sd = Wt.load_seeded(blk, seed=dim + heads)
        x = rnd(2, dim, H, W, seed=dim).requires_grad_(True)
        psd = {"b." + k_: v.clone().requires_grad_(True) for k_, v in sd.items()}
        ref = (O.res_fusion_block if fusion else O.transformer_block)(psd, "b", x, heads)
        dout = rnd(*ref.shape, seed=dim + 7)
        names = list(psd)
        gref = torch.autograd.grad(ref, [x] + [psd[n] for n in names], dout)
        blk = blk.to(DEV)
        p = TR.prep_block_train(blk, A._prep_block(blk))
        tape = []
        y, _ = TR.run_block_train(nhwc(x.detach()), p, tape)
        tag = f"d{dim}_h{heads}_{ln}_b{int(bias)}_f{int(fusion)}"
        out.append(result(f"block_train_fwd_{tag}", nchw(y), ref, 1.5e-2))
        G = TR.Grads()
        dx = TR.run_block_bwd(nhwc(dout), tape[0], G)
        out.append(grad_result(f"block_bwd_dx_{tag}", nchw(dx), gref[0], 3e-2))
        params = dict(blk.named_parameters())
        for n, gr in zip(names, gref[1:]):
            g = G.get(params[n[2:]])
            if g is None:
                out.append(dict(name=f"block_bwd_{tag}_{n[2:]}", ok=False, max_err=None, note="no gradient produced"))
            else:
                out.append(grad_result(f"block_bwd_{tag}_{n[2:]}", g, gr, 4e-2))
    return out


def _net_grads(net, names_ref):
    return {n: p.grad for n, p in net.named_parameters()}


def check_restormer_grad():
    """End-to-end: L1 loss -> loss.backward() through the explicit backward schedule vs autograd through the oracle."""
    from oracle import restormer as O, weights as Wt
    from textualdegremoval_b200.archs import define_network
    out = []
    for name in ("restormer_withbias", "restormer_biasfree", "restormer_gray_bias"):
        meta, _ = _golden(name)
        net = define_network(dict(type="Restormer", **meta["cfg"]))
        sd = Wt.load_seeded(net, meta["seed"])
        x = Wt.seeded_image("x", meta["shape"], meta["seed"])
        gt = Wt.seeded_image("gt", meta["shape"], meta["seed"])
        sdg = {k_: v.clone().requires_grad_(True) for k_, v in sd.items()}
        yr = O.restormer_forward(sdg, x, meta["cfg"]["heads"])
        lr = (yr - gt).abs().mean()
        lr.backward()
        net = net.to(DEV).train()
        y = net(x.to(DEV))
        loss = (y - gt.to(DEV)).abs().mean()
        loss.backward()
        out.append(result(f"train_fwd_{name}", y.detach().cpu(), yr.detach(), E2E_TOL / max(yr.abs().max().item(), 1e-6)))
        out.append(result(f"train_loss_{name}", loss.detach().reshape(1), lr.detach().reshape(1), 1e-2))
        num = den = 0.0
        worst = ("", 0.0)
        for n, p in net.named_parameters():
            gr = sdg[n].grad
            if p.grad is None:
                out.append(dict(name=f"grad_{name}_{n}", ok=False, max_err=None, note="no gradient"))
                continue
            g = p.grad.float().cpu()
            e = (g - gr).norm().item()
            num += e * e
            den += gr.norm().item() ** 2
            rel = e / max(gr.norm().item(), 1e-12)
            if rel > worst[1]:
                worst = (n, rel)
        tot = (num / max(den, 1e-30)) ** 0.5
        r = dict(name=f"grad_global_{name}", max_err=tot, tol=5e-2, ok=bool(tot <= 5e-2), ref_scale=den ** 0.5,
                 note=f"global rel-L2 over all parameters; worst tensor {worst[0]} rel {worst[1]:.3f}")
        out.append(r)
        out.append(dict(name=f"grad_worst_{name}", max_err=worst[1], tol=0.25, ok=bool(worst[1] <= 0.25),
                        note=f"worst per-tensor rel-L2: {worst[0]}"))
    return out


def _grad_compare(tag, net, sdg, out, groups=None):
    """Appends global / worst / per-group rel-L2 between net.parameters().grad and the oracle's autograd grads."""
    num = den = 0.0
    worst = ("", 0.0)
    gsum = {}
    for n, p in net.named_parameters():
        gr = sdg[n].grad
        if gr is None:
            gr = torch.zeros_like(sdg[n])
        if p.grad is None:
            out.append(dict(name=f"grad_{tag}_{n}", ok=False, max_err=None, note="no gradient"))
            continue
        g = p.grad.float().cpu()
        e2 = (g - gr).pow(2).sum().item()
        r2 = gr.pow(2).sum().item()
        num += e2
        den += r2
        rel = (e2 / max(r2, 1e-30)) ** 0.5
        if rel > worst[1] and r2 > 0:
            worst = (n, rel)
        key = n.split(".")[0]
        a = gsum.setdefault(key, [0.0, 0.0])
        a[0] += e2
        a[1] += r2
    tot = (num / max(den, 1e-30)) ** 0.5
    return tot, worst, {k_: (v[0] / max(v[1], 1e-30)) ** 0.5 for k_, v in gsum.items()}


def check_guided_grad():
    """RestormerRefFusion end-to-end: L1 loss backward through fusion blocks, MASA transfer / confidence and the shared
    feature encoder vs autograd through the oracle (the arg-max matches are constants on both sides; a few near-tied
    fine matches flip with bf16 features, which perturbs the guidance gradients inside the affected 8x8 patches)."""
    from oracle import restormer as O, weights as Wt
    from oracle.make_golden import guided_inputs
    from textualdegremoval_b200.archs import define_network
    out = []
    for name in ("guided_restormer_128", "guided_restormer_ragged", "guided_restormer_dual"):
        meta, _ = _golden(name)
        net = define_network(dict(type="RestormerRefFusion", **meta["cfg"]))
        sd = Wt.load_seeded(net, meta["seed"])
        lq, rf = guided_inputs(meta)
        gshape = (meta["lq"][0], meta["cfg"].get("out_channels", meta["lq"][1])) + tuple(meta["lq"][2:])
        gt = Wt.seeded_image("gt", gshape, meta["seed"])
        sdg = {k_: v.clone().requires_grad_(True) for k_, v in sd.items()}
        yr = O.restormer_ref_fusion_forward(sdg, lq, rf, meta["cfg"]["heads"])
        lr = (yr - gt).abs().mean()
        lr.backward()
        net = net.to(DEV).train()
        y = net(lq.to(DEV), rf.to(DEV))
        loss = (y - gt.to(DEV)).abs().mean()
        loss.backward()
        out.append(guided_result(f"train_fwd_{name}", y.detach().cpu(), yr.detach()))
        tot, worst, groups = _grad_compare(name, net, sdg, out)
        body = {k_: v for k_, v in groups.items() if "masa" not in k_}
        masa = {k_: v for k_, v in groups.items() if "masa" in k_}
        note = "per-module rel-L2: " + ", ".join(f"{k_}={v:.3f}" for k_, v in sorted(groups.items(), key=lambda kv: -kv[1])[:6])
        out.append(dict(name=f"grad_global_{name}", max_err=tot, tol=0.02, ok=bool(tot <= 0.02), note=note))
        wb = max(body.values())
        out.append(dict(name=f"grad_body_{name}", max_err=wb, tol=0.05, ok=bool(wb <= 0.05), note="worst non-masa module"))
        wm = max(masa.values())
        out.append(dict(name=f"grad_masa_{name}", max_err=wm, tol=0.06, ok=bool(wm <= 0.06),
                        note=f"worst masa module (match flips perturb these); worst tensor {worst[0]} {worst[1]:.3f}"))
    return out


def check_train_step():
    """RefGuidedTrainer.optimize_parameters (forward, L1, backward, clip 0.01, AdamW with lr / ref_lr groups) for three
    iterations vs the same loop in PyTorch over the oracle (autograd + clip_grad_norm_ + torch.optim.AdamW)."""
    from oracle import restormer as O, weights as Wt
    from oracle.make_golden import guided_inputs
    from textualdegremoval_b200.archs import define_network
    from textualdegremoval_b200.ddp import RefGuidedTrainer
    out = []
    meta, _ = _golden("guided_restormer_128")
    net = define_network(dict(type="RestormerRefFusion", **meta["cfg"]))
    sd = Wt.load_seeded(net, meta["seed"])
    lq, rf = guided_inputs(meta)
    gt = Wt.seeded_image("gt", meta["lq"], meta["seed"])
    lr, ref_lr, wd = 2e-3, 1e-3, 1e-4
    sdg = {k_: v.clone().requires_grad_(True) for k_, v in sd.items()}
    opt = torch.optim.AdamW([dict(params=[v for k_, v in sdg.items() if "masa" not in k_], lr=lr),
                             dict(params=[v for k_, v in sdg.items() if "masa" in k_], lr=ref_lr)],
                            lr=lr, weight_decay=wd, betas=(0.9, 0.999))
    ref_losses = []
    for _ in range(3):
        opt.zero_grad()
        l = (O.restormer_ref_fusion_forward(sdg, lq, rf, meta["cfg"]["heads"]) - gt).abs().mean()
        l.backward()
        torch.nn.utils.clip_grad_norm_(list(sdg.values()), 0.01)
        opt.step()
        ref_losses.append(l.item())
    net = net.to(DEV).train()
    tr = RefGuidedTrainer(net, dict(optim_g=dict(type="AdamW", lr=lr, ref_lr=ref_lr, weight_decay=wd, betas=[0.9, 0.999]),
                                    use_grad_clip=True, pixel_opt=dict(type="L1Loss", loss_weight=1.0)))
    tr.feed_train_data(dict(lq=lq, gt=gt, ref_in=rf))
    losses = []
    for it in range(3):
        tr.optimize_parameters(it)
        losses.append(tr.current_loss())
    out.append(result("train_step_losses", torch.tensor(losses), torch.tensor(ref_losses), 5e-3,
                      note=f"ours {losses} ref {ref_losses}"))
    out.append(dict(name="train_step_loss_decreases", ok=bool(losses[2] < losses[0]), max_err=losses[2] - losses[0], tol=0.0))
    # parameter updates after 3 steps: cosine similarity with the reference trajectory, per LR group
    for grp, sel in (("normal", lambda n: "masa" not in n), ("masa", lambda n: "masa" in n)):
        a = torch.cat([(p.detach().cpu() - sd[n]).reshape(-1) for n, p in net.named_parameters() if sel(n)])
        b = torch.cat([(sdg[n].detach() - sd[n]).reshape(-1) for n, _ in net.named_parameters() if sel(n)])
        cos = (a @ b / (a.norm() * b.norm())).item()
        out.append(dict(name=f"train_step_update_cos_{grp}", max_err=1 - cos, tol=0.05, ok=bool(cos > 0.95),
                        note=f"|dp| ours {a.norm().item():.4e} ref {b.norm().item():.4e}"))
    # fix_iterations: the masa group is frozen (no update, excluded from the clipped norm)
    before = {n: p.detach().clone() for n, p in net.named_parameters()}
    tr.fix_iterations = 10
    tr.optimize_parameters(3)
    tr.fix_iterations = None
    moved_masa = max((p.detach() - before[n]).abs().max().item() for n, p in net.named_parameters() if "masa" in n)
    moved_body = max((p.detach() - before[n]).abs().max().item() for n, p in net.named_parameters() if "masa" not in n)
    out.append(dict(name="fix_iterations_freezes_masa", ok=bool(moved_masa == 0.0 and moved_body > 0.0), max_err=moved_masa,
                    tol=0.0, note=f"body moved {moved_body:.2e}"))
    # DINO reference-crop selection inside the step (full reference image in the batch, image_restoration_ref_model.py:215-247)
    from textualdegremoval_b200.archs import vit_b200 as VB
    ext = VB.DinoVisionTransformer(img_size=70, patch_size=14, embed_dim=64, depth=2, num_heads=4, mlp_ratio=4,
                                   init_values=1.0, ffn_layer="mlp", block_chunks=0)
    Wt.load_seeded(ext, 7)
    ext = ext.to(DEV).eval()
    tr.net_ext = ext
    big = torch.cat([rf, torch.roll(rf, 17, 3)], 3)                     # [B,3,128,256]: N = 1 x 5 candidate crops
    tr.feed_train_data(dict(lq=lq, gt=gt, ref=big))
    tr.optimize_parameters(4)
    sel, idx, cos = VB.select_reference_crop(ext, lq.to(DEV), big.to(DEV))
    out.append(result("dino_selected_ref_in", tr.ref_in, sel, 0.0, note=f"selected crops {idx.tolist()}"))
    l_sel = tr.current_loss()
    out.append(dict(name="train_step_with_dino_select_finite", ok=bool(np.isfinite(l_sel)), max_err=l_sel, tol=None))
    tr.net_ext = None
    tr.feed_train_data(dict(lq=lq, gt=gt, ref_in=rf))
    # plain autograd route (loss.backward() on the module output) still works with the flat grads in place
    y = net(lq.to(DEV), rf.to(DEV))
    tr.engine.zero_grad()
    (y - gt.to(DEV)).abs().mean().backward()
    gn = torch.cat([g.grad[:g.n] for g in tr.engine.groups]).norm().item()
    out.append(dict(name="autograd_into_flat_buffers", ok=bool(gn > 0 and np.isfinite(gn)), max_err=gn, tol=None))
    return out


def check_nafnet_grad():
    """NAFNet / NAFNetRefFusion: L1 loss backward through the explicit schedule vs autograd through the oracle."""
    from oracle import nafnet as ON, weights as Wt
    from oracle.make_golden import denoise_inputs, guided_inputs
    from textualdegremoval_b200.archs import define_network
    out = []
    for name in ("nafnet_tiny_gray64", "nafnet_rgb_ragged"):
        meta, _ = _golden(name)
        net = define_network(dict(type="NAFNet", **meta["cfg"]))
        sd = Wt.load_seeded(net, meta["seed"])
        lq, gt = denoise_inputs(meta)
        sdg = {k_: v.clone().requires_grad_(True) for k_, v in sd.items()}
        yr = ON.nafnet_forward(sdg, lq)
        (yr - gt).abs().mean().backward()
        net = net.to(DEV).train()
        y = net(lq.to(DEV))
        (y - gt.to(DEV)).abs().mean().backward()
        out.append(result(f"train_fwd_{name}", y.detach().cpu(), yr.detach(), E2E_TOL / max(yr.abs().max().item(), 1e-6)))
        tot, worst, groups = _grad_compare(name, net, sdg, out)
        note = f"worst tensor {worst[0]} {worst[1]:.3f}; " + ", ".join(f"{k_}={v:.3f}" for k_, v in sorted(groups.items(), key=lambda kv: -kv[1])[:4])
        out.append(dict(name=f"grad_global_{name}", max_err=tot, tol=0.02, ok=bool(tot <= 0.02), note=note))
        wm = max(groups.values())
        out.append(dict(name=f"grad_modules_{name}", max_err=wm, tol=0.06, ok=bool(wm <= 0.06), note="worst module"))
    meta, _ = _golden("guided_nafnet_256")
    net = define_network(dict(type="NAFNetRefFusion", **meta["cfg"]))
    sd = Wt.load_seeded(net, meta["seed"])
    lq, rf = guided_inputs(meta)
    gt = Wt.seeded_image("gt", meta["lq"], meta["seed"])
    sdg = {k_: v.clone().requires_grad_(True) for k_, v in sd.items()}
    yr = ON.nafnet_ref_fusion_forward(sdg, lq, rf)
    (yr - gt).abs().mean().backward()
    net = net.to(DEV).train()
    y = net(lq.to(DEV), rf.to(DEV))
    (y - gt.to(DEV)).abs().mean().backward()
    out.append(naf_unconditional_result("train_fwd_guided_nafnet_256", y.detach().cpu(), yr.detach()))
    tot, worst, groups = _grad_compare("guided_nafnet_256", net, sdg, out)
    note = f"worst tensor {worst[0]} {worst[1]:.3f}; " + ", ".join(f"{k_}={v:.3f}" for k_, v in sorted(groups.items(), key=lambda kv: -kv[1])[:5])
    out.append(dict(name="grad_global_guided_nafnet_256", max_err=tot, tol=0.03, ok=bool(tot <= 0.03), note=note))
    wm = max(groups.values())
    out.append(dict(name="grad_modules_guided_nafnet_256", max_err=wm, tol=0.1, ok=bool(wm <= 0.1), note="worst module"))
    return out


def check_fullsize_properties():
    """BASELINE.json's full configuration (option 003, 512x512) has no CPU-oracle answer in test time; it is covered by
    size-independent properties: (i) identity at alpha = 0 (SURVEY 8c i) -- the guided net equals the unguided Restormer
    built from its non-masa weights; (ii) samples are independent units -- a batch equals its samples run one by one;
    (iii) the backward pass is linear in the output gradient and the step is deterministic."""
    from textualdegremoval_b200.archs import define_network
    out = []
    cfg = dict(inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4,
               heads=[1, 2, 4, 8], ffn_expansion_factor=2.66, bias=False, LayerNorm_type="WithBias", nf=48,
               ext_n_blocks=[4, 4, 4, 4], reffusion_n_blocks=[2, 2, 2, 2])
    torch.manual_seed(0)
    net = define_network(dict(type="RestormerRefFusion", **cfg))
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("temperature"):
                p.uniform_(0.5, 1.5)
    g = torch.Generator().manual_seed(5)
    lq = torch.rand(2, 3, 512, 512, generator=g).to(DEV)
    ref = torch.rand(2, 3, 512, 512, generator=g).to(DEV)
    net = net.to(DEV).eval()
    # (i) alpha = 0 (the reference's init): guided == unguided on the same non-masa weights
    plain = define_network(dict(type="Restormer", **{k_: v for k_, v in cfg.items()
                                                     if k_ not in ("nf", "ext_n_blocks", "reffusion_n_blocks")}))
    plain.load_state_dict({k_: v for k_, v in net.state_dict().items() if "masa" not in k_}, strict=True)
    plain = plain.to(DEV).eval()
    with torch.no_grad():
        y0 = net(lq[:1], ref[:1])
        yp = plain(lq[:1])
    out.append(result("fullsize_identity_at_zero_alpha", y0, yp, 1e-5))
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("alpha"):
                p.uniform_(0.2, 1.0)
        yb = net(lq, ref)
        y1 = torch.cat([net(lq[i:i + 1], ref[i:i + 1]) for i in range(2)])
    out.append(result("fullsize_batch_equals_samples", yb, y1, 0.0,
                      note="bit-identical: no reduction order depends on the batch size (the MDTA Gram split is a function "
                           "of (P, heads) only)"))
    out.append(dict(name="fullsize_guidance_changes_output", ok=bool((yb[:1] - y0).abs().max().item() > 1e-3),
                    max_err=(yb[:1] - y0).abs().max().item(), tol=None))
    # (iii) backward: deterministic and linear in dout
    net.train()
    gt = torch.rand(1, 3, 512, 512, generator=g).to(DEV)

    def grads(scale):
        for p in net.parameters():
            p.grad = None
        y = net(lq[:1], ref[:1])
        ((y - gt).abs().mean() * scale).backward()
        return torch.cat([p.grad.reshape(-1) for p in net.parameters()])

    g1, g1b, g4 = grads(1.0), grads(1.0), grads(4.0)
    det = (g1 - g1b).abs().max().item()
    out.append(dict(name="fullsize_backward_repeatable", ok=bool(det <= 1e-6 * g1.abs().max().item() + 1e-12), max_err=det, tol=None,
                    note="only the MASA scatter-adds use fp32 atomics"))
    out.append(grad_result("fullsize_backward_linear_in_dout", g4, 4.0 * g1, 2e-3))
    out.append(dict(name="fullsize_all_params_get_gradients", ok=bool(all(p.grad is not None and torch.isfinite(p.grad).all()
                                                                         and p.grad.abs().max() > 0 for p in net.parameters())),
                    max_err=None, tol=None))
    return out

def _book_agreement(tag, aux, z):
    """Arg-max agreement against a full-size fixture's match book (top-3 candidates + fp32 scores of every coarse /
    fine arg-max, written by oracle/make_golden_fullsize.py): a mismatch is a tie when our candidate is one of the
    oracle's top-3 within TIE_TOL of its maximum."""
    cv, ci = torch.from_numpy(z["coarse_top_val"]), torch.from_numpy(z["coarse_top_idx"]).long()
    fv, fi = torch.from_numpy(z["fine_top_val"]), torch.from_numpy(z["fine_top_idx"]).long()
    B, nblk, _ = ci.shape

    def grade(mine, top_i, top_v):
        hit = (top_i == mine.unsqueeze(-1)) & ((top_v[..., :1] - top_v) <= TIE_TOL)
        return (top_i[..., 0] == mine), hit.any(-1)

    ex_c, ok_c = grade(aux["idx"].cpu().long().view(B, nblk), ci, cv)
    y1 = aux["origin"][:, 1].view(B, -1).cpu().long()
    x1 = aux["origin"][:, 2].view(B, -1).cpu().long()
    same_win = ((y1 == torch.from_numpy(z["y1"]).long()) & (x1 == torch.from_numpy(z["x1"]).long())).view(-1)
    nq = fi.shape[1]
    ex_f, ok_f = grade(aux["index"].cpu().long().view(-1, nq), fi, fv)
    a_c, a_f = ok_c.float().mean().item(), ok_f[same_win].float().mean().item()
    return [dict(name=f"{tag}_coarse_idx", max_err=1 - a_c, ref_scale=1, tol=0.0, ok=a_c == 1.0,
                 note=f"coarse arg-max over {B * nblk} blocks: identical {ex_c.float().mean().item():.4f}, up to ties "
                      f"< {TIE_TOL:g} {a_c:.4f}"),
            dict(name=f"{tag}_fine_idx", max_err=1 - a_f, ref_scale=1, tol=1 - FINE_AGREE_MIN, ok=a_f >= FINE_AGREE_MIN,
                 note=f"fine arg-max over {int(same_win.sum()) * nq} positions: identical "
                      f"{ex_f[same_win].float().mean().item():.4f}, up to ties {a_f:.4f} (bar 99.9 %)")]


def check_fullsize_golden():
    """The BASELINE.json configurations at FULL size against outputs of the unmodified reference modules
    (tests/golden/full_*.npz, oracle/make_golden_fullsize.py): guided Restormer option 003 @512x512 (the headline
    config), guided NAFNet option 002 @512x512, Restormer option 017 @256x256, DINOv2 ViT-B/14 @518x518, CLIP ViT-H/14
    @224x224 (third-party transformers; version in the fixture)."""
    from oracle import weights as Wt
    from oracle.make_golden_fullsize import fullsize_inputs
    from textualdegremoval_b200.archs import define_network, vit_b200 as VB
    out = []
    for name in ("full_guided_restormer_512", "full_guided_nafnet_512", "full_restormer_256", "full_dino_vitb_518",
                 "full_clip_vith_224"):
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        meta, ref = json.loads(str(z["meta"])), torch.from_numpy(z["out"])
        kind = meta["kind"]
        lq, rf, _ = fullsize_inputs(meta)
        if kind in ("guided_restormer", "guided_nafnet", "restormer"):
            typ = dict(guided_restormer="RestormerRefFusion", guided_nafnet="NAFNetRefFusion", restormer="Restormer")[kind]
            net = define_network(dict(type=typ, **meta["cfg"]))
        elif kind == "dino":
            net = VB.vit_base(**meta["cfg"])
        else:
            net = VB.CLIPVisionTower(**meta["cfg"])
        Wt.load_seeded(net, meta["seed"])
        net = net.to(DEV).eval()
        with torch.no_grad():
            if kind in ("guided_restormer", "guided_nafnet"):
                y, aux = net(lq.to(DEV), rf.to(DEV), return_aux=True)
                out += _book_agreement(name, aux, z)
                if kind == "guided_nafnet":         # parity GIVEN the oracle's matches is the asserted one (see _ForcedMatches)
                    out.append(naf_unconditional_result(f"golden_{name}_own_matches", y.cpu(), ref))
                    with _ForcedMatches(torch.from_numpy(z["y1"]), torch.from_numpy(z["x1"]),
                                        torch.from_numpy(z["fine_top_idx"][..., 0])):
                        y = net(lq.to(DEV), rf.to(DEV))
            elif kind == "clip":
                y = net(lq.to(DEV), output_hidden_states=True)[0]
            else:
                y = net(lq.to(DEV))
        y = y.cpu()
        if kind in ("dino", "clip"):
            # token features (O(1-5) values after the final / no final norm): relative to the fixture's scale
            r = result(f"golden_{name}", y, ref, 2e-2)
            rel = ((y - ref).norm() / ref.norm()).item()
            r["note"] = f"max|d|={r['max_err']:.2e} of scale {r['ref_scale']:.2f}; rel-L2 {rel:.2e} (<= 5e-3)"
            r["ok"] = r["ok"] and rel <= 5e-3
            out.append(r)
        else:
            out.append(guided_result(f"golden_{name}", y, ref))
            out.append(psnr_delta_result(name, y, ref, meta["seed"]))
        del net
        torch.cuda.empty_cache()
    return out


def check_graph_replay():
    """GraphedForward (CUDA-graph replay of the inference forward) returns the eager path's bits, for new inputs too, and
    refuses a different shape.  The timing note is informational (one image: the coarse levels are launch-bound)."""
    from textualdegremoval_b200.archs import define_network
    from textualdegremoval_b200.graphs import GraphedForward
    from oracle.make_golden import GUIDED_CASES
    out = []
    meta = GUIDED_CASES["guided_restormer_128"]
    torch.manual_seed(3)
    net = define_network(dict(type="RestormerRefFusion", **meta["cfg"]))
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("alpha"):
                p.uniform_(0.2, 1.0)
    net = net.to(DEV).eval()
    g = torch.Generator().manual_seed(11)
    xs = [(torch.rand(1, 3, 128, 128, generator=g).to(DEV), torch.rand(1, 3, 128, 128, generator=g).to(DEV))
          for _ in range(3)]
    with torch.no_grad():
        eager = [net(a, b).clone() for a, b in xs]
    fwd = GraphedForward(net, *xs[0])
    for i, (a, b) in enumerate(xs):
        out.append(result(f"graph_replay_equals_eager_{i}", fwd(a, b).clone(), eager[i], 0.0, note="bit-identical"))
    try:
        fwd(torch.rand(1, 3, 64, 64, device=DEV), xs[0][1])
        refused = False
    except ValueError:
        refused = True
    out.append(dict(name="graph_replay_refuses_other_shape", max_err=0.0, tol=0.0, ok=refused, note=""))

    def ms(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n
    with torch.no_grad():
        t_eager = ms(lambda: net(*xs[0]))
    t_graph = ms(lambda: fwd(*xs[0]))
    out.append(dict(name="graph_replay_timing_128", max_err=0.0, tol=0.0, ok=True,
                    note=f"1 x 128x128 guided forward: eager {t_eager:.3f} ms, graph replay {t_graph:.3f} ms"))
    return out


CHECKS = {
    "graph_replay": check_graph_replay,
    "layout": check_layout,
    "rownorm": check_rownorm,
    "dwconv": check_dwconv,
    "small_convs": check_small_convs,
    "conv_simt": check_conv_simt,
    "conv_tc_basic": check_conv_tc_basic,
    "conv_tc": check_conv_tc,
    "conv_origin": check_conv_origin,
    "conv_ln": check_conv_ln,
    "gdfn_tail": check_gdfn_tail,
    "fp16_path": check_fp16_path,
    "input_pipeline": check_input_pipeline,
    "psnr": check_psnr,
    "mdta": check_mdta,
    "block": check_block,
    "masa": check_masa,
    "restormer_golden": check_restormer_golden,
    "guided_stages": check_guided_stages,
    "guided_golden": check_guided_golden,
    "nafnet": check_nafnet,
    "promptir": check_promptir,
    "drsformer": check_drsformer,
    "vit": check_vit,
    "optim": check_optim,
    "wgrad": check_wgrad,
    "bwd_pointwise": check_bwd_pointwise,
    "block_bwd": check_block_bwd,
    "restormer_grad": check_restormer_grad,
    "guided_grad": check_guided_grad,
    "train_step": check_train_step,
    "nafnet_grad": check_nafnet_grad,
    "fullsize_properties": check_fullsize_properties,
    "fullsize_golden": check_fullsize_golden,
}


def run_group(name):
    t0 = time.time()
    try:
        res = CHECKS[name]()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        res = [dict(name=name, ok=False, max_err=None, note="EXCEPTION: " + "".join(
            traceback.format_exception_only(type(e), e)).strip()[-600:])]
    for r in res:
        r["group"] = name
    return res, time.time() - t0


def main(argv):
    only = [a for a in argv if not a.startswith("--")]
    isolate = "--isolate" in argv
    names = only or list(CHECKS)
    allres = []
    for n in names:
        if isolate:
            try:
                p = subprocess.run([sys.executable, "-m", "tests.gpu_checks", "--json", n], capture_output=True, text=True,
                                   timeout=600, cwd=ROOT)
                line = [l for l in p.stdout.splitlines() if l.startswith("JSON:")]
                res = json.loads(line[-1][5:]) if line else [dict(name=n, group=n, ok=False, max_err=None,
                                                                 note="CRASH: " + (p.stderr or p.stdout)[-800:])]
            except subprocess.TimeoutExpired:
                res = [dict(name=n, group=n, ok=False, max_err=None, note="TIMEOUT")]
        else:
            res, _ = run_group(n)
        allres += res
        for r in res:
            err = "-" if r.get("max_err") is None else f"{r['max_err']:.3e}"
            tol = "-" if r.get("tol") is None else f"{r['tol']:.3e}"
            print(f"[{'ok' if r['ok'] else 'FAIL'}] {r['name']:<44} err {err:>10} tol {tol:>10} {r.get('note', '')}", flush=True)
    if "--json" in argv:
        print("JSON:" + json.dumps(allres))
    else:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "gpu_checks.json"), "w") as fh:
            json.dump(allres, fh, indent=1)
        bad = [r["name"] for r in allres if not r["ok"]]
        print(f"\n{len(allres) - len(bad)}/{len(allres)} checks ok; failing: {bad}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
