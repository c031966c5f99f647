"""Tensor-level wrappers over the C ABI (include/tdr_sm100.h).

Activations are NHWC torch tensors ``[B, H, W, C]`` whose channel dim is contiguous; channel slices of a wider
buffer (``t[..., a:b]``) are passed as (pointer, row stride) without copies, which is how concatenations
(``torch.cat([x, warp], 1)`` in the reference) are expressed.  PyTorch is used for device memory and streams only.
"""
import ctypes as C
import os

import torch

from . import lib

BF16 = torch.bfloat16
F16 = torch.float16
F32 = torch.float32


def operand_dtype(train):
    """16-bit format of the GEMM / depthwise operands and stored intermediates.  Inference: IEEE fp16 -- with bf16's 8
    significand bits the full-size guided-Restormer output sits at mean |delta| 1.05e-3 / 0.032 dB PSNR-delta from the fp32
    reference, with fp16's 11 bits it meets the 1e-3 / 0.01 dB bar (tools/operand_format_study.py; same tcgen05 rate,
    stores saturate at +-65504).  Training: bf16 -- the saved activations are operands of the bf16 gradient GEMMs (tcgen05
    kind::f16 takes no mixed fp16 x bf16 pairs) and gradients need bf16's exponent range."""
    return BF16 if train else F16


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Profiler:
    """Optional per-launch CUDA-event timing (bench.py roofline leg).  Off by default: zero overhead on the hot path
    beyond one counter increment per launch."""

    def __init__(self):
        self.active = False
        self.records = []          # (kernel, tag, bytes, flops, start_event, end_event)
        self.launches = 0

    def start(self):
        self.records, self.active = [], True

    def stop(self):
        self.active = False
        torch.cuda.synchronize()
        out = [(k, t, b, f, s.elapsed_time(e)) for (k, t, b, f, s, e) in self.records]
        self.records = []
        return out


PROF = Profiler()


def _call(name, *args, tag="", nbytes=0, flops=0):
    PROF.launches += 1
    if PROF.active:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        lib.call(name, *args)
        e.record()
        PROF.records.append((name, tag, nbytes, flops, s, e))
    else:
        lib.call(name, *args)


def _nb(*ts):
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


def _ld(t: torch.Tensor) -> int:
    """Row stride (elements) of an NHWC view; checks the layout is a channel slice of a dense NHWC buffer."""
    assert t.dim() == 4 and t.stride(3) == 1, f"expected NHWC view, got strides {t.stride()}"
    ld = t.stride(2)
    b, h, w, _ = t.shape
    assert (w == 1 or t.stride(2) == ld) and (h == 1 or t.stride(1) == w * ld) and (b == 1 or t.stride(0) == h * w * ld), \
        f"not a channel slice of a dense NHWC buffer: shape {tuple(t.shape)} strides {t.stride()}"
    return ld


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ---------------------------------------------------------------------------------------------- weight packing
def _f32c(t):
    t = t.detach()
    return t if (t.dtype == F32 and t.is_contiguous()) else t.float().contiguous()


def pack_conv(w, co_map=None, Co_p=None, ci_map=None, Ci_p=None, scale=None, fwd=True, dgrad=False, dt=BF16):
    """nn.Conv2d weight [Co, Ci, kh, kw] -> (bf16 [T, Co_p, ld], bf16 [T, Ci_p, ld_t]): the tdr_conv_gemm operand and its
    transposed, tap-flipped twin (data gradient), each in ONE kernel launch.  co_map / ci_map: int32 DEVICE tensors mapping
    padded channel -> logical channel (-1 = zero); scale: fp32 [Co] folded into the output channels.  dt: 16-bit format
    of the forward operand (torch.float16 for the inference forward); the twin is always bf16."""
    w = _f32c(w)
    co, ci, kh, kw = w.shape
    Co_p = co if Co_p is None else Co_p
    Ci_p = ci if Ci_p is None else Ci_p
    ld, ld_t = round_up(Ci_p, 8), round_up(Co_p, 8)
    out = torch.empty((kh * kw, Co_p, ld), dtype=dt, device=w.device) if fwd else None
    out_t = torch.empty((kh * kw, Ci_p, ld_t), dtype=BF16, device=w.device) if dgrad else None
    lib.call("tdr_pack_conv_weight", _p(w), co, ci, kh, kw, _p(co_map), Co_p, _p(ci_map), Ci_p, _p(scale), _p(out), ld,
             _p(out_t), ld_t, int(dt == F16), _stream())
    return out, out_t


def pack_conv_weight(w: torch.Tensor, ci_map=None, co_map=None, dt=BF16) -> torch.Tensor:
    """[Co, Ci, kh, kw] fp32 -> bf16 [kh*kw, Co, Ci_p] (tap-major, K contiguous, Ci padded to 8) for tdr_conv_gemm.
    Rows are NOT padded: the kernel's TMA box zero-fills beyond Co.

    ci_map / co_map: optional (n_padded, index tensor) placing logical channels at padded positions (GDFN halves).
    """
    if ci_map is None and co_map is None and w.dim() == 4 and w.is_cuda:
        return pack_conv(w, dt=dt)[0]
    assert dt == BF16
    co, ci, kh, kw = w.shape
    wt = w.detach().permute(2, 3, 0, 1).reshape(kh * kw, co, ci)
    if ci_map is not None:
        n, idx = ci_map
        tmp = wt.new_zeros(kh * kw, co, n)
        tmp[:, :, idx] = wt
        wt = tmp
    if co_map is not None:
        n, idx = co_map
        tmp = wt.new_zeros(kh * kw, n, wt.shape[2])
        tmp[:, idx, :] = wt
        wt = tmp
    cip = round_up(wt.shape[2], 8)
    out = torch.zeros(kh * kw, wt.shape[1], cip, dtype=BF16, device=w.device)
    out[:, :, : wt.shape[2]] = wt.to(BF16)
    return out


def pack_dw(w, bias=None, c_map=None, C_p=None, flip=False):
    """Depthwise weight [C,1,3,3] (+bias) -> (fp32 [9, C_p], flipped twin or None, padded bias or None), one launch."""
    w = _f32c(w)
    c = w.shape[0]
    C_p = c if C_p is None else C_p
    out = torch.empty((9, C_p), dtype=F32, device=w.device)
    out_f = torch.empty((9, C_p), dtype=F32, device=w.device) if flip else None
    out_b = torch.empty(C_p, dtype=F32, device=w.device) if bias is not None else None
    lib.call("tdr_pack_dw_weight", _p(w), _p(_f32c(bias)) if bias is not None else None, c, _p(c_map), C_p, _p(out),
             _p(out_f), _p(out_b), _stream())
    return out, out_f, out_b


def gather_vec(v, cmap, n):
    """out[i] = v[cmap[i]] (0 where cmap[i] < 0): padded bias vectors."""
    if v is None:
        return None
    v = _f32c(v).reshape(-1)
    out = torch.empty(n, dtype=F32, device=v.device)
    lib.call("tdr_gather_vec", _p(v), _p(cmap), n, v.numel(), _p(out), _stream())
    return out


def pad_vec(v, n, idx=None):
    if v is None:
        return None
    out = torch.zeros(n, dtype=F32, device=v.device)
    if idx is None:
        out[: v.numel()] = v.detach().float().reshape(-1)
    else:
        out[idx] = v.detach().float().reshape(-1)
    return out


def pack_dw_weight(w: torch.Tensor, n=None, idx=None) -> torch.Tensor:
    """[C, 1, 3, 3] -> fp32 [9, C_p] tap-major."""
    c = w.shape[0]
    wt = w.detach().float().reshape(c, 9).t().contiguous()
    if n is None:
        return wt
    out = torch.zeros(9, n, dtype=F32, device=w.device)
    out[:, idx] = wt
    return out


# widths kept dense by rows16 (TDR_ROWS16_DENSE="288,576": A/B knob for the pitch trade-off between the GEMM stores and
# the depthwise conv that reads the same buffer)
_ROWS16_DENSE = frozenset(int(v) for v in os.environ.get("TDR_ROWS16_DENSE", "").split(",") if v.strip())


def rows16(B, H, W, Cc, dev, dt=BF16):
    """bf16 NHWC activation buffer whose row pitch is a multiple of 128 B (64 channels): a view [B,H,W,Cc] of a wider
    allocation.  Rows that straddle 128 B lines (C = 48, 96, 144, 288 ...) cost the TMA loads / stores of the GEMMs up to
    25 % of their bandwidth (tools/probe.py qkv48 vs qkv48_ld192); the pad columns are never read or written."""
    ld = round_up(Cc, 64)
    if Cc in _ROWS16_DENSE:
        ld = Cc
    if ld == Cc:
        return torch.empty((B, H, W, Cc), dtype=dt, device=dev)
    return torch.empty((B, H, W, ld), dtype=dt, device=dev)[..., :Cc]


# ---------------------------------------------------------------------------------------------- ops
def conv_gemm(x, wpack, Co, *, Ci=None, k=1, stride=1, pad=0, dil=1, bias=None, relu=False, gelu=False, rowscale=None,
              alpha=1.0, scale_ptr=None, res1=None, res1_scale=1.0, res2=None, out_f32=None, out_bf16=None,
              want="bf16", store_mode=0, w_batched=False, w_raw=None, origin=None, window=None, impl=0, ln=None,
              out_fp16=None):
    """x: bf16 NHWC view.  wpack: bf16 [T, Co_p, Ci_p].  Returns (out_f32, out_bf16) -- allocated if not given
    according to ``want`` in {"bf16", "f32", "both"}.

    ln = (mode, weight, bias, eps, out_bf16_view): also write LayerNorm(out) of the finished fp32 rows (the norm that
    follows on the residual stream) from the same epilogue; see ``conv_ln_ok``.

    fp16 ``x`` and ``wpack`` (the MASA feature encoder's operands) run the same kernel with fp16 operand descriptors;
    out_fp16: the 16-bit output is IEEE fp16."""
    in_fp16 = x.dtype == F16
    assert x.dtype in (BF16, F16) and (w_raw is not None or wpack.dtype == x.dtype)
    if out_fp16 is None:            # 16-bit outputs keep the operand format unless told otherwise
        out_fp16 = (out_bf16.dtype == F16) if out_bf16 is not None else ((ln[4].dtype == F16) if ln is not None else in_fp16)
    o16 = F16 if out_fp16 else BF16
    B, H, W, Cx = x.shape
    Ci = Cx if Ci is None else Ci
    d = lib.ConvGemmDesc()
    d.in_ = x.data_ptr(); d.in_ld = _ld(x)
    if origin is not None:
        wh, ww = window
        d.origin = origin.data_ptr(); d.n_images = B; d.img_h = H; d.img_w = W
        d.B = origin.shape[0]; d.H = wh; d.W = ww
    else:
        d.B = B; d.H = H; d.W = W
    d.Ci = Ci
    if w_raw is not None:
        # raw per-sample weights living inside another tensor: (data_ptr, row stride, batch stride) in elements
        d.weight, d.w_ld, d.w_batch_stride = w_raw
        assert w_batched and k == 1
    else:
        d.weight = wpack.data_ptr(); d.w_ld = wpack.shape[2]
        assert wpack.shape[1] >= Co and wpack.shape[2] >= Ci and wpack.is_contiguous()
        # the weight tensor map uses Co as the row extent and w_ld*Co as the tap stride
        assert wpack.shape[1] == Co, f"packed weight rows {wpack.shape[1]} != Co {Co}"
    d.Co = Co; d.KH = k; d.KW = k; d.stride = stride; d.pad = pad; d.dil = dil
    d.w_batched = int(w_batched)
    d.bias = bias.data_ptr() if bias is not None else None
    d.rowscale = rowscale.data_ptr() if rowscale is not None else None
    d.alpha = alpha
    d.scale_ptr = scale_ptr.data_ptr() if scale_ptr is not None else None
    d.act = 2 if gelu else int(relu)
    nB, nH, nW = d.B, d.H, d.W
    OH = (nH + 2 * pad - dil * (k - 1) - 1) // stride + 1
    OW = (nW + 2 * pad - dil * (k - 1) - 1) // stride + 1
    if store_mode == 0:
        oshape = (nB, OH, OW, Co)
    elif store_mode == 1:
        oshape = (nB, OH // 2, OW // 2, Co * 4)
    else:
        oshape = (nB, OH * 2, OW * 2, Co // 4)
    if out_f32 is None and want in ("f32", "both"):
        out_f32 = torch.empty(oshape, dtype=F32, device=x.device)
    if out_bf16 is None and want in ("bf16", "both") and not (want == "bf16" and out_f32 is not None):
        out_bf16 = torch.empty(oshape, dtype=o16, device=x.device)
    for o in (out_f32, out_bf16, res1, res2):
        if o is not None:
            assert tuple(o.shape) == oshape, f"output/residual shape {tuple(o.shape)} != {oshape}"
    if res1 is not None:
        assert res1.dtype == F32
        d.res1 = res1.data_ptr(); d.res1_ld = _ld(res1); d.res1_scale = res1_scale
    if res2 is not None:
        d.res2 = res2.data_ptr(); d.res2_ld = _ld(res2); d.res2_bf16 = int(res2.dtype == BF16)
    if out_f32 is not None:
        assert out_f32.dtype == F32
        d.out_f32 = out_f32.data_ptr(); d.out_f32_ld = _ld(out_f32)
    if out_bf16 is not None:
        assert out_bf16.dtype == o16
        d.out_bf16 = out_bf16.data_ptr(); d.out_bf16_ld = _ld(out_bf16)
    d.in_fp16 = int(in_fp16)
    d.out_fp16 = int(bool(out_fp16) and (out_bf16 is not None or ln is not None))
    d.store_mode = store_mode
    d.impl = impl
    ln_out = None
    if ln is not None:
        mode, lw, lb, eps, ln_out = ln
        assert mode in (1, 2) and ln_out.dtype == o16 and tuple(ln_out.shape) == oshape and lw.dtype == F32
        d.ln_mode = mode; d.ln_eps = eps; d.ln_weight = lw.data_ptr()
        d.ln_bias = lb.data_ptr() if (lb is not None and mode == 1) else None
        d.ln_out_bf16 = ln_out.data_ptr(); d.ln_out_ld = _ld(ln_out)
    taps = k * k
    _call("tdr_conv_gemm", C.byref(d), _stream(),
          tag=f"k{k}s{stride}_Ci{Ci}_Co{Co}_{OH}x{OW}" + ("_wb" if w_batched else "") + ("_ln" if ln is not None else "")
              + ("_h" if in_fp16 else ""),
          nbytes=nB * nH * nW * Ci * 2 + Co * Ci * taps * 2 * (nB if w_batched else 1) + _nb(out_f32, out_bf16, res1, res2, ln_out),
          flops=2 * nB * OH * OW * Co * Ci * taps)
    return out_f32, out_bf16


def conv_ln_ok(Co):
    """Can a 1x1 ``conv_gemm`` with fp32 output + fp32 res2 (no res1) also emit the LayerNorm of its output rows?  Mirrors
    ``tdr_conv_gemm_ln_supported`` for the buffers this package allocates.  TDR_NO_LN_FUSION=1 turns the fusion off
    (A/B measurements)."""
    return Co <= 96 and Co % 8 == 0 and os.environ.get("TDR_NO_LN_FUSION", "0") in ("", "0")


def rownorm(x32, mode, weight=None, bias=None, eps=1e-5, out=None, leaky=False, out_f32=None, want_bf16=True, dt=BF16):
    """fp32 NHWC view -> 16-bit NHWC (bf16, or fp16 when dt / out say so) and/or fp32.  mode 0 cast, 1 WithBias LN,
    2 BiasFree LN; leaky: LeakyReLU(0.01)."""
    assert x32.dtype == F32
    B, H, W, Cc = x32.shape
    if out is None and want_bf16:
        out = torch.empty((B, H, W, Cc), dtype=dt, device=x32.device)
    act = (3 if leaky else 0) | (16 if (out is not None and out.dtype == F16) else 0)
    _call("tdr_rownorm", _p(x32), _ld(x32), B * H * W, Cc, mode, _p(weight), _p(bias), eps, act, _p(out),
          _ld(out) if out is not None else 0, _p(out_f32), _ld(out_f32) if out_f32 is not None else 0, _stream(),
          tag=f"m{mode}_C{Cc}", nbytes=B * H * W * Cc * (4 + (2 if out is not None else 0) + (4 if out_f32 is not None else 0)))
    return out if out is not None else out_f32


def dwconv3x3(x, w9, bias=None, gate=0, out=None, relu=False):
    assert x.dtype in (BF16, F16) and w9.dtype == F32 and not (relu and gate)
    B, H, W, Cc = x.shape
    co = Cc // 2 if gate else Cc
    if out is None:
        out = torch.empty((B, H, W, co), dtype=x.dtype, device=x.device)
    assert out.dtype == x.dtype
    _call("tdr_dwconv3x3", _p(x), _ld(x), B, H, W, Cc, _p(w9), _p(bias), gate | (16 if x.dtype == F16 else 0) | (32 if relu else 0),
          _p(out), _ld(out), _stream(),
          tag=f"g{gate}_C{Cc}_{H}x{W}", nbytes=B * H * W * (Cc + co) * 2, flops=2 * 9 * B * H * W * Cc)
    return out


def _gdfn_desc(hid, w9, bias_dw, w_out, Cc, bias, scale_ptr, res1, res1_scale, res2, out):
    B, H, W, C2 = hid.shape
    d = lib.GdfnTailDesc()
    d.hidden = hid.data_ptr(); d.hidden_ld = _ld(hid)
    d.B = B; d.H = H; d.W = W; d.hp = C2 // 2; d.C = Cc
    d.dw_weight = w9.data_ptr(); d.dw_bias = bias_dw.data_ptr() if bias_dw is not None else None
    d.w_out = w_out.data_ptr(); d.w_ld = w_out.shape[-1]
    d.bias = bias.data_ptr() if bias is not None else None
    d.alpha = 1.0
    d.scale_ptr = scale_ptr.data_ptr() if scale_ptr is not None else None
    if res1 is not None:
        d.res1 = res1.data_ptr(); d.res1_ld = _ld(res1)
    d.res1_scale = res1_scale
    if res2 is not None:
        d.res2 = res2.data_ptr(); d.res2_ld = _ld(res2)
    d.out = out.data_ptr(); d.out_ld = _ld(out)
    d.fp16 = int(hid.dtype == F16)
    return d


def gdfn_tail_supported(hid, w_out, Cc):
    """Can ``gdfn_tail`` run this shape (16 <= C <= 192, C % 16 == 0, ...)?"""
    return 16 <= Cc <= 192 and Cc % 16 == 0 and (hid.shape[3] // 2) % 8 == 0 and w_out.dtype == hid.dtype \
        and w_out.shape[-2] == Cc


def gdfn_tail_ok(hid, w_out, Cc):
    """Does the block schedule use the fused GDFN tail?  OFF by default: the kernel is correct (tests/gpu_checks.py
    ``gdfn_tail``) but measured 859 us against 713 us for depthwise-gate + project_out as two kernels at hp 256 -> C 96,
    512x512, batch 4 (tools/probe_gdfn.py; DESIGN.md "Fused GDFN tail: a negative result").  TDR_GDFN_FUSION=1 turns it
    on for A/B runs."""
    return os.environ.get("TDR_GDFN_FUSION", "0") not in ("", "0") and gdfn_tail_supported(hid, w_out, Cc)


def gdfn_tail(hid, w9, bias_dw, w_out, Cc, bias=None, scale_ptr=None, res1=None, res1_scale=1.0, res2=None, out=None):
    """Fused GDFN tail (reference R:236-240, :329, :345-353): out = g * (W_out . (gelu(dw(hid)[:hp]) * dw(hid)[hp:]) + bias)
    + g * res1_scale * res1 + res2, fp32 rows; the gated tensor never reaches HBM.  hid: 16-bit NHWC [B,H,W,2*hp]; w9 fp32
    [9, 2*hp]; w_out 16-bit [1, C, hp]."""
    assert hid.dtype in (BF16, F16) and w9.dtype == F32 and w_out.dtype == hid.dtype
    B, H, W, C2 = hid.shape
    if out is None:
        out = torch.empty((B, H, W, Cc), dtype=F32, device=hid.device)
    d = _gdfn_desc(hid, w9, bias_dw, w_out, Cc, bias, scale_ptr, res1, res1_scale, res2, out)
    hp = C2 // 2
    _call("tdr_gdfn_tail", C.byref(d), _stream(), tag=f"hp{hp}_C{Cc}_{H}x{W}",
          nbytes=B * H * W * (C2 * 2 + Cc * 4 * (1 + int(res1 is not None) + int(res2 is not None))) + Cc * hp * 2,
          flops=2 * B * H * W * (9 * C2 + hp * Cc))
    return out


def mdta_weff(qkv, C_, heads, temperature, w_out, want_attn=False, save=None, topk_w=None):
    """qkv: bf16 NHWC [B,H,W,>=3C].  Returns Weff bf16 [B, C, C_p] (and attn fp32 [B,heads,c,c]).
    save: optional dict that receives partials / attn / weff / weff_t (what the backward pass needs).
    topk_w: optional fp32 [4] device tensor -> top-k sparse attention mix (DRSformer TKSA)."""
    B, H, W, _ = qkv.shape
    P = H * W
    nbytes = lib.load().tdr_mdta_partials_bytes(B, P, C_, heads)
    if nbytes == 0:
        raise lib.TdrError(f"MDTA head width unsupported: C={C_} heads={heads}")
    partials = torch.empty(nbytes // 4, dtype=F32, device=qkv.device)
    h16 = int(qkv.dtype == F16)
    _call("tdr_mdta_gram", _p(qkv), _ld(qkv), B, P, C_, heads, _p(partials), h16, _stream(), tag=f"C{C_}_h{heads}_P{P}",
          nbytes=B * P * 2 * C_ * 2, flops=3 * 2 * B * P * C_ * (C_ // heads))
    cp = round_up(C_, 8)
    weff = torch.empty((B, C_, cp), dtype=qkv.dtype, device=qkv.device)
    attn = torch.empty((B, heads, C_ // heads, C_ // heads), dtype=F32, device=qkv.device)
    weff_t = shat = None
    if save is not None:
        c = C_ // heads
        weff_t = torch.zeros((B, C_, cp), dtype=BF16, device=qkv.device)
        shat = torch.empty((B, heads, c * c + 2 * c), dtype=F32, device=qkv.device)
    _call("tdr_mdta_weff", _p(partials), B, P, C_, heads, _p(temperature), _p(w_out), _p(weff), cp, _p(attn),
          _p(weff_t), _p(shat), h16, _p(topk_w), _stream(), tag=f"C{C_}_h{heads}", nbytes=nbytes + _nb(weff))
    if save is not None:
        save.update(shat=shat, attn=attn, weff=weff, weff_t=weff_t)
    return (weff, attn) if want_attn else weff


def gate_mul(x16, out=None):
    """SimpleGate: bf16 [B,H,W,2C] -> [B,H,W,C]."""
    B, H, W, C2 = x16.shape
    Cc = C2 // 2
    if out is None:
        out = torch.empty((B, H, W, Cc), dtype=x16.dtype, device=x16.device)
    _call("tdr_gate_mul", _p(x16), _ld(x16), B * H * W, Cc, _p(out), _ld(out), int(x16.dtype == F16), _stream(), tag=f"C{Cc}",
          nbytes=B * H * W * Cc * 6)
    return out


def naf_sca_fold(g16, w_sca, b_sca, w3, rowscale=None, save=None):
    """g16: bf16 [B,H,W,C] (gated features).  Returns per-sample folded conv3 weights bf16 [B, Co, C_p].
    save: optional dict receiving mean / s / weff_t (training)."""
    B, H, W, Cc = g16.shape
    Co = w3.shape[0]
    nbytes = lib.load().tdr_naf_sca_workspace_bytes(B, H * W, Cc)
    ws = torch.empty(nbytes // 4, dtype=F32, device=g16.device)
    cp = round_up(Cc, 8)
    weff = torch.empty((B, Co, cp), dtype=g16.dtype, device=g16.device)
    mean = s_vec = weff_t = None
    cop = round_up(Co, 8)
    if save is not None:
        mean = torch.empty((B, Cc), dtype=F32, device=g16.device)
        s_vec = torch.empty((B, Cc), dtype=F32, device=g16.device)
        weff_t = torch.zeros((B, Cc, cop), dtype=BF16, device=g16.device)
        save.update(mean=mean, s=s_vec, weff_t=weff_t)
    _call("tdr_naf_sca_fold", _p(g16), _ld(g16), B, H * W, Cc, _p(w_sca), _p(b_sca), _p(w3), Co, _p(rowscale), _p(weff),
          cp, _p(ws), _p(mean), _p(s_vec), _p(weff_t), cop, int(g16.dtype == F16), _stream(), tag=f"C{Cc}",
          nbytes=B * H * W * Cc * 2)
    return weff


def nchw_to_nhwc(x, pad_h, pad_w, want_bf16=False):
    x = x.contiguous().float()
    B, Cc, H, W = x.shape
    o32 = torch.empty((B, pad_h, pad_w, Cc), dtype=F32, device=x.device)
    o16 = torch.empty((B, pad_h, pad_w, Cc), dtype=BF16, device=x.device) if want_bf16 else None
    _call("tdr_nchw_to_nhwc", _p(x), B, Cc, H, W, pad_h, pad_w, _p(o32), Cc, _p(o16), Cc, _stream())
    return (o32, o16) if want_bf16 else o32


def nchw_to_nhwc_into(x, pad_h, pad_w, dst32=None, dst16=None):
    """NCHW fp32 -> existing NHWC buffers (any channel width >= C; only the first C channels are written)."""
    x = x.contiguous().float()
    B, Cc, H, W = x.shape
    _call("tdr_nchw_to_nhwc", _p(x), B, Cc, H, W, pad_h, pad_w, _p(dst32), _ld(dst32) if dst32 is not None else 0,
          _p(dst16), _ld(dst16) if dst16 is not None else 0, _stream())


# Opt-in knob (measured, not adopted): the image-boundary 3x3 convs (3 input channels) on the tensor-core path instead of the
# exact fp32 tdr_conv3x3_small_ci.  "1": patch_embed, "2": also the MASA encoder's first conv.  On B200 it saves 0.15-0.4 ms of
# the 47 ms step but rounds the image and the first layer's weights to fp16: guided Restormer 512^2 mean |delta| 1.09e-4 ->
# 1.27e-4 (63.97 -> 63.26 dB), Restormer 256^2 66.6 -> 65.8 dB.  Parity comes first, so the default stays off.
SMALL_CI_TC = os.environ.get("TDR_SMALL_CI_TC", "0") not in ("", "0")
SMALL_CI_TC_MASA = os.environ.get("TDR_SMALL_CI_TC", "0") == "2"


def image_to_rows16(x, pad_h, pad_w, dt, out=None, c16=8):
    """NCHW fp32 image -> zero-padded 16-bit NHWC rows [B, pad_h, pad_w, c16] (tdr_image_to_rows16)."""
    x = x.contiguous().float()
    B, Cc, H, W = x.shape
    if out is None:
        out = torch.empty((B, pad_h, pad_w, c16), dtype=dt, device=x.device)
    assert out.is_contiguous() and tuple(out.shape) == (B, pad_h, pad_w, c16)
    _call("tdr_image_to_rows16", _p(x), B, Cc, H, W, pad_h, pad_w, c16, _p(out), int(out.dtype == F16), _stream())
    return out


def nhwc_to_nchw(x32, out_h, out_w, res=None):
    """NHWC fp32 view -> dense NCHW (cropped); optionally adds an NHWC fp32 residual with the same channel count."""
    B, H, W, Cc = x32.shape
    out = torch.empty((B, Cc, out_h, out_w), dtype=F32, device=x32.device)
    if res is not None:
        assert res.shape == x32.shape and res.dtype == F32
    _call("tdr_nhwc_to_nchw", _p(x32), _ld(x32), B, Cc, H, W, out_h, out_w, _p(res), _ld(res) if res is not None else 0,
          _p(out), _stream())
    return out


def copy_rows(src32, dst32=None, dst16=None):
    B, H, W, Cc = src32.shape
    _call("tdr_copy_rows_f32", _p(src32), _ld(src32), B * H * W, Cc, _p(dst32), _ld(dst32) if dst32 is not None else 0,
          _p(dst16), _ld(dst16) if dst16 is not None else 0, int(dst16 is not None and dst16.dtype == F16), _stream(),
          tag=f"C{Cc}",
          nbytes=B * H * W * Cc * (4 + (4 if dst32 is not None else 0) + (2 if dst16 is not None else 0)))


def conv3x3_small_ci(x32, weight, bias, relu=False, out_f32=None, out_bf16=None):
    """x32: dense fp32 NHWC [B,H,W,Ci<=8]; weight fp32 [Co,Ci,3,3].  out_bf16 may be an fp16 tensor (MASA encoder)."""
    B, H, W, Ci = x32.shape
    assert x32.is_contiguous()
    Co = weight.shape[0]
    flags = int(relu) | (2 if (out_bf16 is not None and out_bf16.dtype == F16) else 0)
    _call("tdr_conv3x3_small_ci", _p(x32), B, H, W, Ci, _p(weight), _p(bias), Co, flags, _p(out_f32),
          _ld(out_f32) if out_f32 is not None else 0, _p(out_bf16), _ld(out_bf16) if out_bf16 is not None else 0,
          _stream(), tag=f"Ci{Ci}_Co{Co}", flops=2 * 9 * B * H * W * Ci * Co,
          nbytes=B * H * W * (Ci * 4 + Co * ((4 if out_f32 is not None else 0) + (2 if out_bf16 is not None else 0))))


def conv3x3_small_co(x16, weight, bias, res32=None):
    B, H, W, Ci = x16.shape
    Co = weight.shape[0]
    out = torch.empty((B, H, W, Co), dtype=F32, device=x16.device)
    if res32 is not None:
        assert res32.is_contiguous() and res32.shape == out.shape
    _call("tdr_conv3x3_small_co", _p(x16), _ld(x16), B, H, W, Ci, _p(weight), _p(bias), Co, _p(res32), _p(out),
          _stream(), tag=f"Ci{Ci}_Co{Co}", flops=2 * 9 * B * H * W * Ci * Co, nbytes=B * H * W * (Ci * 2 + Co * 8))
    return out


# ---------------------------------------------------------------------------------------------- MASA
def sqnorm_rows(x32):
    assert x32.dtype == F32
    B, H, W, Cc = x32.shape
    n2 = torch.empty((B, H, W), dtype=F32, device=x32.device)
    _call("tdr_sqnorm_rows", _p(x32), _ld(x32), B * H * W, Cc, _p(n2), _stream())
    return n2


def masa_split3(x32):
    """fp32 NHWC [B,H,W,C] -> bf16 [B,H,W,3C] = [hi | lo | hi]: the A operand of the split-bf16 MASA correlations."""
    assert x32.dtype == F32
    B, H, W, Cc = x32.shape
    out = torch.empty((B, H, W, 3 * Cc), dtype=BF16, device=x32.device)
    _call("tdr_masa_split3", _p(x32), _ld(x32), B * H * W, Cc, _p(out), 3 * Cc, _stream(), tag=f"C{Cc}",
          nbytes=B * H * W * Cc * 10)
    return out


def cast_rows(x32, want_bf16=True, want_fp16=False, scale16=None, scale_bf16=None, rescale_in=False):
    """fp32 NHWC view -> (bf16 copy, fp16 copy) in one pass (None for the one not requested).  scale16 / scale_bf16:
    1-element DEVICE tensors multiplying the fp16 / bf16 copy; rescale_in: also rewrite x32 *= scale16 in place."""
    assert x32.dtype == F32
    B, H, W, Cc = x32.shape
    o16 = torch.empty((B, H, W, Cc), dtype=BF16, device=x32.device) if want_bf16 else None
    oh = torch.empty((B, H, W, Cc), dtype=F16, device=x32.device) if want_fp16 else None
    _call("tdr_cast_rows", _p(x32), _ld(x32), B * H * W, Cc, _p(o16), Cc, _p(oh), Cc, _p(scale16), _p(scale_bf16),
          int(rescale_in), _stream(), tag=f"C{Cc}",
          nbytes=B * H * W * Cc * (4 + 2 * (int(want_bf16) + int(want_fp16)) + (4 if rescale_in else 0)))
    return o16, oh


def masa_level_scale(x32, state_prev, state_cur):
    """state_cur (zeroed fp32 [4]) <- {s, 1/s, 1/r, max|x|}: the power-of-two level scale of the MASA encoder."""
    B, H, W, Cc = x32.shape
    _call("tdr_masa_level_scale", _p(x32), _ld(x32), B * H * W, Cc, _p(state_prev), _p(state_cur), _stream(),
          tag=f"C{Cc}", nbytes=B * H * W * Cc * 4)


def scale_vec(v, scale, out=None):
    out = torch.empty_like(v) if out is None else out
    lib.call("tdr_scale_vec", _p(v), v.numel(), _p(scale), _p(out), _stream())
    return out


def cvt_f16_bf16(xh):
    B, H, W, Cc = xh.shape
    out = torch.empty((B, H, W, Cc), dtype=BF16, device=xh.device)
    _call("tdr_cvt_f16_bf16", _p(xh), _ld(xh), B * H * W, Cc, _p(out), Cc, _stream(), tag=f"C{Cc}",
          nbytes=B * H * W * Cc * 4)
    return out


def pack_conv_weight_f16(w):
    """nn.Conv2d weight [Co,Ci,kh,kw] -> fp16 [kh*kw, Co, Ci_p8] (tdr_conv_gemm operand of the MASA encoder)."""
    return pack_conv(w, dt=F16)[0]


def _iarr(vals):
    return (C.c_int * len(vals))(*vals)


def masa_ref_invnorm(n2, dils):
    B, H, W = n2.shape
    inv = torch.empty((len(dils), B, H, W), dtype=F32, device=n2.device)
    _call("tdr_masa_ref_invnorm", _p(n2), B, H, W, _iarr(dils), len(dils), _p(inv), _stream())
    return inv


def masa_coarse_filters(f_lq, k_y, k_x, dils, co_pad):
    """f_lq fp32 dense NHWC -> bf16 [ndil, B*9, co_pad, 3C] split filters [hi | hi | lo]."""
    B, H, W, Cc = f_lq.shape
    assert f_lq.is_contiguous() and f_lq.dtype == F32
    w = torch.empty((len(dils), B * 9, co_pad, 3 * Cc), dtype=BF16, device=f_lq.device)
    _call("tdr_masa_coarse_filters", _p(f_lq), B, H, W, Cc, k_y, k_x, _iarr(dils), len(dils), co_pad, _p(w), _stream())
    return w


def masa_coarse_argmax(score, nblk, d_y, d_x):
    B, Hr, Wr, co_pad = score.shape
    idx = torch.empty((B, nblk), dtype=torch.int32, device=score.device)
    origin = torch.empty((B * nblk, 3), dtype=torch.int32, device=score.device)
    _call("tdr_masa_coarse_argmax", _p(score), B, Hr, Wr, nblk, co_pad, d_y, d_x, _p(idx), _p(origin), _stream())
    return idx, origin


def masa_fine_filters(f_lq, k_y, k_x):
    B, H, W, Cc = f_lq.shape
    assert f_lq.is_contiguous() and f_lq.dtype == F32
    nwin = B * (H // k_y) * (W // k_x)
    w = torch.empty((nwin * 9, k_y * k_x, 3 * Cc), dtype=BF16, device=f_lq.device)
    _call("tdr_masa_fine_filters", _p(f_lq), B, H, W, Cc, k_y, k_x, _p(w), _stream())
    return w


def masa_win_invnorm(n2_ref, origin, d_y, d_x):
    B, Hr, Wr = n2_ref.shape
    nwin = origin.shape[0]
    inv = torch.empty((nwin, d_y, d_x), dtype=F32, device=n2_ref.device)
    _call("tdr_masa_win_invnorm", _p(n2_ref), Hr, Wr, _p(origin), nwin, d_y, d_x, _p(inv), _stream())
    return inv


def masa_fine_argmax(corr):
    nwin, dy, dx, nq = corr.shape
    index = torch.empty((nwin, nq), dtype=torch.int32, device=corr.device)
    att = torch.empty((nwin, nq), dtype=F32, device=corr.device)
    _call("tdr_masa_fine_argmax", _p(corr), nwin, dy * dx, nq, _p(index), _p(att), _stream())
    return index, att


def masa_transfer(f_ref, origin, index, att, py, px, k_y, k_x, d_x, s, out32=None, out16=None):
    B, Hs, Ws, Cc = f_ref.shape
    assert f_ref.is_contiguous()
    npx = B * py * k_y * s * px * k_x * s
    _call("tdr_masa_transfer", _p(f_ref), B, Hs, Ws, Cc, _p(origin), _p(index), _p(att), py, px, k_y, k_x, d_x, s,
          _p(out32), _ld(out32) if out32 is not None else 0, _p(out16), _ld(out16) if out16 is not None else 0,
          _stream(), tag=f"s{s}_C{Cc}",
          nbytes=npx * Cc * (2 + (4 if out32 is not None else 0) + (2 if out16 is not None else 0)))


# ---------------------------------------------------------------------------------------------- ViT glue
def vit_patchify(img, patch):
    img = img.contiguous().float()
    B, Cc, H, W = img.shape
    K = Cc * patch * patch
    n = (H // patch) * (W // patch)
    out = torch.zeros((B, 1, n, round_up(K, 8)), dtype=BF16, device=img.device)
    _call("tdr_vit_patchify", _p(img), B, Cc, H, W, patch, _p(out), out.shape[3], _stream())
    return out


def vit_assemble_tokens(patch_tokens, cls, pos):
    B, _, N, D = patch_tokens.shape
    x = torch.empty((B, 1, N + 1, D), dtype=F32, device=patch_tokens.device)
    _call("tdr_vit_assemble_tokens", _p(patch_tokens), _p(cls), _p(pos), B, N, D, _p(x), _stream())
    return x


def softmax_rows(s32, n, scale, out16):
    rows = s32.numel() // s32.shape[-1]
    _call("tdr_softmax_rows", _p(s32), s32.shape[-1], rows, n, scale, _p(out16), out16.shape[-1], _stream(),
          nbytes=rows * n * 6)
    return out16


def vit_transpose_v(qkv, heads, hd, voff, n_pad):
    B, _, N, _ = qkv.shape
    vt = torch.empty((B, heads, hd, n_pad), dtype=BF16, device=qkv.device)
    _call("tdr_vit_transpose_v", _p(qkv), _ld(qkv), B, N, heads, hd, voff, _p(vt), n_pad, _stream())
    return vt


def grouped_stencil(x16, idx, weight, bias, K, out16, dil=1, relu=False, pool=False):
    """Generic grouped / depthwise stencil (tdr_grouped_stencil): x16 NHWC 16-bit view; idx int32 [Co * ipg] input channel
    table; weight fp32 [Co, ipg, K, K]; out16: NHWC 16-bit view whose channel count is Co (a slice of a wider buffer)."""
    B, H, W, _ = x16.shape
    Co = out16.shape[-1]
    ipg = idx.numel() // Co
    assert ipg * Co == idx.numel() and out16.dtype == x16.dtype
    _call("tdr_grouped_stencil", _p(x16), _ld(x16), B, H, W, Co, ipg, _p(idx), _p(weight), _p(bias), K, dil, int(relu),
          int(pool), _p(out16), _ld(out16), int(x16.dtype == F16), _stream(), tag=f"k{K}d{dil}g{ipg}_C{Co}",
          nbytes=B * H * W * (Co * ipg + Co) * 2)
    return out16


def mefc_gate(emb, w1, b1, w2, b2, num_ops):
    """softmax over ops of Linear(ReLU(Linear(emb))): emb fp32 [B, C] -> fp32 [B, steps, num_ops] (tdr_mefc_gate)."""
    B, C_ = emb.shape
    H1, O = w1.shape[0], w2.shape[0]
    out = torch.empty((B, O // num_ops, num_ops), dtype=F32, device=emb.device)
    _call("tdr_mefc_gate", _p(emb), emb.stride(0), B, C_, _p(w1), _p(b1), H1, _p(w2), _p(b2), O, num_ops, _p(out), _stream())
    return out


def mefc_mix_weights(w32, gate, C_, dt):
    """w32 fp32 [Co, n*C]; gate fp32 [B, n] view (row stride gate.stride(0)) -> 16-bit per-sample weights [B, Co, ld]."""
    Co, Ci = w32.shape
    B = gate.shape[0]
    ld = round_up(Ci, 8)
    out = torch.empty((B, Co, ld), dtype=dt, device=w32.device)
    _call("tdr_mefc_mix_weights", _p(w32), Co, Ci, C_, _p(gate), gate.stride(0), B, _p(out), ld, int(dt == F16), _stream())
    return out


def prompt_weights(emb, weight, bias):
    """softmax(linear_layer(emb), dim=1): emb fp32 [B, C] (row stride emb.stride(0)), weight fp32 [L, C] -> fp32 [B, L]."""
    B, C_ = emb.shape
    L = weight.shape[0]
    out = torch.empty((B, L), dtype=F32, device=emb.device)
    _call("tdr_prompt_weights", _p(emb), emb.stride(0), _p(weight), _p(bias), B, C_, L, _p(out), _stream())
    return out


def prompt_mix_resize(prompt, wts, H, W, out16):
    """prompt fp32 [L, D, S, S] (prompt_param), wts fp32 [B, L] -> out16 (NHWC 16-bit view [B,H,W,D]) =
    bilinear resize of the weighted prompt sum (tdr_prompt_mix_resize)."""
    L, D, S, _ = prompt.shape
    B = wts.shape[0]
    _call("tdr_prompt_mix_resize", _p(prompt), L, D, S, _p(wts), B, H, W, _p(out16), _ld(out16),
          1 if out16.dtype == F16 else 0, _stream(), nbytes=B * H * W * D * 2)
    return out16


def vit_attention_supported(hd):
    return bool(lib.load().tdr_vit_attention_supported(int(hd)))


def vit_attention(qkv, heads, hd, scale, out=None):
    """Fused softmax(scale q k^T) v over bf16 qkv rows [B,1,N,3D] -> bf16 [B,1,N,D] (tdr_vit_attention, one launch)."""
    B, _, N, _ = qkv.shape
    D = heads * hd
    if out is None:
        out = torch.empty((B, 1, N, D), dtype=BF16, device=qkv.device)
    _call("tdr_vit_attention", _p(qkv), _ld(qkv), B, N, heads, hd, float(scale), _p(out), _ld(out), _stream(),
          nbytes=B * N * D * 2 * 4)
    return out


def mean_tokens(x32, t0, n, out, accumulate=False):
    """x32 fp32 [B,1,T,C]; out fp32 [B, C] view (row stride out.stride(0))."""
    B, _, T, Cc = x32.shape
    _call("tdr_mean_tokens", _p(x32), _ld(x32), B, T, t0, n, Cc, _p(out), out.stride(0), int(accumulate), _stream())
    return out


def crop_resize(img, origin, crop_hw, out_hw):
    """img NCHW fp32; origin int32 [n, 3] = (image, y0, x0) on device -> [n, C, oh, ow] fp32 (bilinear)."""
    img = img.contiguous().float()
    _, Cc, H, W = img.shape
    n = origin.shape[0]
    out = torch.empty((n, Cc, out_hw[0], out_hw[1]), dtype=F32, device=img.device)
    _call("tdr_crop_resize", _p(img), Cc, H, W, _p(origin), n, crop_hw[0], crop_hw[1], out_hw[0], out_hw[1], _p(out),
          _stream())
    return out


def cosine_rows(fl, fr, n):
    """fl fp32 [B, F], fr fp32 [B*n, F] (both dense) -> cos fp32 [B, n]."""
    assert fl.is_contiguous() and fr.is_contiguous() and fl.dtype == F32 and fr.dtype == F32
    B, Fdim = fl.shape[0], fl.numel() // fl.shape[0]
    out = torch.empty((B, n), dtype=F32, device=fl.device)
    _call("tdr_cosine_rows", _p(fl), _p(fr), B, n, Fdim, _p(out), _stream())
    return out


# ---------------------------------------------------------------------------------------------- backward (training)
_WS = {}


def workspace(nbytes, dev):
    """Grow-only fp32 scratch buffer per device (stream-ordered reuse: every consumer runs on the current stream)."""
    t = _WS.get(dev)
    n = (int(nbytes) + 3) // 4
    if t is None or t.numel() < n:
        t = torch.empty(max(n, 1 << 20), dtype=F32, device=dev)
        _WS[dev] = t
    return t


def wgrad(dy16, x16, out, *, Co=None, Ci=None, k=1, stride=1, pad=0, dil=1, per_sample=False, strides=None,
          co_map=None, ci_map=None, accumulate=True, scale=1.0, scale_ptr=None):
    """Weight gradient of a dense conv.  dy16 bf16 NHWC [B,OH,OW,>=Co], x16 bf16 NHWC [B,H,W,>=Ci]; ``out`` fp32 in
    the PARAMETER layout [Co, Ci, k, k] (default strides) or any layout given by strides=(s_b, s_co, s_ci, s_tap)."""
    assert dy16.dtype == BF16 and out.dtype == F32
    if x16.dtype == F16:            # MASA encoder tape: the GEMM takes bf16 pairs only (kind::f16 rejects mixed operands)
        x16 = cvt_f16_bf16(x16)
    assert x16.dtype == BF16
    B, H, W, Cx = x16.shape
    Co = dy16.shape[3] if Co is None else Co
    Ci = Cx if Ci is None else Ci
    d = lib.WgradDesc()
    d.dy = dy16.data_ptr(); d.dy_ld = _ld(dy16)
    d.x = x16.data_ptr(); d.x_ld = _ld(x16)
    d.B = B; d.H = H; d.W = W; d.Ci = Ci; d.Co = Co; d.KH = k; d.KW = k; d.stride = stride; d.pad = pad; d.dil = dil
    d.per_sample = int(per_sample)
    OH = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    OW = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    assert dy16.shape[0] == B and dy16.shape[1] == OH and dy16.shape[2] == OW, \
        f"wgrad: dy {tuple(dy16.shape)} does not match conv output {(B, OH, OW)}"
    if strides is None:
        ci_l = out.shape[1]
        strides = (0, ci_l * k * k, k * k, 1)
    d.out = out.data_ptr()
    d.out_stride_b, d.out_stride_co, d.out_stride_ci, d.out_stride_tap = strides
    d.co_map = co_map.data_ptr() if co_map is not None else None
    d.ci_map = ci_map.data_ptr() if ci_map is not None else None
    d.accumulate = int(accumulate); d.scale = scale
    d.scale_ptr = scale_ptr.data_ptr() if scale_ptr is not None else None
    nbytes = lib.load().tdr_wgrad_workspace_bytes(C.byref(d))
    if nbytes == 0:
        raise lib.TdrError(f"tdr_wgrad: unsupported shape Co={Co} Ci={Ci} k={k}")
    ws = workspace(nbytes, x16.device)
    d.workspace = ws.data_ptr(); d.workspace_bytes = ws.numel() * 4
    _call("tdr_wgrad", C.byref(d), _stream(), tag=f"k{k}s{stride}_Ci{Ci}_Co{Co}_{OH}x{OW}" + ("_ps" if per_sample else ""),
          nbytes=B * H * W * Ci * 2 * (1 if k == 1 else 1) + B * OH * OW * Co * 2, flops=2 * B * OH * OW * Co * Ci * k * k)
    return out


def _red_ws(Cc, dev):
    return workspace(lib.load().tdr_reduce_workspace_bytes(Cc), dev)


def colsum(x16, out, c_map=None, accumulate=True, Cc=None):
    B, H, W, Cx = x16.shape
    Cc = Cx if Cc is None else Cc
    _call("tdr_colsum", _p(x16), _ld(x16), B * H * W, Cc, _p(out), 1, _p(c_map), int(accumulate), _p(_red_ws(Cc, x16.device)),
          _stream(), tag=f"C{Cc}", nbytes=B * H * W * Cc * 2)
    return out


def dwconv3x3_wgrad(dy16, x16, dw, db=None, c_map=None, accumulate=True):
    B, H, W, Cc = x16.shape
    assert dy16.shape == x16.shape
    _call("tdr_dwconv3x3_wgrad", _p(dy16), _ld(dy16), _p(x16), _ld(x16), B, H, W, Cc, _p(dw), _p(db), _p(c_map),
          int(accumulate), _p(_red_ws(Cc, x16.device)), _stream(), tag=f"C{Cc}_{H}x{W}", nbytes=B * H * W * Cc * 4,
          flops=2 * 9 * B * H * W * Cc)


def rownorm_bwd(x32, dy16, mode, weight=None, eps=1e-5, add=None, out=None, dweight=None, dbias=None, accumulate=True,
                want16=False):
    """dx = add + LN_bwd(dy) (fp32).  mode 0: dx = add + float(dy).  out may alias add.  want16: also return a bf16
    copy of dx (what the next dgrad / wgrad GEMMs consume) -> (dx, dx16)."""
    B, H, W, Cc = dy16.shape
    if out is None:
        out = torch.empty((B, H, W, Cc), dtype=F32, device=dy16.device)
    out16 = torch.empty((B, H, W, Cc), dtype=BF16, device=dy16.device) if want16 else None
    _call("tdr_rownorm_bwd", _p(x32), _ld(x32) if x32 is not None else 0, _p(dy16), _ld(dy16), B * H * W, Cc, mode,
          _p(weight), eps, _p(add), _ld(add) if add is not None else 0, _p(out), _ld(out), _p(out16),
          _ld(out16) if out16 is not None else 0, _p(dweight), _p(dbias), int(accumulate),
          _p(_red_ws(Cc, dy16.device)) if dweight is not None else None, _stream(), tag=f"m{mode}_C{Cc}_{H}x{W}",
          nbytes=B * H * W * Cc * (2 + 4 + (4 if mode else 0) + (4 if add is not None else 0) + (2 if want16 else 0)))
    return (out, out16) if want16 else out


def gate_bwd(y16, dg16, gate, out=None, dg_add=None):
    """dg_add: optional fp32 [B, C2/2] added to dg of every pixel of sample b (SCA pool gradient)."""
    B, H, W, C2 = y16.shape
    if out is None:
        out = y16
    _call("tdr_gate_bwd", _p(y16), _ld(y16), _p(dg16), _ld(dg16), B * H * W, C2 // 2, gate, _p(out), _ld(out), _p(dg_add),
          H * W, _stream(), tag=f"g{gate}_C{C2}", nbytes=B * H * W * C2 * 5)
    return out


def dwconv3x3_gated_train(x16, w9, bias, gate):
    """Gated depthwise conv that also returns the pre-gate tensor: (g [B,H,W,C/2], y [B,H,W,C])."""
    B, H, W, Cc = x16.shape
    out = torch.empty((B, H, W, Cc // 2), dtype=BF16, device=x16.device)
    y = torch.empty((B, H, W, Cc), dtype=BF16, device=x16.device)
    _call("tdr_dwconv3x3_gated_train", _p(x16), _ld(x16), B, H, W, Cc, _p(w9), _p(bias), gate, _p(out), _ld(out), _p(y),
          _ld(y), _stream(), tag=f"g{gate}_C{Cc}_{H}x{W}", nbytes=B * H * W * Cc * 5, flops=2 * 9 * B * H * W * Cc)
    return out, y


def dwconv3x3_gate_bwd(x16, w9, bias, gate, dg16, dg_add=None):
    """Fused recompute of the gated depthwise conv + gate backward: returns d(pre-gate) bf16 [B,H,W,C]."""
    B, H, W, Cc = x16.shape
    out = torch.empty((B, H, W, Cc), dtype=BF16, device=x16.device)
    _call("tdr_dwconv3x3_gate_bwd", _p(x16), _ld(x16), B, H, W, Cc, _p(w9), _p(bias), gate, _p(dg16), _ld(dg16),
          _p(dg_add), _p(out), _ld(out), _stream(), tag=f"g{gate}_C{Cc}_{H}x{W}", nbytes=B * H * W * Cc * 5,
          flops=2 * 9 * B * H * W * Cc)
    return out


def naf_scaled_conv_bwd(raw, W, bias, scale, cs, s, dW, dbias, dscale):
    """raw fp32 [nb, Co, C]; accumulates dW [Co, C], dbias [Co], dscale [Co] (see include/tdr_sm100.h)."""
    nb, Co, Cc = raw.shape
    _call("tdr_naf_scaled_conv_bwd", _p(raw), nb, Co, Cc, _p(W), _p(bias), _p(scale), _p(cs), _p(s), _p(dW), _p(dbias),
          _p(dscale), _stream(), tag=f"Co{Co}_C{Cc}")


def naf_sca_bwd(raw, w3, scale, mean, w_sca, P, dw_sca, db_sca):
    """Returns dg_add fp32 [B, C] (to be passed to gate_bwd); accumulates dw_sca [C, C] and db_sca [C]."""
    B, Co, Cc = raw.shape
    dg_add = torch.empty((B, Cc), dtype=F32, device=raw.device)
    ws = torch.empty(B * Cc, dtype=F32, device=raw.device)
    _call("tdr_naf_sca_bwd", _p(raw), B, Co, Cc, _p(w3), _p(scale), _p(mean), _p(w_sca), P, _p(dw_sca), _p(db_sca),
          _p(dg_add), _p(ws), _stream(), tag=f"C{Cc}")
    return dg_add


def mdta_bwd(saved, B, P, C_, heads, temperature, w_out, dweff, dw_out, dtemp):
    """Returns mqk bf16 [B, 2C, 2C_p]: [dq; dk] = mqk[b] . [q; k].  Accumulates dw_out [C,C] and dtemp [heads]."""
    cp2 = round_up(2 * C_, 8)
    mqk = torch.zeros((B, 2 * C_, cp2), dtype=BF16, device=dweff.device)
    ws = workspace(lib.load().tdr_mdta_bwd_workspace_bytes(B, C_, heads), dweff.device)
    _call("tdr_mdta_bwd", _p(saved["shat"]), _p(saved["attn"]), B, P, C_, heads, _p(temperature), _p(w_out), _p(dweff),
          _p(mqk), cp2, _p(dw_out), _p(dtemp), 1, _p(ws), _stream(), tag=f"C{C_}_h{heads}")
    return mqk


def scale_add(x32, y32=None, scale_ptr=None, scale=1.0, out=None):
    B, H, W, Cc = x32.shape
    if out is None:
        out = torch.empty((B, H, W, Cc), dtype=F32, device=x32.device)
    _call("tdr_scale_add_f32", _p(x32), _ld(x32), _p(y32), _ld(y32) if y32 is not None else 0, B * H * W, Cc,
          _p(scale_ptr), scale, _p(out), _ld(out), _stream(), tag=f"C{Cc}",
          nbytes=B * H * W * Cc * (8 + (4 if y32 is not None else 0)))
    return out


def dot_f32(x32, y32, out, accumulate=True):
    B, H, W, Cc = x32.shape
    _call("tdr_dot_f32", _p(x32), _ld(x32), _p(y32), _ld(y32), B * H * W, Cc, _p(out), int(accumulate),
          _p(workspace(1 << 16, x32.device)), _stream(), tag=f"C{Cc}", nbytes=B * H * W * Cc * 8)
    return out


def pixel_shuffle(x16, mode):
    """mode 1: PixelUnshuffle(2) [B,H,W,C] -> [B,H/2,W/2,4C]; mode 2: PixelShuffle(2) [B,H,W,C] -> [B,2H,2W,C/4]."""
    B, H, W, Cc = x16.shape
    oshape = (B, H // 2, W // 2, Cc * 4) if mode == 1 else (B, H * 2, W * 2, Cc // 4)
    if oshape[3] % 8:            # keep 16 B rows for TMA: zero-padded channel stride, the view has the logical width
        full = torch.zeros(oshape[:3] + (round_up(oshape[3], 8),), dtype=BF16, device=x16.device)
        out = full[..., :oshape[3]]
    else:
        out = torch.empty(oshape, dtype=BF16, device=x16.device)
    _call("tdr_pixel_shuffle_nhwc", _p(x16), _ld(x16), B, H, W, Cc, mode, _p(out), _ld(out), _stream(),
          tag=f"m{mode}_C{Cc}", nbytes=B * H * W * Cc * 4)
    return out


def relu_mask(y16, dy16, out=None):
    B, H, W, Cc = y16.shape
    if out is None:
        out = torch.empty((B, H, W, Cc), dtype=BF16, device=y16.device)
    _call("tdr_relu_mask", _p(y16), _ld(y16), _p(dy16), _ld(dy16), B * H * W, Cc, _p(out), _ld(out), _stream(),
          tag=f"C{Cc}", nbytes=B * H * W * Cc * 6)
    return out


def masa_transfer_bwd(dout32, f_ref, origin, index, att, py, px, k_y, k_x, d_x, s, dref32, datt):
    B, Hs, Ws, Cc = f_ref.shape
    assert f_ref.is_contiguous() and dref32.is_contiguous() and dref32.shape == f_ref.shape and dref32.dtype == F32
    assert tuple(dout32.shape) == (B, py * k_y * s, px * k_x * s, Cc) and dout32.dtype == F32
    _call("tdr_masa_transfer_bwd", _p(dout32), _ld(dout32), _p(f_ref), B, Hs, Ws, Cc, _p(origin), _p(index), _p(att),
          py, px, k_y, k_x, d_x, s, _p(dref32), _p(datt), _stream(), tag=f"s{s}_C{Cc}",
          nbytes=dout32.shape[0] * dout32.shape[1] * dout32.shape[2] * Cc * (4 + 2 + 8))


def masa_fine_bwd(f_lq, f_ref, origin, index, datt, k_y, k_x, d_x, dlq32, dref32):
    B, H, W, Cc = f_lq.shape
    _, Hr, Wr, _ = f_ref.shape
    assert f_lq.is_contiguous() and f_ref.is_contiguous() and dlq32.is_contiguous() and dref32.is_contiguous()
    _call("tdr_masa_fine_bwd", _p(f_lq), B, H, W, _p(f_ref), Hr, Wr, Cc, k_y, k_x, d_x, _p(origin), _p(index), _p(datt),
          _p(dlq32), _p(dref32), _stream(), tag=f"C{Cc}")


def dilate2(x16, OH, OW):
    """Zero insertion: [B,H,W,C] -> [B,OH,OW,C] with out[:, 2y, 2x] = x[:, y, x] (adjoint of a stride-2 conv's subsampling)."""
    B, H, W, Cc = x16.shape
    out = torch.zeros((B, OH, OW, Cc), dtype=BF16, device=x16.device)
    _call("tdr_dilate2_nhwc", _p(x16), _ld(x16), B, H, W, Cc, _p(out), _ld(out), OH, OW, _stream(), tag=f"C{Cc}",
          nbytes=B * H * W * Cc * 4)
    return out


# ------------------------------------------------------------------------------------------------ input pipeline
def prepare_patches(frames, crops, size, bgr2rgb=True, mean=None, std=None, out=None):
    """Device-side ``Dataset_PairedImageWithRef.__getitem__`` tensor preparation (data/restoration_dataset.py:194-253).

    frames: list of uint8 CUDA tensors [h, w, c] (decoded BGR frames, HWC); crops: list of (top, left, mode) -- the
    reference's ``random.randint`` decisions for paired_random_crop / random_augmentation; size: int or (out_h, out_w).
    Returns fp32 [n, c, out_h, out_w] in [0, 1] (or normalised with mean / std), RGB, bit-identical to the reference's
    numpy / cv2 chain (reflect padding of frames smaller than the patch included)."""
    n = len(frames)
    assert n > 0 and len(crops) == n
    out_h, out_w = (size, size) if isinstance(size, int) else size
    ch = frames[0].shape[2]
    descs = (lib.PatchDesc * n)()
    for i, (f, (top, left, mode)) in enumerate(zip(frames, crops)):
        assert f.dtype == torch.uint8 and f.is_cuda and f.dim() == 3 and f.is_contiguous() and f.shape[2] == ch
        descs[i].image = f.data_ptr(); descs[i].h = f.shape[0]; descs[i].w = f.shape[1]
        descs[i].top = top; descs[i].left = left; descs[i].mode = mode
    raw = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).pin_memory()
    d_dev = raw.to(frames[0].device, non_blocking=True)
    if out is None:
        out = torch.empty((n, ch, out_h, out_w), dtype=F32, device=frames[0].device)
    assert out.dtype == F32 and out.is_contiguous() and tuple(out.shape) == (n, ch, out_h, out_w)
    m = (C.c_float * ch)(*mean) if mean is not None else None
    sd = (C.c_float * ch)(*std) if std is not None else None
    _call("tdr_prepare_patches", d_dev.data_ptr(), C.cast(descs, C.c_void_p), n, ch, out_h, out_w, int(bgr2rgb),
          C.cast(m, C.c_void_p) if m is not None else None, C.cast(sd, C.c_void_p) if sd is not None else None,
          out.data_ptr(), _stream(), tag=f"n{n}_{out_h}x{out_w}", nbytes=out.numel() * 4 + n * out_h * out_w * ch)
    return out


def add_gaussian_noise(img, sigma, noise=None, seed=0, out=None):
    """Training noise of the Gaussian-denoising datasets (data/restoration_dataset.py:474-476) on the device.
    img: fp32 CUDA [B, ...]; sigma: per-sample noise levels in 8-bit units (float, sequence or tensor) -> img + n * sigma/255.
    noise: standard normals with img's shape (then the result equals the reference's for that draw, bit for bit);
    None: drawn on the device from a Philox stream keyed by ``seed`` (reproducible, no host RNG)."""
    assert img.dtype == F32 and img.is_cuda and img.is_contiguous()
    B = img.shape[0]
    lv = torch.as_tensor(sigma, dtype=F32).reshape(-1)
    if lv.numel() == 1:
        lv = lv.expand(B)
    assert lv.numel() == B
    lv = (lv / 255.0).contiguous().to(img.device)       # torch.FloatTensor([sigma]) / 255.0, as the reference computes it
    if noise is not None:
        assert noise.shape == img.shape and noise.dtype == F32 and noise.is_cuda and noise.is_contiguous()
    if out is None:
        out = torch.empty_like(img)
    _call("tdr_add_gaussian_noise", _p(img), _p(noise), _p(lv), B, img.numel() // B, int(seed), _p(out), _stream(),
          nbytes=img.numel() * (12 if noise is not None else 8))
    return out


def psnr_u8(result, gt, crop_border=0):
    """Validation PSNR of ``calculate_psnr(tensor2img(result), tensor2img(gt), crop_border)`` (metrics/psnr_ssim.py:9-63,
    utils/utils_image.py:129-191; ``test_y_channel=False``) per image of fp32 NCHW CUDA batches: the uint8 quantisation
    and the integer sums run on the device, 12 bytes per image come back, and the float64 tail is the reference's own
    formula -- the returned doubles are identical to the reference's.  Returns a list of floats (``inf`` for equal images)."""
    import numpy as np
    assert result.dtype == F32 and gt.dtype == F32 and result.shape == gt.shape and result.dim() == 4
    result, gt = result.contiguous(), gt.contiguous()
    B, Cc, H, W = result.shape
    sse = torch.empty(B, dtype=torch.int64, device=result.device)
    mx = torch.empty(B, dtype=torch.int32, device=result.device)
    _call("tdr_psnr_u8_sums", result.data_ptr(), gt.data_ptr(), B, Cc, H, W, crop_border, sse.data_ptr(), mx.data_ptr(),
          _stream(), tag=f"{Cc}x{H}x{W}", nbytes=2 * result.numel() * 4)
    n = Cc * (H - 2 * crop_border) * (W - 2 * crop_border)
    out = []
    for s_, m_ in zip(sse.tolist(), mx.tolist()):
        mse = np.float64(s_) / np.float64(n)                       # np.mean of exact integers in float64
        if mse == 0:
            out.append(float("inf"))
            continue
        max_value = 1. if m_ <= 1 else 255.
        out.append(float(20. * np.log10(max_value / np.sqrt(mse))))
    return out
