"""B200-native ``DRSformerRefFusion`` and ``DRSformer200L_SPA_RefFusion`` (inference).

Drop-ins for the classes of the same names in the reference's ``models/archs/network_drsformer_guided_arch.py``
(:679-1123; options 008-010) and ``network_drsformer_guided_arch_200L_SPA.py`` (option 007): same constructor kwargs,
``state_dict`` keys / order / shapes, ``net(lq, ref)`` on NCHW tensors.  The MASA guidance, the U-Net wiring and the
Res-fusion scheme are the guided Restormer's (the schedule is shared through ``RestormerRefFusion._guided_encode``); what
differs is inside the blocks:

  * **TKSA** (:260-332): channel attention whose softmax is a mix of four top-k masked softmaxes.  The mix is still one
    c x c matrix per (sample, head), so it folds into ``project_out`` exactly like MDTA: ``tdr_mdta_weff(topk_w=...)``.
  * **MSFN** (:216-256): depthwise 3x3 and 5x5 (+ReLU) on the hidden tensor, then two grouped convs (2 input channels per
    output channel) over the re-interleaved halves: ``tdr_grouped_stencil`` with channel index tables, no chunk / cat copies.
  * **MEFC** ``subnet`` (:371-549, full model only: ``encoder_level0`` and ``refinement``): four steps of eight experts
    (separable / dilated depthwise convs, average pooling) gated per sample; the gate is folded into the per-sample
    weights of each step's 1x1 reduction (``tdr_mefc_mix_weights``), the experts' 1x1 convs are ``tdr_conv_gemm``.

Inference only: the explicit backward is not implemented for this family (``TdrError`` under grad mode).
"""
import torch
import torch.nn as nn

from .. import ops
from ..lib import TdrError
from .masa import Encoder, MasaMixin, prep_conv as _prep_conv, _f
from .restormer_b200_arch import (Downsample, LayerNorm, OverlapPatchEmbed, RestormerRefFusion, Upsample, _blocks,
                                  operand_dtype)

F32, I32 = torch.float32, torch.int32


# ----------------------------------------------------------------------------------------------- parameter holders
class FeedForward(nn.Module):
    """:216-256 MSFN."""

    def __init__(self, dim, ffn_expansion_factor, bias):
        super().__init__()
        h = int(dim * ffn_expansion_factor)
        self.project_in = nn.Conv2d(dim, h * 2, 1, bias=bias)
        self.dwconv3x3 = nn.Conv2d(h * 2, h * 2, 3, 1, 1, groups=h * 2, bias=bias)
        self.dwconv5x5 = nn.Conv2d(h * 2, h * 2, 5, 1, 2, groups=h * 2, bias=bias)
        self.relu3, self.relu5 = nn.ReLU(), nn.ReLU()
        self.dwconv3x3_1 = nn.Conv2d(h * 2, h, 3, 1, 1, groups=h, bias=bias)
        self.dwconv5x5_1 = nn.Conv2d(h * 2, h, 5, 1, 2, groups=h, bias=bias)
        self.relu3_1, self.relu5_1 = nn.ReLU(), nn.ReLU()
        self.project_out = nn.Conv2d(h * 2, dim, 1, bias=bias)


class Attention(nn.Module):
    """:260-332 TKSA."""

    def __init__(self, dim, num_heads, bias):
        super().__init__()
        self.num_heads = num_heads
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.qkv = nn.Conv2d(dim, dim * 3, 1, bias=bias)
        self.qkv_dwconv = nn.Conv2d(dim * 3, dim * 3, 3, 1, 1, groups=dim * 3, bias=bias)
        self.project_out = nn.Conv2d(dim, dim, 1, bias=bias)
        self.attn_drop = nn.Dropout(0.)
        self.attn1 = nn.Parameter(torch.tensor([0.2]))
        self.attn2 = nn.Parameter(torch.tensor([0.2]))
        self.attn3 = nn.Parameter(torch.tensor([0.2]))
        self.attn4 = nn.Parameter(torch.tensor([0.2]))


class TransformerBlock(nn.Module):
    def __init__(self, dim, num_heads, ffn_expansion_factor, bias, LayerNorm_type):
        super().__init__()
        self.norm1 = LayerNorm(dim, LayerNorm_type)
        self.attn = Attention(dim, num_heads, bias)
        self.norm2 = LayerNorm(dim, LayerNorm_type)
        self.ffn = FeedForward(dim, ffn_expansion_factor, bias)


class TransformerResFusionBlock(TransformerBlock):
    def __init__(self, dim, num_heads, ffn_expansion_factor, bias, LayerNorm_type):
        super().__init__(dim, num_heads, ffn_expansion_factor, bias, LayerNorm_type)
        self.alpha = nn.Parameter(torch.zeros(1))


def _dw(C, k, pad, dil=1):
    return nn.Conv2d(C, C, k, 1, pad, dilation=dil, groups=C, bias=False)


class SepConv(nn.Module):
    def __init__(self, C, k):
        super().__init__()
        self.op = nn.Sequential(_dw(C, k, k // 2), nn.Conv2d(C, C, 1, bias=False), nn.ReLU(), _dw(C, k, k // 2),
                                nn.Conv2d(C, C, 1, bias=False))


class DilConv(nn.Module):
    def __init__(self, C, k):
        super().__init__()
        self.op = nn.Sequential(_dw(C, k, k - 1, 2), nn.Conv2d(C, C, 1, bias=False))


class OperationLayer(nn.Module):
    def __init__(self, C):
        super().__init__()
        self._ops = nn.ModuleList([SepConv(C, 1), SepConv(C, 3), SepConv(C, 5), SepConv(C, 7), DilConv(C, 3), DilConv(C, 5),
                                   DilConv(C, 7), nn.AvgPool2d(3, stride=1, padding=1, count_include_pad=False)])
        self._out = nn.Sequential(nn.Conv2d(C * 8, C, 1, bias=False), nn.ReLU())


class ReLUConv(nn.Module):
    def __init__(self, C):
        super().__init__()
        self.op = nn.Sequential(nn.Conv2d(C, C, 1, bias=False), nn.ReLU())


class GroupOLs(nn.Module):
    def __init__(self, steps, C):
        super().__init__()
        self.preprocess = ReLUConv(C)
        self._steps = steps
        self._ops = nn.ModuleList([OperationLayer(C) for _ in range(steps)])
        self.relu = nn.ReLU()


class OALayer(nn.Module):
    def __init__(self, channel, k, num_ops):
        super().__init__()
        self.k, self.num_ops, self.output = k, num_ops, k * num_ops
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.ca_fc = nn.Sequential(nn.Linear(channel, self.output * 2), nn.ReLU(), nn.Linear(self.output * 2, self.output))


class subnet(nn.Module):
    """:522-549 MEFC."""

    def __init__(self, dim, layer_num=1, steps=4):
        super().__init__()
        self._C, self.num_ops, self._layer_num, self._steps = dim, 8, layer_num, steps
        self.layers = nn.ModuleList()
        for _ in range(layer_num):
            self.layers += [OALayer(dim, steps, 8)]
            self.layers += [GroupOLs(steps, dim)]


# ----------------------------------------------------------------------------------------------- weight preparation
def _arange(n, dev, pad_to=None):
    t = torch.full((pad_to or n,), -1, dtype=I32, device=dev)
    t[:n] = torch.arange(n, dtype=I32, device=dev)
    return t


def _prep_block(blk):
    a, f = blk.attn, blk.ffn
    C_ = a.qkv.in_channels
    h = f.dwconv3x3_1.out_channels
    P2 = ops.round_up(2 * h, 8)
    dev = a.qkv.weight.device
    dt = operand_dtype(False)
    p = dict(C=C_, heads=a.num_heads, h=h, P2=P2, dt=dt)
    p["ln1_w"], p["ln1_b"] = _f(blk.norm1.body.weight), _f(blk.norm1.body.bias)
    p["ln2_w"], p["ln2_b"] = _f(blk.norm2.body.weight), _f(blk.norm2.body.bias)
    p["ln_mode"] = 1 if blk.norm1.body.bias is not None else 2
    p["w_qkv"], _ = ops.pack_conv(a.qkv.weight, dt=dt)
    p["b_qkv"] = _f(a.qkv.bias)
    p["w_qkv_dw"], _, p["b_qkv_dw"] = ops.pack_dw(a.qkv_dwconv.weight, a.qkv_dwconv.bias)
    p["temp"] = _f(a.temperature).reshape(-1)
    p["w_po"], p["b_po"] = _f(a.project_out.weight).reshape(C_, C_), _f(a.project_out.bias)
    p["topk_w"] = torch.cat([_f(t).reshape(1) for t in (a.attn1, a.attn2, a.attn3, a.attn4)]).contiguous()
    m2 = _arange(2 * h, dev, P2)                       # padded hidden slot -> logical channel
    p["w_in"], _ = ops.pack_conv(f.project_in.weight, co_map=m2, Co_p=P2, dt=dt)
    p["b_in"] = ops.gather_vec(f.project_in.bias, m2, P2)
    # depthwise 3x3 / 5x5 on the 2h hidden channels -> D = [dw3 | dw5], each half P2 slots wide
    p["idx_dw"] = _arange(2 * h, dev)
    p["w3"], p["b3"] = _f(f.dwconv3x3.weight), _f(f.dwconv3x3.bias)
    p["w3_tma"], _, p["b3_tma"] = ops.pack_dw(f.dwconv3x3.weight, f.dwconv3x3.bias, c_map=m2, C_p=P2)   # tdr_dwconv3x3 (TMA)
    p["w5"], p["b5"] = _f(f.dwconv5x5.weight), _f(f.dwconv5x5.bias)
    # grouped convs: x1 = cat[dw3[:h], dw5[:h]], x2 = cat[dw3[h:], dw5[h:]] (:246-247); group g reads channels 2g, 2g+1
    i = torch.arange(2 * h, device=dev)
    x1 = torch.where(i < h, i, P2 + (i - h))
    x2 = torch.where(i < h, h + i, P2 + h + (i - h))
    p["idx_g3"] = x1.to(I32).contiguous()
    n5 = P2 - h                                       # the second half also writes the pad slots (zeros)
    idx5 = torch.full((n5 * 2,), -1, dtype=I32, device=dev)
    idx5[: 2 * h] = x2.to(I32)
    p["idx_g5"] = idx5
    p["w3_1"], p["b3_1"] = _f(f.dwconv3x3_1.weight), _f(f.dwconv3x3_1.bias)
    w5 = torch.zeros((n5, 2, 5, 5), dtype=F32, device=dev)
    w5[:h] = f.dwconv5x5_1.weight.detach().float()
    p["w5_1"] = w5
    p["b5_1"] = ops.pad_vec(f.dwconv5x5_1.bias, n5)
    p["w_out"], _ = ops.pack_conv(f.project_out.weight, ci_map=m2, Ci_p=P2, dt=dt)
    p["b_out"] = _f(f.project_out.bias)
    p["alpha"] = _f(blk.alpha) if hasattr(blk, "alpha") else None
    return p


def _prep_subnet(net, dt):
    oa, grp = net.layers[0], net.layers[1]
    C_ = net._C
    dev = oa.ca_fc[0].weight.device
    S = dict(C=C_, dt=dt, steps=grp._steps, idx=_arange(C_, dev),
             w1=_f(oa.ca_fc[0].weight), b1=_f(oa.ca_fc[0].bias), w2=_f(oa.ca_fc[2].weight), b2=_f(oa.ca_fc[2].bias),
             pre=_prep_conv(grp.preprocess.op[0], dt), ops=[])
    for ol in grp._ops:
        experts = []
        for op in ol._ops:
            if isinstance(op, SepConv):
                experts.append(("sep", op.op[0].kernel_size[0], _f(op.op[0].weight), _prep_conv(op.op[1], dt),
                                _f(op.op[3].weight), _prep_conv(op.op[4], dt)))
            elif isinstance(op, DilConv):
                experts.append(("dil", op.op[0].kernel_size[0], _f(op.op[0].weight), _prep_conv(op.op[1], dt)))
            else:
                experts.append(("pool",))
        S["ops"].append(dict(experts=experts, w_out=_f(ol._out[0].weight).reshape(C_, 8 * C_)))
    return S


# ----------------------------------------------------------------------------------------------- kernel schedules
def run_block(x32, p):
    """One sparse transformer block (:334-369) on the fp32 residual stream x32 (NHWC view), in place."""
    C_, heads, h, P2, dt = p["C"], p["heads"], p["h"], p["P2"], p["dt"]
    fusion = p["alpha"] is not None
    B, H, W, _ = x32.shape
    dev = x32.device
    xn = ops.rownorm(x32, p["ln_mode"], p["ln1_w"], p["ln1_b"], 1e-5, out=ops.rows16(B, H, W, C_, dev, dt))
    _, qkv = ops.conv_gemm(xn, p["w_qkv"], 3 * C_, bias=p["b_qkv"], out_bf16=ops.rows16(B, H, W, 3 * C_, dev, dt))
    qkv = ops.dwconv3x3(qkv, p["w_qkv_dw"], p["b_qkv_dw"], out=ops.rows16(B, H, W, 3 * C_, dev, dt))
    weff = ops.mdta_weff(qkv, C_, heads, p["temp"], p["w_po"], topk_w=p["topk_w"])
    v = qkv[..., 2 * C_:]
    if fusion:
        x1, _ = ops.conv_gemm(v, weff, C_, Ci=C_, bias=p["b_po"], res2=x32, want="f32", w_batched=True)
    else:
        ops.conv_gemm(v, weff, C_, Ci=C_, bias=p["b_po"], res2=x32, out_f32=x32, w_batched=True)
        x1 = x32
    xn = ops.rownorm(x1, p["ln_mode"], p["ln2_w"], p["ln2_b"], 1e-5, out=xn)
    _, hid = ops.conv_gemm(xn, p["w_in"], P2, bias=p["b_in"])
    D = torch.empty((B, H, W, 2 * P2), dtype=dt, device=dev)
    ops.dwconv3x3(hid, p["w3_tma"], p["b3_tma"], out=D[..., :P2], relu=True)        # pad channels: zero taps -> zeros
    ops.grouped_stencil(hid, p["idx_dw"], p["w5"], p["b5"], 5, D[..., P2:P2 + 2 * h], relu=True)
    Y = torch.empty((B, H, W, P2), dtype=dt, device=dev)
    ops.grouped_stencil(D, p["idx_g3"], p["w3_1"], p["b3_1"], 3, Y[..., :h], relu=True)
    ops.grouped_stencil(D, p["idx_g5"], p["w5_1"], p["b5_1"], 5, Y[..., h:], relu=True)
    if fusion:      # out = (x1 + ffn) * alpha + x0
        ops.conv_gemm(Y, p["w_out"], C_, bias=p["b_out"], scale_ptr=p["alpha"], res1=x1, res2=x32, out_f32=x32)
    else:
        ops.conv_gemm(Y, p["w_out"], C_, bias=p["b_out"], res2=x32, out_f32=x32)


def run_stack(x32, preps, xn=None, nxt=None, tail=None):
    for p in preps:
        run_block(x32, p)
    if tail is not None:
        tail.append(None)
    return x32


def run_subnet(x32, S):
    """MEFC (:522-549) on the fp32 NHWC view x32, in place."""
    B, H, W, C_ = x32.shape
    dev, dt = x32.device, S["dt"]
    emb = torch.empty((B, C_), dtype=F32, device=dev)
    ops.mean_tokens(x32.as_strided((B, 1, H * W, C_), (x32.stride(0), x32.stride(0), x32.stride(2), 1)), 0, H * W, emb)
    gate = ops.mefc_gate(emb, S["w1"], S["b1"], S["w2"], S["b2"], 8)                      # [B, steps, 8]
    s32, _ = ops.conv_gemm(ops.rownorm(x32, 0, dt=dt), S["pre"]["w"], C_, relu=True, want="f32")      # preprocess
    cat16 = torch.empty((B, H, W, 8 * C_), dtype=dt, device=dev)
    t = torch.empty((B, H, W, C_), dtype=dt, device=dev)
    t2 = torch.empty_like(t)
    for i, ol in enumerate(S["ops"]):
        s16 = ops.rownorm(s32, 0, dt=dt)
        for k, ex in enumerate(ol["experts"]):
            dst = cat16[..., k * C_:(k + 1) * C_]
            if ex[0] == "sep":
                _, K, dw1, pw1, dw2, pw2 = ex
                ops.grouped_stencil(s16, S["idx"], dw1, None, K, t)
                ops.conv_gemm(t, pw1["w"], C_, relu=True, out_bf16=t2)
                ops.grouped_stencil(t2, S["idx"], dw2, None, K, t)
                ops.conv_gemm(t, pw2["w"], C_, out_bf16=dst)
            elif ex[0] == "dil":
                _, K, dw1, pw1 = ex
                ops.grouped_stencil(s16, S["idx"], dw1, None, K, t, dil=2)
                ops.conv_gemm(t, pw1["w"], C_, out_bf16=dst)
            else:
                ops.grouped_stencil(s16, S["idx"], None, None, 3, dst, pool=True)
        # s0 = relu(relu(conv_out(cat(states * gate))) + s0): both terms are >= 0, so the outer ReLU is the identity
        wmix = ops.mefc_mix_weights(ol["w_out"], gate[:, i], C_, dt)
        ops.conv_gemm(cat16, wmix, C_, Ci=8 * C_, relu=True, res2=s32, out_f32=s32, w_batched=True)
    ops.copy_rows(s32, dst32=x32)
    return x32


# ----------------------------------------------------------------------------------------------- models
class _DRSBase(MasaMixin, nn.Module):
    _guided_encode = RestormerRefFusion._guided_encode
    _down = RestormerRefFusion._down
    _check = RestormerRefFusion._check
    _prep_key = RestormerRefFusion._prep_key
    prepared = RestormerRefFusion.prepared
    _run_stack = staticmethod(run_stack)
    dual_pixel_task = False
    with_mefc = False

    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], heads=[1, 2, 4, 8],
                 ffn_expansion_factor=2.66, bias=False, LayerNorm_type="WithBias", nf=64, ext_n_blocks=[4, 4, 4, 4],
                 reffusion_n_blocks=[1, 1, 1, 1], reffusion_n_blocks_middle=1, scale=1, num_nbr=1, psize=3, lr_block_size=8,
                 ref_down_block_size=1.5, dilations=[1, 2, 3]):
        super().__init__()
        name = type(self).__name__
        if num_nbr != 1 or psize != 3:
            raise TdrError(f"{name} (B200): only num_nbr=1, psize=3 are implemented (all shipped options)")
        if not 1 <= len(dilations) <= 3:
            raise TdrError(f"{name} (B200): 1..3 dilations supported")
        if nf != dim:
            raise TdrError(f"{name}: nf must equal dim (warped reference features are concatenated channel-for-channel)")
        self.scale, self.num_nbr, self.psize = scale, num_nbr, psize
        self.lr_block_size, self.ref_down_block_size, self.dilations = lr_block_size, ref_down_block_size, list(dilations)
        self.padder_size = 2 ** 3
        kw = dict(ffn_expansion_factor=ffn_expansion_factor, bias=bias, LayerNorm_type=LayerNorm_type)
        d = self.dims = [dim, dim * 2, dim * 4, dim * 8]
        self.masa_enc = Encoder(inp_channels, nf, ext_n_blocks, levels=4)
        self.masa_blk_enc, self.masa_blk_middle, self.masa_blk_dec = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.patch_embed = OverlapPatchEmbed(inp_channels, dim)
        if self.with_mefc:
            self.encoder_level0 = subnet(dim)
        names = ["encoder_level1", "encoder_level2", "encoder_level3", "latent"]
        downs = ["down1_2", "down2_3", "down3_4", None]
        for i in range(4):
            setattr(self, f"masa_blk_enc_level{i + 1}",
                    _blocks(reffusion_n_blocks[i], TransformerResFusionBlock, dim=2 * d[i], num_heads=heads[i], **kw))
            setattr(self, names[i], _blocks(num_blocks[i], TransformerBlock, dim=d[i], num_heads=heads[i], **kw))
            if downs[i]:
                setattr(self, downs[i], Downsample(d[i]))
        self.up4_3 = Upsample(d[3])
        self.reduce_chan_level3 = nn.Conv2d(d[3], d[2], 1, bias=bias)
        self.decoder_level3 = _blocks(num_blocks[2], TransformerBlock, dim=d[2], num_heads=heads[2], **kw)
        self.up3_2 = Upsample(d[2])
        self.reduce_chan_level2 = nn.Conv2d(d[2], d[1], 1, bias=bias)
        self.decoder_level2 = _blocks(num_blocks[1], TransformerBlock, dim=d[1], num_heads=heads[1], **kw)
        self.up2_1 = Upsample(d[1])
        self.decoder_level1 = _blocks(num_blocks[0], TransformerBlock, dim=d[1], num_heads=heads[0], **kw)
        if self.with_mefc:
            self.refinement = subnet(dim=d[1])
        self.output = nn.Conv2d(d[1], out_channels, 3, 1, 1, bias=bias)
        self.nf = nf
        self._prep_cache = None

    def _prepare(self):
        if getattr(self, "_prep_train_flag", False):
            raise TdrError(f"{type(self).__name__} (B200): inference only -- the explicit backward is not implemented")
        P = {}
        dt = P["dt"] = operand_dtype(False)
        for name in ["encoder_level1", "encoder_level2", "encoder_level3", "latent", "decoder_level3", "decoder_level2",
                     "decoder_level1"] + [f"masa_blk_enc_level{i}" for i in range(1, 5)]:
            P[name] = [_prep_block(b) for b in getattr(self, name)]
        for name in ["down1_2", "down2_3", "down3_4", "up4_3", "up3_2", "up2_1"]:
            P[name] = _prep_conv(getattr(self, name).body[0], dt)
        for name in ["reduce_chan_level3", "reduce_chan_level2"]:
            P[name] = _prep_conv(getattr(self, name), dt)
        if self.with_mefc:
            P["encoder_level0"] = _prep_subnet(self.encoder_level0, dt)
            P["refinement"] = _prep_subnet(self.refinement, dt)
        P["patch_embed"] = dict(w=_f(self.patch_embed.proj.weight), b=_f(self.patch_embed.proj.bias))
        ow = self.output.weight
        co = ow.shape[0]
        w8 = torch.zeros(8, ow.shape[1], 3, 3, dtype=ow.dtype, device=ow.device)
        w8[:co] = ow.detach()
        P["output"] = dict(w=ops.pack_conv_weight(w8, dt=dt), b=ops.pad_vec(self.output.bias, 8), Co=co)
        P["masa_enc"] = self.prepare_masa_enc()
        return P

    def _after_patch_embed(self, P, x32):
        if self.with_mefc:                                  # :1063 inp_enc_level0 = encoder_level0(patch_embed(x))
            run_subnet(x32, P["encoder_level0"])

    def forward(self, inp_img, ref_img):
        """:913-1123.  NCHW in, NCHW out, arbitrary H, W (zero-padded to x64, cropped)."""
        self._check(inp_img, ref_img)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise TdrError(f"{type(self).__name__} (B200): inference only -- call it under torch.no_grad()")
        P = self.prepared()
        dt, d = P["dt"], self.dims
        xs, _, lq32, (oh, ow), _, _ = self._guided_encode(P, inp_img, ref_img)
        e1, e2, e3, lat = xs
        dev = lat.device

        def up_cat_reduce(x32, enc, up, red, Cn):
            b, hh, ww, _ = x32.shape
            cat16 = torch.empty((b, hh * 2, ww * 2, 2 * Cn), dtype=dt, device=dev)
            ops.conv_gemm(ops.rownorm(x32, 0, dt=dt), P[up]["w"], P[up]["Co"], k=3, pad=1, out_bf16=cat16[..., :Cn],
                          store_mode=2)
            ops.copy_rows(enc, dst16=cat16[..., Cn:])
            y32, _ = ops.conv_gemm(cat16, P[red]["w"], Cn, bias=P[red]["b"], want="f32")
            return y32

        d3 = run_stack(up_cat_reduce(lat, e3, "up4_3", "reduce_chan_level3", d[2]), P["decoder_level3"])
        d2 = run_stack(up_cat_reduce(d3, e2, "up3_2", "reduce_chan_level2", d[1]), P["decoder_level2"])
        b, hh, ww, _ = d2.shape
        d1 = torch.empty((b, hh * 2, ww * 2, d[1]), dtype=F32, device=dev)
        ops.conv_gemm(ops.rownorm(d2, 0, dt=dt), P["up2_1"]["w"], P["up2_1"]["Co"], k=3, pad=1, out_f32=d1[..., :d[0]],
                      store_mode=2)
        ops.copy_rows(e1, dst32=d1[..., d[0]:])
        run_stack(d1, P["decoder_level1"])
        if self.with_mefc:
            run_subnet(d1, P["refinement"])
        o8, _ = ops.conv_gemm(ops.rownorm(d1, 0, dt=dt), P["output"]["w"], 8, k=3, pad=1, bias=P["output"]["b"], want="f32")
        return ops.nhwc_to_nchw(o8[..., :P["output"]["Co"]], oh, ow, res=lq32)


class DRSformer200L_SPA_RefFusion(_DRSBase):
    """network_drsformer_guided_arch_200L_SPA.py :582-1020 (no MEFC).  Faithful to a quirk of that file: the level-1 fusion
    result is assigned to ``inp_enc_level0`` (:973) and never read -- ``encoder_level1`` consumes the patch embedding
    (:975) -- so ``masa_blk_enc_level1`` holds parameters that do not influence the output and is not run."""
    with_mefc = False
    skip_level1_fusion = True


class DRSformerRefFusion(_DRSBase):
    """network_drsformer_guided_arch.py :679-1123."""
    with_mefc = True
