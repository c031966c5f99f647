"""B200-native ``Restormer`` and ``RestormerRefFusion``.

Drop-in for the classes of the same name in the reference's
``models/archs/network_restormer_guided_arch.py`` (:396-501, :504-964): same constructor kwargs (option files
``options/train_restoration/003*.yml``, ``017*.yml``), same call signature ``net(inp)`` / ``net(lq, ref)`` on NCHW
tensors, same ``state_dict`` keys and shapes, every guidance parameter name contains ``masa``.

The module tree only HOLDS parameters; the forward pass is a schedule of libtdr_sm100 kernels over NHWC buffers
(fp32 residual stream, bf16 GEMM operands) -- see DESIGN.md.  There is no PyTorch fallback: without the CUDA
library, or on a non-CUDA tensor, the forward raises.

RestormerRefFusion implements the reference's *consistent* reading of its own forward: the MASA encoder has four
levels, so the deepest feature is the 1/8-scale one (upstream indexes ``feat[4]``, an off-by-one that raises
IndexError as shipped).
"""
import math

import torch
import torch.nn as nn

from .. import ops
from ..lib import TdrError
from .masa import Encoder, MasaMixin, MasaTrainMixin, ResidualBlock, prep_conv as _prep_conv, conv3x3, _f  # noqa: F401
from .restormer_train import GuidedRestormerTrainMixin, RestormerTrainMixin, train_call

F32, BF16, F16 = torch.float32, torch.bfloat16, torch.float16


operand_dtype = ops.operand_dtype


# ----------------------------------------------------------------------------------------------- parameter holders
class _Norm(nn.Module):
    def __init__(self, dim, with_bias):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        if with_bias:
            self.bias = nn.Parameter(torch.zeros(dim))
        else:
            self.bias = None


class LayerNorm(nn.Module):
    """Keys ``body.weight`` (+ ``body.bias`` for WithBias) as in the reference (:208-218)."""

    def __init__(self, dim, LayerNorm_type):
        super().__init__()
        self.body = _Norm(dim, LayerNorm_type != "BiasFree")


class Attention(nn.Module):
    def __init__(self, dim, num_heads, bias):
        super().__init__()
        self.num_heads = num_heads
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.qkv = nn.Conv2d(dim, dim * 3, 1, bias=bias)
        self.qkv_dwconv = nn.Conv2d(dim * 3, dim * 3, 3, 1, 1, groups=dim * 3, bias=bias)
        self.project_out = nn.Conv2d(dim, dim, 1, bias=bias)


class FeedForward(nn.Module):
    def __init__(self, dim, ffn_expansion_factor, bias):
        super().__init__()
        hidden = int(dim * ffn_expansion_factor)
        self.project_in = nn.Conv2d(dim, hidden * 2, 1, bias=bias)
        self.dwconv = nn.Conv2d(hidden * 2, hidden * 2, 3, 1, 1, groups=hidden * 2, bias=bias)
        self.project_out = nn.Conv2d(hidden, dim, 1, bias=bias)


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


class OverlapPatchEmbed(nn.Module):
    def __init__(self, in_c=3, embed_dim=48, bias=False):
        super().__init__()
        self.proj = nn.Conv2d(in_c, embed_dim, 3, 1, 1, bias=bias)


class Downsample(nn.Module):
    def __init__(self, n_feat):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(n_feat, n_feat // 2, 3, 1, 1, bias=False), nn.PixelUnshuffle(2))


class Upsample(nn.Module):
    def __init__(self, n_feat):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(n_feat, n_feat * 2, 3, 1, 1, bias=False), nn.PixelShuffle(2))


def _blocks(n, cls, **kw):
    return nn.Sequential(*[cls(**kw) for _ in range(n)])


# ----------------------------------------------------------------------------------------------- weight preparation
def _block_maps(blk, h, hp, dev):
    """int32 padded->logical channel maps of the GDFN halves (static per block, cached on the module)."""
    m = getattr(blk, "_tdr_maps", None)
    if m is None or m[0].device != dev:
        idx2 = torch.cat([torch.arange(h, device=dev), hp + torch.arange(h, device=dev)])
        m2 = torch.full((2 * hp,), -1, dtype=torch.int32, device=dev)
        m2[idx2] = torch.arange(2 * h, dtype=torch.int32, device=dev)
        m1 = torch.full((hp,), -1, dtype=torch.int32, device=dev)
        m1[:h] = torch.arange(h, dtype=torch.int32, device=dev)
        m = (m2, m1)
        blk._tdr_maps = m
    return m


def _prep_block(blk: TransformerBlock, train=False):
    """Pack one transformer block's parameters for the kernels (bf16 GEMM weights, padded GDFN halves): one launch per
    parameter.  train=True also emits the transposed / flipped twins the data-gradient kernels consume."""
    a, f = blk.attn, blk.ffn
    C_ = a.qkv.in_channels
    h = f.project_out.in_channels
    hp = ops.round_up(h, 8)
    dev = a.qkv.weight.device
    m2, m1 = _block_maps(blk, h, hp, dev)
    dt = operand_dtype(train)
    p = dict(C=C_, heads=a.num_heads, h=h, hp=hp, map_2h=m2, map_h=m1, mod=blk, dt=dt)
    p["ln1_w"], p["ln1_b"] = _f(blk.norm1.body.weight), _f(blk.norm1.body.bias)
    p["ln2_w"], p["ln2_b"] = _f(blk.norm2.body.weight), _f(blk.norm2.body.bias)
    p["ln_mode"] = 1 if blk.norm1.body.bias is not None else 2
    p["w_qkv"], p["w_qkv_T"] = ops.pack_conv(a.qkv.weight, dgrad=train, dt=dt)
    p["b_qkv"] = _f(a.qkv.bias)
    p["w_qkv_dw"], p["w_qkv_dw_f"], p["b_qkv_dw"] = ops.pack_dw(a.qkv_dwconv.weight, a.qkv_dwconv.bias, flip=train)
    p["temp"] = _f(a.temperature).reshape(-1)
    p["w_po"] = _f(a.project_out.weight).reshape(C_, C_)
    p["b_po"] = _f(a.project_out.bias)
    p["w_in"], p["w_in_T"] = ops.pack_conv(f.project_in.weight, co_map=m2, Co_p=2 * hp, dgrad=train, dt=dt)
    p["b_in"] = ops.gather_vec(f.project_in.bias, m2, 2 * hp)
    p["w_dw"], p["w_dw_f"], p["b_dw"] = ops.pack_dw(f.dwconv.weight, f.dwconv.bias, c_map=m2, C_p=2 * hp, flip=train)
    p["w_out"], p["w_out_T"] = ops.pack_conv(f.project_out.weight, ci_map=m1, Ci_p=hp, dgrad=train, dt=dt)
    p["b_out"] = _f(f.project_out.bias)
    p["alpha"] = _f(blk.alpha) if hasattr(blk, "alpha") else None
    p["train"] = train
    return p


# ----------------------------------------------------------------------------------------------- kernel schedules
def _ln_fusable(p):
    """Plain (non Res-fusion) blocks whose width fits the conv epilogue's fused LayerNorm (ops.conv_ln_ok)."""
    return p["alpha"] is None and ops.conv_ln_ok(p["C"])


def run_block(x32, p, xn=None, nxt=None):
    """One (Res-fusion) transformer block on the fp32 residual stream x32 (NHWC view), updated IN PLACE.

    Reference :318-331 / :334-353.  Kernel sequence: LN -> qkv 1x1 (tcgen05) -> depthwise 3x3 -> Gram (tcgen05) ->
    softmax+fold -> attn.v.project_out + residual (tcgen05) -> LN -> project_in (tcgen05) -> depthwise 3x3 + GELU gate
    -> project_out + residual (tcgen05).  For C <= 96 the two convs that finish a residual row also write the
    LayerNorm that follows it (norm2 of this block; norm1 of ``nxt``, the next block of the stack), so the norm
    kernels' re-read of the fp32 stream disappears: ``xn`` is that already-normalised input when the producer made it.
    Returns the normalised input for ``nxt`` (or None when ``nxt`` has to run its own norm1).
    """
    C_, heads, hp = p["C"], p["heads"], p["hp"]
    fusion = p["alpha"] is not None
    fuse_ln = _ln_fusable(p)
    B, H, W, _ = x32.shape
    dev = x32.device
    dt = p["dt"]
    # 16-bit operands with 128 B-aligned row pitch (ops.rows16): C = 48 / 96 rows would straddle lines otherwise
    if xn is None:
        xn = ops.rownorm(x32, p["ln_mode"], p["ln1_w"], p["ln1_b"], 1e-5, out=ops.rows16(B, H, W, C_, dev, dt))
    _, qkv = ops.conv_gemm(xn, p["w_qkv"], 3 * C_, bias=p["b_qkv"], out_bf16=ops.rows16(B, H, W, 3 * C_, dev, dt))
    qkv = ops.dwconv3x3(qkv, p["w_qkv_dw"], p["b_qkv_dw"], out=ops.rows16(B, H, W, 3 * C_, dev, dt))
    weff = ops.mdta_weff(qkv, C_, heads, p["temp"], p["w_po"])
    v = qkv[..., 2 * C_:]
    if fusion:
        x1, _ = ops.conv_gemm(v, weff, C_, Ci=C_, bias=p["b_po"], res2=x32, want="f32", w_batched=True)
    else:
        ln2 = (p["ln_mode"], p["ln2_w"], p["ln2_b"], 1e-5, xn) if fuse_ln else None     # xn1 is consumed: reuse it
        ops.conv_gemm(v, weff, C_, Ci=C_, bias=p["b_po"], res2=x32, out_f32=x32, w_batched=True, ln=ln2)
        x1 = x32
    if fusion or not fuse_ln:
        xn = ops.rownorm(x1, p["ln_mode"], p["ln2_w"], p["ln2_b"], 1e-5, out=xn)
    _, hid = ops.conv_gemm(xn, p["w_in"], 2 * hp, bias=p["b_in"])
    if ops.gdfn_tail_ok(hid, p["w_out"], C_):
        # depthwise 3x3 + GELU gate + project_out + residual(s) in one kernel: the gated tensor never reaches HBM
        if fusion:  # out = (x1 + ffn) * alpha + x0
            ops.gdfn_tail(hid, p["w_dw"], p["b_dw"], p["w_out"], C_, bias=p["b_out"], scale_ptr=p["alpha"], res1=x1,
                          res2=x32, out=x32)
        else:
            ops.gdfn_tail(hid, p["w_dw"], p["b_dw"], p["w_out"], C_, bias=p["b_out"], res2=x32, out=x32)
        return None
    g = ops.dwconv3x3(hid, p["w_dw"], p["b_dw"], gate=1)
    if fusion:      # out = (x1 + ffn) * alpha + x0
        ops.conv_gemm(g, p["w_out"], C_, bias=p["b_out"], scale_ptr=p["alpha"], res1=x1, res2=x32, out_f32=x32)
        return None
    ln1 = None
    if fuse_ln and nxt is not None and _ln_fusable(nxt) and nxt["C"] == C_:
        ln1 = (nxt["ln_mode"], nxt["ln1_w"], nxt["ln1_b"], 1e-5, xn)                    # xn2 is consumed: reuse it
    ops.conv_gemm(g, p["w_out"], C_, bias=p["b_out"], res2=x32, out_f32=x32, ln=ln1)
    return xn if ln1 is not None else None


def run_stack(x32, preps, xn=None, nxt=None, tail=None):
    """Blocks of one level, in place.  ``xn``: LayerNorm of x32 already made by the producer; ``nxt``: first block of the
    stack that continues on the same stream (its norm1 is then emitted by this stack's last conv and appended to ``tail``)."""
    for i, p in enumerate(preps):
        xn = run_block(x32, p, xn, preps[i + 1] if i + 1 < len(preps) else nxt)
    if tail is not None:
        tail.append(xn)
    return x32


class _RestormerBase(nn.Module):
    """Shared U-Net body (:412-461) + decoder schedule."""

    _run_stack = staticmethod(run_stack)       # block-stack runner (the DRSformer family swaps in its own blocks)

    def _after_patch_embed(self, P, x32):      # hook: DRSformer's MEFC ``encoder_level0`` runs here
        pass

    def _build_body(self, inp_channels, out_channels, dim, num_blocks, num_refinement_blocks, heads,
                    ffn_expansion_factor, bias, LayerNorm_type, dual_pixel_task, fusion_blocks=None):
        kw = dict(ffn_expansion_factor=ffn_expansion_factor, bias=bias, LayerNorm_type=LayerNorm_type)
        self.patch_embed = OverlapPatchEmbed(inp_channels, dim)
        dims = [dim, dim * 2, dim * 4, dim * 8]
        names = ["encoder_level1", "encoder_level2", "encoder_level3", "latent"]
        downs = [None, "down1_2", "down2_3", "down3_4"]
        for i in range(4):
            if downs[i]:
                setattr(self, downs[i], Downsample(dims[i - 1]))
            if fusion_blocks is not None:
                setattr(self, f"masa_blk_enc_level{i + 1}",
                        _blocks(fusion_blocks[i], TransformerResFusionBlock, dim=2 * dims[i], num_heads=heads[i], **kw))
            setattr(self, names[i], _blocks(num_blocks[i], TransformerBlock, dim=dims[i], num_heads=heads[i], **kw))
        self.up4_3 = Upsample(dims[3])
        self.reduce_chan_level3 = nn.Conv2d(dims[3], dims[2], 1, bias=bias)
        self.decoder_level3 = _blocks(num_blocks[2], TransformerBlock, dim=dims[2], num_heads=heads[2], **kw)
        self.up3_2 = Upsample(dims[2])
        self.reduce_chan_level2 = nn.Conv2d(dims[2], dims[1], 1, bias=bias)
        self.decoder_level2 = _blocks(num_blocks[1], TransformerBlock, dim=dims[1], num_heads=heads[1], **kw)
        self.up2_1 = Upsample(dims[1])
        self.decoder_level1 = _blocks(num_blocks[0], TransformerBlock, dim=dims[1], num_heads=heads[0], **kw)
        self.refinement = _blocks(num_refinement_blocks, TransformerBlock, dim=dims[1], num_heads=heads[0], **kw)
        self.dual_pixel_task = dual_pixel_task
        if dual_pixel_task:
            self.skip_conv = nn.Conv2d(dim, dims[1], 1, bias=bias)
        self.output = nn.Conv2d(dims[1], out_channels, 3, 1, 1, bias=bias)
        self.dims = dims
        self._prep_cache = None

    # ---- weight cache -------------------------------------------------------------------------
    def _prep_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def prepared(self, train=False):
        """Packed kernel operands, rebuilt when a parameter changed; one cache per mode: inference packs are IEEE fp16
        (operand_dtype), train=True packs are bf16 and come with the data-gradient twins."""
        key = self._prep_key()
        if not isinstance(self._prep_cache, dict):
            self._prep_cache = {}
        c = self._prep_cache.get(bool(train))
        if c is None or c[0] != key:
            with torch.no_grad():
                self._prep_train_flag = train
                c = self._prep_cache[bool(train)] = (key, self._prepare())
        return c[1]

    def _prepare_body(self):
        P = {}
        train = getattr(self, "_prep_train_flag", False)
        dt = P["dt"] = operand_dtype(train)
        for name in ["encoder_level1", "encoder_level2", "encoder_level3", "latent", "decoder_level3",
                     "decoder_level2", "decoder_level1", "refinement"] + \
                    [f"masa_blk_enc_level{i}" for i in range(1, 5) if hasattr(self, f"masa_blk_enc_level{i}")]:
            P[name] = [_prep_block(b, train) for b in getattr(self, name)]
        for name in ["down1_2", "down2_3", "down3_4", "up4_3", "up3_2", "up2_1"]:
            P[name] = _prep_conv(getattr(self, name).body[0], dt)
        for name in ["reduce_chan_level3", "reduce_chan_level2"]:
            P[name] = _prep_conv(getattr(self, name), dt)
        P["patch_embed"] = dict(w=_f(self.patch_embed.proj.weight), b=_f(self.patch_embed.proj.bias))
        if not train and self.patch_embed.proj.in_channels <= 8:
            P["patch_embed"]["w16"] = ops.pack_conv(self.patch_embed.proj.weight, Ci_p=8, dt=dt)[0]
        # output conv (:640): 3 (or 1) output channels are zero-padded to 8 so that it runs on the tensor-core path
        ow = self.output.weight
        co = ow.shape[0]
        w8 = torch.zeros(8, ow.shape[1], 3, 3, dtype=ow.dtype, device=ow.device)
        w8[:co] = ow.detach()
        P["output"] = dict(w=ops.pack_conv_weight(w8, dt=dt), b=ops.pad_vec(self.output.bias, 8), Co=co)
        if self.dual_pixel_task:
            P["skip_conv"] = _prep_conv(self.skip_conv, dt)
        return P

    def _wants_grad(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    # ---- schedules ----------------------------------------------------------------------------
    def _check(self, *ts):
        for t in ts:
            if not t.is_cuda:
                raise TdrError("textualdegremoval_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")

    def _down(self, x32, pc, out32):
        """Downsample (:372-380): conv3x3 C->C/2 then PixelUnshuffle(2), written straight into out32."""
        x16 = ops.rownorm(x32, 0, dt=pc["w"].dtype)
        ops.conv_gemm(x16, pc["w"], pc["Co"], k=3, pad=1, out_f32=out32, store_mode=1)

    def _decode(self, P, lat, e1, e2, e3, x_in1):
        """Decoder half (:477-501).  e*: fp32 NHWC views of the encoder outputs."""
        d = self.dims
        B, H8, W8, _ = lat.shape
        dev = lat.device
        dt = P["dt"]

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
        tail = []
        run_stack(d1, P["decoder_level1"], nxt=P["refinement"][0] if P["refinement"] else None, tail=tail)
        run_stack(d1, P["refinement"], xn=tail[0])
        if self.dual_pixel_task:      # :494-496  out = output(d1 + skip_conv(inp_enc_level1))
            ops.conv_gemm(ops.rownorm(x_in1, 0, dt=dt), P["skip_conv"]["w"], d[1], bias=P["skip_conv"]["b"], res2=d1,
                          out_f32=d1)
        o8, _ = ops.conv_gemm(ops.rownorm(d1, 0, dt=dt), P["output"]["w"], 8, k=3, pad=1, bias=P["output"]["b"], want="f32")
        return o8[..., :P["output"]["Co"]]


class Restormer(RestormerTrainMixin, _RestormerBase):
    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4,
                 heads=[1, 2, 4, 8], ffn_expansion_factor=2.66, bias=False, LayerNorm_type="WithBias",
                 dual_pixel_task=False):
        super().__init__()
        self._build_body(inp_channels, out_channels, dim, num_blocks, num_refinement_blocks, heads,
                         ffn_expansion_factor, bias, LayerNorm_type, dual_pixel_task)

    def _prepare(self):
        return self._prepare_body()

    def forward(self, inp_img):
        """:463-501.  inp_img NCHW; H and W must be multiples of 8 (as in the reference, which fails otherwise).
        Under autograd (grad mode on and trainable parameters) the call records the training tape and is
        differentiable w.r.t. the parameters (restormer_train.NetFunction)."""
        self._check(inp_img)
        if self._wants_grad():
            return train_call(self, inp_img)
        B, Cin, H, W = inp_img.shape
        if H % 8 or W % 8:
            raise ValueError(f"Restormer needs H, W multiples of 8 (got {H}x{W})")
        P = self.prepared()
        d = self.dims
        dev = inp_img.device
        inp32 = ops.nchw_to_nhwc(inp_img, H, W)
        x1 = torch.empty((B, H, W, d[0]), dtype=F32, device=dev)
        if ops.SMALL_CI_TC and "w16" in P["patch_embed"]:
            ops.conv_gemm(ops.image_to_rows16(inp_img, H, W, F16), P["patch_embed"]["w16"], d[0], Ci=8, k=3, pad=1,
                          bias=P["patch_embed"]["b"], out_f32=x1)
        else:
            ops.conv3x3_small_ci(inp32, P["patch_embed"]["w"], P["patch_embed"]["b"], out_f32=x1)
        x_in1 = x1.clone() if self.dual_pixel_task else None
        e1 = run_stack(x1, P["encoder_level1"])
        e2 = torch.empty((B, H // 2, W // 2, d[1]), dtype=F32, device=dev)
        self._down(e1, P["down1_2"], e2)
        run_stack(e2, P["encoder_level2"])
        e3 = torch.empty((B, H // 4, W // 4, d[2]), dtype=F32, device=dev)
        self._down(e2, P["down2_3"], e3)
        run_stack(e3, P["encoder_level3"])
        lat = torch.empty((B, H // 8, W // 8, d[3]), dtype=F32, device=dev)
        self._down(e3, P["down3_4"], lat)
        run_stack(lat, P["latent"])
        out = self._decode(P, lat, e1, e2, e3, x_in1)
        return ops.nhwc_to_nchw(out, H, W, res=None if self.dual_pixel_task else inp32)     # + inp_img (:499)


class RestormerRefFusion(GuidedRestormerTrainMixin, MasaTrainMixin, MasaMixin, _RestormerBase):
    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4,
                 heads=[1, 2, 4, 8], ffn_expansion_factor=2.66, bias=False, LayerNorm_type="WithBias",
                 dual_pixel_task=False, nf=64, ext_n_blocks=[4, 4, 4, 4], reffusion_n_blocks=[1, 1, 1, 1],
                 reffusion_n_blocks_middle=1, scale=1, num_nbr=1, psize=3, lr_block_size=8, ref_down_block_size=1.5,
                 dilations=[1, 2, 3]):
        super().__init__()
        if num_nbr != 1 or psize != 3:
            raise TdrError("RestormerRefFusion (B200): only num_nbr=1, psize=3 are implemented (all shipped options)")
        if not 1 <= len(dilations) <= 3:
            raise TdrError("RestormerRefFusion (B200): 1..3 dilations supported")
        self.scale, self.num_nbr, self.psize = scale, num_nbr, psize
        self.lr_block_size, self.ref_down_block_size, self.dilations = lr_block_size, ref_down_block_size, list(dilations)
        self.padder_size = 2 ** 3
        self.masa_enc = Encoder(inp_channels, nf, ext_n_blocks, levels=4)
        # empty lists kept for key/structure parity with the reference (:547-549)
        self.masa_blk_enc = nn.ModuleList()
        self.masa_blk_middle = nn.ModuleList()
        self.masa_blk_dec = nn.ModuleList()
        self._build_body(inp_channels, out_channels, dim, num_blocks, num_refinement_blocks, heads,
                         ffn_expansion_factor, bias, LayerNorm_type, dual_pixel_task, fusion_blocks=reffusion_n_blocks)
        self.nf = nf
        if nf != dim:
            raise TdrError("RestormerRefFusion: nf must equal dim (the warped reference features are concatenated "
                           "channel-for-channel with the U-Net features, :907-936)")

    def _prepare(self):
        P = self._prepare_body()
        enc = self.prepare_masa_enc()
        P["masa_enc"] = enc
        return P

    def _guided_encode(self, P, inp_img, ref_img, return_aux=False):
        """MASA guidance + the four encoder levels (:753-936).  Returns (per-level fp32 NHWC encoder outputs, the level-1
        input copy for ``dual_pixel_task``, the padded NHWC lq image, (H, W) of the input, aux, the [x || warp] buffers)."""
        x_in1 = None
        d = self.dims
        dev = inp_img.device
        B, _, oh, ow = inp_img.shape
        mult = self.padder_size * self.lr_block_size
        h, w = ops.round_up(oh, mult), ops.round_up(ow, mult)
        hr, wr = ops.round_up(ref_img.shape[2], mult), ops.round_up(ref_img.shape[3], mult)
        E = P["masa_enc"]
        if (h, w) == (hr, wr):           # shared weights: lq and ref run through the encoder as one batch
            both = torch.empty((2 * B, h, w, inp_img.shape[1]), dtype=F32, device=dev)
            lq32, ref32 = both[:B], both[B:]
            ops.nchw_to_nhwc_into(inp_img, h, w, dst32=lq32)
            ops.nchw_to_nhwc_into(ref_img, hr, wr, dst32=ref32)
            both16 = lq16 = None
            if ops.SMALL_CI_TC:          # 8-channel fp16 image rows: conv_L1 / patch_embed run on tdr_conv_gemm
                both16 = torch.empty((2 * B, h, w, 8), dtype=F16, device=dev)
                ops.image_to_rows16(inp_img, h, w, F16, out=both16[:B])
                ops.image_to_rows16(ref_img, hr, wr, F16, out=both16[B:])
                lq16 = both16[:B]
            fb, d32 = self._masa_encode(E, both, img16=both16 if ops.SMALL_CI_TC_MASA else None)
            f_lq, f_ref, lq_d32, ref_d32 = [t[:B] for t in fb], [t[B:] for t in fb], d32[:B], d32[B:]
        else:
            lq32, ref32 = ops.nchw_to_nhwc(inp_img, h, w), ops.nchw_to_nhwc(ref_img, hr, wr)
            lq16 = ops.image_to_rows16(inp_img, h, w, F16) if ops.SMALL_CI_TC else None
            ref16 = ops.image_to_rows16(ref_img, hr, wr, F16) if ops.SMALL_CI_TC else None
            m16 = ops.SMALL_CI_TC_MASA
            (f_lq, lq_d32), (f_ref, ref_d32) = (self._masa_encode(E, lq32, img16=lq16 if m16 else None),
                                                self._masa_encode(E, ref32, img16=ref16 if m16 else None))
        # fusion buffers [x || warp] per level, fp32 residual streams
        fbuf = [torch.empty((B, h >> i, w >> i, 2 * d[i]), dtype=F32, device=dev) for i in range(4)]
        aux = self._masa_warp(lq_d32, ref_d32, f_ref, h, w, hr, wr, [fbuf[i][..., d[i]:] for i in range(4)])
        if return_aux:                   # the fusion blocks overwrite the warp halves in place
            aux.update(feat_lq=f_lq, feat_ref=f_ref, deep32_lq=lq_d32, deep32_ref=ref_d32, deep_scale=self._masa_last_scale[-1],
                       warps=[fbuf[i][..., d[i]:].clone() for i in range(4)])
        if lq16 is not None and "w16" in P["patch_embed"]:
            ops.conv_gemm(lq16, P["patch_embed"]["w16"], d[0], Ci=8, k=3, pad=1, bias=P["patch_embed"]["b"],
                          out_f32=fbuf[0][..., :d[0]])
        else:
            ops.conv3x3_small_ci(lq32, P["patch_embed"]["w"], P["patch_embed"]["b"], out_f32=fbuf[0][..., :d[0]])
        self._after_patch_embed(P, fbuf[0][..., :d[0]])
        enc_names = ["encoder_level1", "encoder_level2", "encoder_level3", "latent"]
        downs = [None, "down1_2", "down2_3", "down3_4"]
        xs = []
        for i in range(4):
            if i:
                self._down(xs[-1], P[downs[i]], fbuf[i][..., :d[i]])
            if not (i == 0 and getattr(self, "skip_level1_fusion", False)):
                self._run_stack(fbuf[i], P[f"masa_blk_enc_level{i + 1}"])  # fuse on 2C channels, keep the first C (:907-909)
            x = fbuf[i][..., :d[i]]
            if i == 0 and self.dual_pixel_task:                      # skip_conv reads inp_enc_level1 (:957-959); the
                x_in1 = torch.empty((B, h, w, d[0]), dtype=F32, device=dev)     # encoder stack updates x in place
                ops.copy_rows(x, dst32=x_in1)
            self._run_stack(x, P[enc_names[i]])
            xs.append(x)
        return xs, x_in1, lq32, (oh, ow), aux, fbuf

    def forward(self, inp_img, ref_img, return_aux=False):
        """:747-964 (with the B1 index shift).  NCHW in, NCHW out, arbitrary H, W (zero-padded to x64, cropped)."""
        self._check(inp_img, ref_img)
        if self._wants_grad() and not return_aux:
            return train_call(self, inp_img, ref_img)
        P = self.prepared()
        xs, x_in1, lq32, (oh, ow), aux, _ = self._guided_encode(P, inp_img, ref_img, return_aux)
        out = self._decode(P, xs[3], xs[0], xs[1], xs[2], x_in1 if self.dual_pixel_task else None)
        out = ops.nhwc_to_nchw(out, oh, ow, res=None if self.dual_pixel_task else lq32)      # + inp_img (:962)
        return (out, aux) if return_aux else out
