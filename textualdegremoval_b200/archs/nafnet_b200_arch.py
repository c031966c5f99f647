"""B200-native ``NAFNet`` and ``NAFNetRefFusion``.

Drop-in for the classes of the same name in the reference's ``models/archs/network_nafnet_guided_arch.py``
(:305-386, :389-740): same constructor kwargs (option ``002_nafnet_single_image_motion_deblurring.yml``), same call
signature, same ``state_dict`` keys and shapes.  The module tree holds parameters; the forward
is a schedule of libtdr_sm100 kernels over NHWC buffers (fp32 residual stream, bf16 GEMM operands).

Schedule of one NAFBlock (:178-238):
  rownorm(LayerNorm2d, eps 1e-6) -> conv_gemm(conv1) -> dwconv3x3 + SimpleGate -> naf_sca_fold (avg-pool, SCA 1x1,
  fold ``x * sca(x)`` and ``beta`` into per-sample conv3 weights) -> conv_gemm(conv3, +x) -> rownorm ->
  conv_gemm(conv4) -> gate_mul -> conv_gemm(conv5 scaled by gamma, +y)

``NAFNetRefFusion`` reads ``reffusion_n_blocks[index + 1]`` for the middle fusion stage (:463-465), i.e. needs FIVE
entries while the shipped option passes four (IndexError upstream): a 4-entry list is accepted here
by repeating its last entry.
"""
import torch
import torch.nn as nn

from .. import ops
from ..lib import TdrError
from .masa import Encoder, MasaMixin, MasaTrainMixin, _f
from .nafnet_train import GuidedNAFTrainMixin, NAFTrainMixin
from .restormer_train import train_call

operand_dtype = ops.operand_dtype

F32, BF16 = torch.float32, torch.bfloat16


class LayerNorm2d(nn.Module):
    def __init__(self, channels, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))
        self.eps = eps


class NAFBlock(nn.Module):
    def __init__(self, c, DW_Expand=2, FFN_Expand=2, drop_out_rate=0.0):
        super().__init__()
        if drop_out_rate > 0:
            raise TdrError("NAFBlock (B200): dropout is not implemented (no shipped option uses it)")
        dw = c * DW_Expand
        self.conv1 = nn.Conv2d(c, dw, 1)
        self.conv2 = nn.Conv2d(dw, dw, 3, 1, 1, groups=dw)
        self.conv3 = nn.Conv2d(dw // 2, c, 1)
        self.sca = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(dw // 2, dw // 2, 1))
        ffn = FFN_Expand * c
        self.conv4 = nn.Conv2d(c, ffn, 1)
        self.conv5 = nn.Conv2d(ffn // 2, c, 1)
        self.norm1 = LayerNorm2d(c)
        self.norm2 = LayerNorm2d(c)
        self.beta = nn.Parameter(torch.zeros((1, c, 1, 1)))
        self.gamma = nn.Parameter(torch.zeros((1, c, 1, 1)))


class NAFResFuseBlock(NAFBlock):
    """Same math as NAFBlock on the concatenated [x || warped-ref] channels (:241-302)."""


def _prep_naf(blk: NAFBlock, dt=torch.float16):
    """dt: 16-bit operand format (restormer_b200_arch.operand_dtype: fp16 for inference, bf16 for the training tape)."""
    c = blk.conv1.in_channels
    dw = blk.conv1.out_channels
    ffn = blk.conv4.out_channels
    beta, gamma = _f(blk.beta).reshape(-1), _f(blk.gamma).reshape(-1)
    p = dict(C=c, dw=dw, ffn=ffn, dt=dt)
    p["n1_w"], p["n1_b"], p["n2_w"], p["n2_b"] = _f(blk.norm1.weight), _f(blk.norm1.bias), _f(blk.norm2.weight), _f(blk.norm2.bias)
    p["eps"] = blk.norm1.eps
    p["w1"], p["b1"] = ops.pack_conv_weight(blk.conv1.weight, dt=dt), _f(blk.conv1.bias)
    p["w2"], p["b2"] = ops.pack_dw_weight(blk.conv2.weight), _f(blk.conv2.bias)
    p["w_sca"], p["b_sca"] = _f(blk.sca[1].weight).reshape(dw // 2, dw // 2), _f(blk.sca[1].bias)
    p["w3"] = _f(blk.conv3.weight).reshape(c, dw // 2)
    p["b3_beta"] = _f(blk.conv3.bias) * beta                       # y = inp + beta * (conv3(.) + b3)
    p["beta"] = beta
    p["w4"], p["b4"] = ops.pack_conv_weight(blk.conv4.weight, dt=dt), _f(blk.conv4.bias)
    p["w5"] = ops.pack_conv_weight(blk.conv5.weight.detach() * gamma.view(-1, 1, 1, 1), dt=dt)   # out = y + gamma * (conv5 + b5)
    p["b5_gamma"] = _f(blk.conv5.bias) * gamma
    return p


def run_naf_block(x32, p):
    """One NAFBlock on the fp32 residual stream x32 (NHWC view), updated in place."""
    c = p["C"]
    xn = ops.rownorm(x32, 1, p["n1_w"], p["n1_b"], p["eps"], dt=p["dt"])
    _, t = ops.conv_gemm(xn, p["w1"], p["dw"], bias=p["b1"])
    g = ops.dwconv3x3(t, p["w2"], p["b2"], gate=2)                                   # conv2 + SimpleGate
    w3eff = ops.naf_sca_fold(g, p["w_sca"], p["b_sca"], p["w3"], rowscale=p["beta"])  # x*sca(x) and beta folded
    ops.conv_gemm(g, w3eff, c, bias=p["b3_beta"], res2=x32, out_f32=x32, w_batched=True)
    xn = ops.rownorm(x32, 1, p["n2_w"], p["n2_b"], p["eps"], out=xn)
    _, t = ops.conv_gemm(xn, p["w4"], p["ffn"], bias=p["b4"])
    g = ops.gate_mul(t)
    ops.conv_gemm(g, p["w5"], c, bias=p["b5_gamma"], res2=x32, out_f32=x32)
    return x32


def run_naf_stack(x32, preps):
    for p in preps:
        run_naf_block(x32, p)
    return x32


class _NAFBase(nn.Module):
    def _build_unet(self, img_channel, width, middle_blk_num, enc_blk_nums, dec_blk_nums, fusion=None):
        self.intro = nn.Conv2d(img_channel, width, 3, 1, 1)
        self.ending = nn.Conv2d(width, img_channel, 3, 1, 1)
        # registration order as in the reference (:319-323) so that state_dict() iterates identically
        self.encoders = nn.ModuleList()
        self.decoders = nn.ModuleList()
        self.middle_blks = nn.ModuleList()
        self.ups = nn.ModuleList()
        self.downs = nn.ModuleList()
        chan = width
        for i, num in enumerate(enc_blk_nums):
            self.encoders.append(nn.Sequential(*[NAFBlock(chan) for _ in range(num)]))
            self.downs.append(nn.Conv2d(chan, 2 * chan, 2, 2))
            if fusion is not None:
                self.masa_blk_enc.append(nn.Sequential(*[NAFResFuseBlock(chan * 2) for _ in range(fusion[i])]))
            chan *= 2
        self.middle_blks = nn.Sequential(*[NAFBlock(chan) for _ in range(middle_blk_num)])
        if fusion is not None:
            self.masa_blk_middle.append(
                nn.Sequential(*[NAFResFuseBlock(chan * 2) for _ in range(fusion[len(enc_blk_nums)])]))
        for num in dec_blk_nums:
            self.ups.append(nn.Sequential(nn.Conv2d(chan, chan * 2, 1, bias=False), nn.PixelShuffle(2)))
            chan //= 2
            self.decoders.append(nn.Sequential(*[NAFBlock(chan) for _ in range(num)]))
        self.width = width
        self.img_channel = img_channel
        if len(dec_blk_nums) != len(enc_blk_nums):
            raise ValueError("NAFNet needs as many decoder stages as encoder stages")
        self._prep_cache = None

    def _prep_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _wants_grad(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def prepared(self, train=False):
        """Packed operands, one cache per mode (inference: fp16 operands; train=True: bf16, see operand_dtype)."""
        key = self._prep_key()
        if not isinstance(self._prep_cache, dict):
            self._prep_cache = {}
        c = self._prep_cache.get(bool(train))
        if c is None or c[0] != key:
            with torch.no_grad():
                self._prep_dt = operand_dtype(train)
                c = self._prep_cache[bool(train)] = (key, self._prepare())
        return c[1]

    def _prepare_unet(self):
        dt = getattr(self, "_prep_dt", torch.float16)
        P = dict(encoders=[[_prep_naf(b, dt) for b in st] for st in self.encoders],
                 decoders=[[_prep_naf(b, dt) for b in st] for st in self.decoders],
                 middle=[_prep_naf(b, dt) for b in self.middle_blks], dt=dt)
        P["downs"] = [dict(w=ops.pack_conv_weight(d.weight, dt=dt), b=_f(d.bias), Co=d.out_channels) for d in self.downs]
        P["ups"] = [dict(w=ops.pack_conv_weight(u[0].weight, dt=dt), Co=u[0].out_channels) for u in self.ups]
        P["intro"] = dict(w=_f(self.intro.weight), b=_f(self.intro.bias))
        ew = self.ending.weight
        w8 = torch.zeros(8, ew.shape[1], 3, 3, dtype=ew.dtype, device=ew.device)
        w8[: ew.shape[0]] = ew.detach()
        P["ending"] = dict(w=ops.pack_conv_weight(w8, dt=dt), b=ops.pad_vec(self.ending.bias, 8), Co=ew.shape[0])
        return P

    def _check(self, *ts):
        for t in ts:
            if not t.is_cuda:
                raise TdrError("textualdegremoval_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")
        if self.width % 8:
            raise TdrError("NAFNet (B200): width must be a multiple of 8")

    def _unet(self, P, x32, fuse=None):
        """x32: fp32 NHWC stream of the first stage (possibly the first half of a fusion buffer).
        fuse(i, x32) -> stream after the i-th fusion stage (i == n_enc for the middle one)."""
        dev = x32.device
        dt = P["dt"]
        encs = []
        n_enc = len(P["encoders"])
        for i in range(n_enc):
            if fuse is not None:
                x32 = fuse(i, x32)
            run_naf_stack(x32, P["encoders"][i])
            encs.append(x32)
            B, H, W, c = x32.shape
            nxt = self._alloc_stream(i + 1, B, H // 2, W // 2, 2 * c, dev)
            pd = P["downs"][i]
            ops.conv_gemm(ops.rownorm(x32, 0, dt=dt), pd["w"], pd["Co"], k=2, stride=2, pad=0, bias=pd["b"], out_f32=nxt)
            x32 = nxt
        if fuse is not None:
            x32 = fuse(n_enc, x32)
        run_naf_stack(x32, P["middle"])
        for i, skip in enumerate(encs[::-1]):
            B, H, W, c = x32.shape
            up = torch.empty((B, H * 2, W * 2, c // 2), dtype=F32, device=dev)
            ops.conv_gemm(ops.rownorm(x32, 0, dt=dt), P["ups"][i]["w"], P["ups"][i]["Co"], out_f32=up, res2=skip, store_mode=2)
            x32 = run_naf_stack(up, P["decoders"][i])
        o8, _ = ops.conv_gemm(ops.rownorm(x32, 0, dt=dt), P["ending"]["w"], 8, k=3, pad=1, bias=P["ending"]["b"], want="f32")
        return o8[..., : P["ending"]["Co"]]

    def _alloc_stream(self, stage, B, H, W, c, dev):
        return torch.empty((B, H, W, c), dtype=F32, device=dev)


class NAFNet(NAFTrainMixin, _NAFBase):
    def __init__(self, img_channel=3, width=16, middle_blk_num=1, enc_blk_nums=[], dec_blk_nums=[]):
        super().__init__()
        self._build_unet(img_channel, width, middle_blk_num, enc_blk_nums, dec_blk_nums)
        self.padder_size = 2 ** len(self.encoders)

    def _prepare(self):
        return self._prepare_unet()

    def forward(self, inp):
        """:356-379.  NCHW in/out, zero-padded to a multiple of 2**stages and cropped back.  Differentiable w.r.t. the
        parameters under autograd (nafnet_train / restormer_train.NetFunction)."""
        self._check(inp)
        if self._wants_grad():
            return train_call(self, inp)
        P = self.prepared()
        B, _, H, W = inp.shape
        h, w = ops.round_up(H, self.padder_size), ops.round_up(W, self.padder_size)
        inp32 = ops.nchw_to_nhwc(inp, h, w)
        x = torch.empty((B, h, w, self.width), dtype=F32, device=inp.device)
        ops.conv3x3_small_ci(inp32, P["intro"]["w"], P["intro"]["b"], out_f32=x)
        out = self._unet(P, x)
        return ops.nhwc_to_nchw(out, H, W, res=inp32)


class NAFNetRefFusion(GuidedNAFTrainMixin, MasaTrainMixin, MasaMixin, _NAFBase):
    def __init__(self, img_channel=3, width=16, middle_blk_num=1, enc_blk_nums=[], dec_blk_nums=[], nf=64,
                 ext_n_blocks=[4, 4, 4, 4], reffusion_n_blocks=[1, 1, 1, 1], reffusion_n_blocks_middle=1, scale=1,
                 num_nbr=1, psize=3, lr_block_size=8, ref_down_block_size=1.5, dilations=[1, 2, 3]):
        super().__init__()
        if num_nbr != 1 or psize != 3:
            raise TdrError("NAFNetRefFusion (B200): only num_nbr=1, psize=3 are implemented (all shipped options)")
        if nf != width:
            raise TdrError("NAFNetRefFusion: nf must equal width (warped reference features are concatenated "
                           "channel-for-channel with the U-Net features, :717-727)")
        n_enc = len(enc_blk_nums)
        fusion = list(reffusion_n_blocks)
        if len(fusion) == n_enc:                 # B2: the reference needs n_enc + 1 entries
            fusion.append(fusion[-1])
        self.scale, self.num_nbr, self.psize = scale, num_nbr, psize
        self.lr_block_size, self.ref_down_block_size, self.dilations = lr_block_size, ref_down_block_size, list(dilations)
        self.masa_enc = Encoder(img_channel, nf, ext_n_blocks, levels=n_enc + 1)
        self.masa_blk_enc = nn.ModuleList()
        self.masa_blk_middle = nn.ModuleList()
        self.masa_blk_dec = nn.ModuleList()
        self._build_unet(img_channel, width, middle_blk_num, enc_blk_nums, dec_blk_nums, fusion=fusion)
        self.padder_size = 2 ** len(self.encoders)

    def _prepare(self):
        P = self._prepare_unet()
        P["masa_enc"] = self.prepare_masa_enc()
        dt = P["dt"]
        P["fuse"] = [[_prep_naf(b, dt) for b in st] for st in self.masa_blk_enc] + \
                    [[_prep_naf(b, dt) for b in self.masa_blk_middle[0]]]
        return P

    def forward(self, inp, ref, return_aux=False):
        """:587-740.  NCHW in/out, arbitrary H, W (zero-padded to 2**stages * lr_block_size, cropped back)."""
        self._check(inp, ref)
        if self._wants_grad() and not return_aux:
            return train_call(self, inp, ref)
        P = self.prepared()
        dev = inp.device
        B, _, oh, ow = inp.shape
        mult = self.padder_size * self.lr_block_size
        h, w = ops.round_up(oh, mult), ops.round_up(ow, mult)
        hr, wr = ops.round_up(ref.shape[2], mult), ops.round_up(ref.shape[3], mult)
        E = P["masa_enc"]
        if (h, w) == (hr, wr):           # shared weights: lq and ref run through the encoder as one batch
            both = torch.empty((2 * B, h, w, inp.shape[1]), dtype=F32, device=dev)
            lq32, ref32 = both[:B], both[B:]
            ops.nchw_to_nhwc_into(inp, h, w, dst32=lq32)
            ops.nchw_to_nhwc_into(ref, hr, wr, dst32=ref32)
            fb, d32 = self._masa_encode(E, both)
            f_lq, f_ref, lq_d32, ref_d32 = [t[:B] for t in fb], [t[B:] for t in fb], d32[:B], d32[B:]
        else:
            lq32, ref32 = ops.nchw_to_nhwc(inp, h, w), ops.nchw_to_nhwc(ref, hr, wr)
            (f_lq, lq_d32), (f_ref, ref_d32) = self._masa_encode(E, lq32), self._masa_encode(E, ref32)
        nlev = len(f_ref)
        chans = [self.width * 2 ** i for i in range(nlev)]
        fbuf = [torch.empty((B, h >> i, w >> i, 2 * chans[i]), dtype=F32, device=dev) for i in range(nlev)]
        self._fbuf = fbuf
        aux = self._masa_warp(lq_d32, ref_d32, f_ref, h, w, hr, wr, [fbuf[i][..., chans[i]:] for i in range(nlev)])
        if return_aux:
            aux.update(feat_lq=f_lq, feat_ref=f_ref, deep32_lq=lq_d32, deep32_ref=ref_d32, deep_scale=self._masa_last_scale[-1], warps=[fbuf[i][..., chans[i]:].clone() for i in range(nlev)])
        ops.conv3x3_small_ci(lq32, P["intro"]["w"], P["intro"]["b"], out_f32=fbuf[0][..., :chans[0]])

        def fuse(i, x32):
            run_naf_stack(fbuf[i], P["fuse"][i])         # NAFResFuseBlock on 2C channels, keep the first C (:719)
            return fbuf[i][..., :chans[i]]

        out = self._unet(P, fbuf[0][..., :chans[0]], fuse=fuse)
        self._fbuf = None
        out = ops.nhwc_to_nchw(out, oh, ow, res=lq32)
        return (out, aux) if return_aux else out

    def _alloc_stream(self, stage, B, H, W, c, dev):
        # the stream entering fusion stage `stage` lives in the first half of that stage's fusion buffer
        return self._fbuf[stage][..., :c]
