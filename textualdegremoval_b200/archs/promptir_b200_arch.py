"""B200-native ``PromptIRRefFusion`` (inference).

Drop-in for the class of the same name in the reference's ``models/archs/network_promptir_guided_arch.py`` (:593-1092;
option ``options/train_restoration/001_promptir_all_in_one_restoration.yml``): same constructor kwargs, same
``state_dict`` keys and shapes, ``net(lq, ref)`` on NCHW tensors.  The encoder half (MASA guidance, Res-fusion blocks,
four encoder levels) is the guided Restormer's; the decoder adds the three prompt stages (:1049-1075):

    prompt = PromptGenBlock(x)            spatial mean -> linear -> softmax -> weighted prompt sum -> bilinear -> conv3x3
    x      = reduce_noise(noise_block(cat[x, prompt]))

Two facts about the reference that this module reproduces rather than repairs:

  * with ``decoder=False`` (what the shipped option file sets) the reference's forward cannot run: ``up4_3`` is
    ``Upsample(dim * 4)`` but is applied to the ``dim * 8`` latent when the prompt stage that reduces it is skipped.  The
    same ``RuntimeError`` is raised here, before any kernel is launched;
  * the MASA encoder index shift is the guided Restormer's.

The prompt-interaction blocks have 176 / 80 / 40 channels per attention head: the 176-wide Gram takes the generic (SIMT)
path of ``tdr_mdta_gram`` -- it runs on 1/8-scale features only.  Training (explicit backward) is not implemented for
this family: calling it with parameters that require grad under grad mode raises ``TdrError``.
"""
import torch
import torch.nn as nn

from .. import ops
from ..lib import TdrError
from .masa import Encoder, MasaMixin, prep_conv as _prep_conv, _f
from .restormer_b200_arch import (Downsample, OverlapPatchEmbed, RestormerRefFusion, TransformerBlock,
                                  TransformerResFusionBlock, Upsample, _blocks, _prep_block, operand_dtype, run_stack)

F32 = torch.float32


class PromptGenBlock(nn.Module):
    """:417-440 (parameter holder)."""

    def __init__(self, prompt_dim=128, prompt_len=5, prompt_size=96, lin_dim=192):
        super().__init__()
        self.prompt_param = nn.Parameter(torch.rand(1, prompt_len, prompt_dim, prompt_size, prompt_size))
        self.linear_layer = nn.Linear(lin_dim, prompt_len)
        self.conv3x3 = nn.Conv2d(prompt_dim, prompt_dim, kernel_size=3, stride=1, padding=1, bias=False)


class PromptIRRefFusion(MasaMixin, nn.Module):
    _guided_encode = RestormerRefFusion._guided_encode
    _down = RestormerRefFusion._down
    _check = RestormerRefFusion._check
    _prep_key = RestormerRefFusion._prep_key
    prepared = RestormerRefFusion.prepared
    _run_stack = staticmethod(run_stack)
    _after_patch_embed = RestormerRefFusion._after_patch_embed
    dual_pixel_task = False

    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4,
                 heads=[1, 2, 4, 8], ffn_expansion_factor=2.66, bias=False, LayerNorm_type="WithBias", decoder=False,
                 nf=64, ext_n_blocks=[4, 4, 4, 4], reffusion_n_blocks=[1, 1, 1, 1], reffusion_n_blocks_middle=1, scale=1,
                 num_nbr=1, psize=3, lr_block_size=8, ref_down_block_size=1.5, dilations=[1, 2, 3]):
        super().__init__()
        if num_nbr != 1 or psize != 3:
            raise TdrError("PromptIRRefFusion (B200): only num_nbr=1, psize=3 are implemented (all shipped options)")
        if not 1 <= len(dilations) <= 3:
            raise TdrError("PromptIRRefFusion (B200): 1..3 dilations supported")
        if nf != dim:
            raise TdrError("PromptIRRefFusion: nf must equal dim (warped reference features are concatenated "
                           "channel-for-channel with the U-Net features)")
        self.scale, self.num_nbr, self.psize = scale, num_nbr, psize
        self.lr_block_size, self.ref_down_block_size, self.dilations = lr_block_size, ref_down_block_size, list(dilations)
        self.padder_size = 2 ** 3
        self.decoder = decoder
        kw = dict(ffn_expansion_factor=ffn_expansion_factor, bias=bias, LayerNorm_type=LayerNorm_type)
        d = self.dims = [dim, dim * 2, dim * 4, dim * 8]
        self.masa_enc = Encoder(inp_channels, nf, ext_n_blocks, levels=4)
        self.masa_blk_enc, self.masa_blk_middle, self.masa_blk_dec = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.patch_embed = OverlapPatchEmbed(inp_channels, dim)
        if decoder:                                                         # :645-648
            self.prompt1 = PromptGenBlock(prompt_dim=64, prompt_len=5, prompt_size=64, lin_dim=96)
            self.prompt2 = PromptGenBlock(prompt_dim=128, prompt_len=5, prompt_size=32, lin_dim=192)
            self.prompt3 = PromptGenBlock(prompt_dim=320, prompt_len=5, prompt_size=16, lin_dim=384)
        # chnl_reduce* / reduce_noise_channel_* are constructed by the reference and never called by its forward
        # (:650-654, :672, :694): kept for state_dict parity.  Registration order = the reference's (key order).
        self.chnl_reduce1 = nn.Conv2d(64, 64, 1, bias=bias)
        self.chnl_reduce2 = nn.Conv2d(128, 128, 1, bias=bias)
        self.chnl_reduce3 = nn.Conv2d(320, 256, 1, bias=bias)
        names = ["encoder_level1", "encoder_level2", "encoder_level3", "latent"]
        downs = ["down1_2", "down2_3", "down3_4", None]
        extra = [64, 128, 256, None]
        for i in range(4):
            if extra[i]:
                setattr(self, f"reduce_noise_channel_{i + 1}", nn.Conv2d(d[i] + extra[i], d[i], 1, bias=bias))
            setattr(self, f"masa_blk_enc_level{i + 1}",
                    _blocks(reffusion_n_blocks[i], TransformerResFusionBlock, dim=2 * d[i], num_heads=heads[i], **kw))
            setattr(self, names[i], _blocks(num_blocks[i], TransformerBlock, dim=d[i], num_heads=heads[i], **kw))
            if downs[i]:
                setattr(self, downs[i], Downsample(d[i]))
        self.up4_3 = Upsample(d[2])                                         # :745 (dim*4, not dim*8)
        self.reduce_chan_level3 = nn.Conv2d(d[1] + 192, d[2], 1, bias=bias)
        self.noise_level3 = TransformerBlock(dim=d[2] + 512, num_heads=heads[2], **kw)
        self.reduce_noise_level3 = nn.Conv2d(d[2] + 512, d[2], 1, bias=bias)
        self.decoder_level3 = _blocks(num_blocks[2], TransformerBlock, dim=d[2], num_heads=heads[2], **kw)
        self.up3_2 = Upsample(d[2])
        self.reduce_chan_level2 = nn.Conv2d(d[2], d[1], 1, bias=bias)
        self.noise_level2 = TransformerBlock(dim=d[1] + 224, num_heads=heads[2], **kw)
        self.reduce_noise_level2 = nn.Conv2d(d[1] + 224, d[2], 1, bias=bias)
        self.decoder_level2 = _blocks(num_blocks[1], TransformerBlock, dim=d[1], num_heads=heads[1], **kw)
        self.up2_1 = Upsample(d[1])
        self.noise_level1 = TransformerBlock(dim=d[1] + 64, num_heads=heads[2], **kw)
        self.reduce_noise_level1 = nn.Conv2d(d[1] + 64, d[1], 1, bias=bias)
        self.decoder_level1 = _blocks(num_blocks[0], TransformerBlock, dim=d[1], num_heads=heads[0], **kw)
        self.refinement = _blocks(num_refinement_blocks, TransformerBlock, dim=d[1], num_heads=heads[0], **kw)
        self.output = nn.Conv2d(d[1], out_channels, 3, 1, 1, bias=bias)
        self.nf = nf
        self._prep_cache = None

    # ---- weight cache -------------------------------------------------------------------------
    def _prepare(self):
        if getattr(self, "_prep_train_flag", False):
            raise TdrError("PromptIRRefFusion (B200): inference only -- the explicit backward is not implemented")
        P = {}
        dt = P["dt"] = operand_dtype(False)
        stacks = ["encoder_level1", "encoder_level2", "encoder_level3", "latent", "decoder_level3", "decoder_level2",
                  "decoder_level1", "refinement"] + [f"masa_blk_enc_level{i}" for i in range(1, 5)]
        for name in stacks:
            P[name] = [_prep_block(b) for b in getattr(self, name)]
        for name in ["down1_2", "down2_3", "down3_4", "up4_3", "up3_2", "up2_1"]:
            P[name] = _prep_conv(getattr(self, name).body[0], dt)
        for name in ["reduce_chan_level3", "reduce_chan_level2"]:
            P[name] = _prep_conv(getattr(self, name), dt)
        if self.decoder:
            for i in (1, 2, 3):
                P[f"noise_level{i}"] = [_prep_block(getattr(self, f"noise_level{i}"))]
                P[f"reduce_noise_level{i}"] = _prep_conv(getattr(self, f"reduce_noise_level{i}"), dt)
                g = getattr(self, f"prompt{i}")
                pc = _prep_conv(g.conv3x3, dt)
                pc.update(param=_f(g.prompt_param)[0].contiguous(), lin_w=_f(g.linear_layer.weight),
                          lin_b=_f(g.linear_layer.bias), D=g.conv3x3.out_channels)
                P[f"prompt{i}"] = pc
        P["patch_embed"] = dict(w=_f(self.patch_embed.proj.weight), b=_f(self.patch_embed.proj.bias))
        ow = self.output.weight
        co = ow.shape[0]
        w8 = torch.zeros(8, ow.shape[1], 3, 3, dtype=ow.dtype, device=ow.device)
        w8[:co] = ow.detach()
        P["output"] = dict(w=ops.pack_conv_weight(w8, dt=dt), b=ops.pad_vec(self.output.bias, 8), Co=co)
        P["masa_enc"] = self.prepare_masa_enc()
        return P

    # ---- schedules ----------------------------------------------------------------------------
    @staticmethod
    def _prompt(pp, x32, dst32, dt):
        """PromptGenBlock.forward (:424-440) on the fp32 NHWC view x32; the conv3x3 output lands in dst32 (a channel
        slice of the [x || prompt] buffer)."""
        B, H, W, C = x32.shape
        emb = torch.empty((B, C), dtype=F32, device=x32.device)
        tok = x32.as_strided((B, 1, H * W, C), (x32.stride(0), x32.stride(0), x32.stride(2), 1))
        ops.mean_tokens(tok, 0, H * W, emb)
        wts = ops.prompt_weights(emb, pp["lin_w"], pp["lin_b"])
        p16 = torch.empty((B, H, W, pp["D"]), dtype=dt, device=x32.device)
        ops.prompt_mix_resize(pp["param"], wts, H, W, p16)
        ops.conv_gemm(p16, pp["w"], pp["D"], k=3, pad=1, out_f32=dst32)

    def _prompt_stage(self, P, i, buf, C, dt):
        """x = reduce_noise_level_i(noise_level_i(cat[x, prompt_i(x)])) (:1049-1053): buf is the fp32 [x || prompt] buffer
        whose first C channels hold x.  Returns a dense fp32 tensor."""
        pp = P[f"prompt{i}"]
        self._prompt(pp, buf[..., :C], buf[..., C:C + pp["D"]], dt)
        cat = buf[..., :C + pp["D"]]
        run_stack(cat, P[f"noise_level{i}"])
        red = P[f"reduce_noise_level{i}"]
        y32, _ = ops.conv_gemm(ops.rownorm(cat, 0, dt=dt), red["w"], red["Co"], bias=red["b"], want="f32")
        return y32

    def _up_cat_reduce(self, P, x32, enc, up, red, width, dt):
        """Upsample (conv3x3 + PixelShuffle) || encoder skip -> 1x1 reduce, written into the first channels of a new
        fp32 buffer that is ``width`` channels wide (room for the prompt of the next stage)."""
        b, hh, ww, _ = x32.shape
        cu, ce = P[up]["Co"] // 4, enc.shape[-1]
        cat16 = torch.empty((b, hh * 2, ww * 2, cu + ce), dtype=dt, device=x32.device)
        ops.conv_gemm(ops.rownorm(x32, 0, dt=dt), P[up]["w"], P[up]["Co"], k=3, pad=1, out_bf16=cat16[..., :cu],
                      store_mode=2)
        ops.copy_rows(enc, dst16=cat16[..., cu:])
        co = P[red]["Co"]
        buf = torch.empty((b, hh * 2, ww * 2, width), dtype=F32, device=x32.device)
        ops.conv_gemm(cat16, P[red]["w"], co, bias=P[red]["b"], out_f32=buf[..., :co])
        return buf

    def forward(self, inp_img, ref_img, noise_emb=None):
        """:913-1092.  NCHW in, NCHW out, arbitrary H, W (zero-padded to x64, cropped)."""
        self._check(inp_img, ref_img)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise TdrError("PromptIRRefFusion (B200): inference only -- call it under torch.no_grad()")
        d = self.dims
        if not self.decoder:
            raise RuntimeError(f"Given groups=1, weight of size [{2 * d[2]}, {d[2]}, 3, 3], expected input to have "
                               f"{d[2]} channels, but got {d[3]} channels instead (PromptIRRefFusion with decoder=False: "
                               "up4_3 = Upsample(dim*4) is applied to the dim*8 latent; the reference fails the same way)")
        P = self.prepared()
        dt = P["dt"]
        xs, _, lq32, (oh, ow), _, fbuf = self._guided_encode(P, inp_img, ref_img)
        e1, e2, e3, lat = xs
        # level 4: the latent lives in the first dim*8 channels of its (2 * dim*8 wide) fusion buffer: the prompt takes
        # the place of the warped reference features, which the fusion blocks have consumed
        fb3 = fbuf[3]
        if fb3.shape[-1] < d[3] + P["prompt3"]["D"]:
            raise TdrError("PromptIRRefFusion (B200): prompt3 does not fit next to the latent")
        x = self._prompt_stage(P, 3, fb3, d[3], dt)                                             # -> dim*4 @ 1/8
        buf3 = self._up_cat_reduce(P, x, e3, "up4_3", "reduce_chan_level3", d[2] + P["prompt2"]["D"], dt)
        run_stack(buf3[..., :d[2]], P["decoder_level3"])
        x = self._prompt_stage(P, 2, buf3, d[2], dt)                                            # -> dim*4 @ 1/4
        buf2 = self._up_cat_reduce(P, x, e2, "up3_2", "reduce_chan_level2", d[1] + P["prompt1"]["D"], dt)
        run_stack(buf2[..., :d[1]], P["decoder_level2"])
        x = self._prompt_stage(P, 1, buf2, d[1], dt)                                            # -> dim*2 @ 1/2
        b, hh, ww, _ = x.shape
        cu = P["up2_1"]["Co"] // 4
        d1 = torch.empty((b, hh * 2, ww * 2, cu + e1.shape[-1]), dtype=F32, device=x.device)
        ops.conv_gemm(ops.rownorm(x, 0, dt=dt), P["up2_1"]["w"], P["up2_1"]["Co"], k=3, pad=1, out_f32=d1[..., :cu],
                      store_mode=2)
        ops.copy_rows(e1, dst32=d1[..., cu:])
        tail = []
        run_stack(d1, P["decoder_level1"], nxt=P["refinement"][0] if P["refinement"] else None, tail=tail)
        run_stack(d1, P["refinement"], xn=tail[0])
        o8, _ = ops.conv_gemm(ops.rownorm(d1, 0, dt=dt), P["output"]["w"], 8, k=3, pad=1, bias=P["output"]["b"], want="f32")
        return ops.nhwc_to_nchw(o8[..., :P["output"]["Co"]], oh, ow, res=lq32)                  # + inp_img (:1089)
