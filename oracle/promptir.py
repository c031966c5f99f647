"""TEST INFRASTRUCTURE ONLY -- fp32 CPU restatement of the reference's ``PromptIRRefFusion``
(/root/reference/models/archs/network_promptir_guided_arch.py:593-1092, ``decoder=True``), functional over a
``state_dict``.  The encoder half (MASA guidance, Res-fusion blocks, encoder levels :913-1044) is the guided Restormer's
(``oracle/restormer.py``, the classes are textually identical in the two reference files); this file adds the prompt
stages of the decoder (:1046-1091) and ``PromptGenBlock`` (:417-440).  Pinned to the unmodified reference module by
``tests/golden/guided_promptir_128.npz`` (``oracle/make_golden.py promptir``).  Never imported by the product.
"""
import torch
import torch.nn.functional as F

from .restormer import (_conv, _stack, downsample, masa_encoder, masa_warp, pad_to, res_fusion_block, transformer_block,
                        upsample)


def prompt_gen(sd, p, x):
    """PromptGenBlock.forward (:424-440)."""
    b, c, h, w = x.shape
    emb = x.mean(dim=(-2, -1))
    wts = F.softmax(F.linear(emb, sd[p + ".linear_layer.weight"], sd[p + ".linear_layer.bias"]), dim=1)
    prompt = (wts.view(b, -1, 1, 1, 1) * sd[p + ".prompt_param"]).sum(1)          # [B, D, S, S]
    prompt = F.interpolate(prompt, (h, w), mode="bilinear")
    return F.conv2d(prompt, sd[p + ".conv3x3.weight"], padding=1)


def promptir_ref_fusion_forward(sd, inp_img, ref_img, heads=(1, 2, 4, 8), lr_block_size=8, ref_down_block_size=1.5,
                                dilations=(1, 2, 3)):
    padder = 8
    _, _, oh, ow = inp_img.shape
    inp_img = pad_to(inp_img, padder * lr_block_size)
    ref_img = pad_to(ref_img, padder * lr_block_size)
    _, _, h, w = inp_img.shape
    _, _, hr, wr = ref_img.shape
    f_lq = masa_encoder(sd, inp_img)
    f_ref = masa_encoder(sd, ref_img)
    warps = masa_warp(f_lq[-1], f_ref, padder, lr_block_size, ref_down_block_size, dilations, h, w, hr, wr)

    def fuse(x, warp, name, hd):
        cat = torch.cat([x, warp], 1)
        return _stack(sd, name, cat, hd, fn=res_fusion_block)[:, : x.shape[1]]

    def prompt_stage(x, i):                       # :1049-1053 and the two stages after it; heads[2] for all three
        x = torch.cat([x, prompt_gen(sd, f"prompt{i}", x)], 1)
        x = transformer_block(sd, f"noise_level{i}", x, heads[2])
        return _conv(sd, f"reduce_noise_level{i}", x)

    x1 = fuse(_conv(sd, "patch_embed.proj", inp_img, padding=1), warps[0], "masa_blk_enc_level1", heads[0])
    e1 = _stack(sd, "encoder_level1", x1, heads[0])
    e2 = _stack(sd, "encoder_level2", fuse(downsample(sd, "down1_2", e1), warps[1], "masa_blk_enc_level2", heads[1]), heads[1])
    e3 = _stack(sd, "encoder_level3", fuse(downsample(sd, "down2_3", e2), warps[2], "masa_blk_enc_level3", heads[2]), heads[2])
    lat = _stack(sd, "latent", fuse(downsample(sd, "down3_4", e3), warps[3], "masa_blk_enc_level4", heads[3]), heads[3])
    lat = prompt_stage(lat, 3)
    d3 = _conv(sd, "reduce_chan_level3", torch.cat([upsample(sd, "up4_3", lat), e3], 1))
    d3 = prompt_stage(_stack(sd, "decoder_level3", d3, heads[2]), 2)
    d2 = _conv(sd, "reduce_chan_level2", torch.cat([upsample(sd, "up3_2", d3), e2], 1))
    d2 = prompt_stage(_stack(sd, "decoder_level2", d2, heads[1]), 1)
    d1 = torch.cat([upsample(sd, "up2_1", d2), e1], 1)
    d1 = _stack(sd, "refinement", _stack(sd, "decoder_level1", d1, heads[0]), heads[0])
    return (_conv(sd, "output", d1, padding=1) + inp_img)[:, :, :oh, :ow]
