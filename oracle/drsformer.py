"""TEST INFRASTRUCTURE ONLY -- fp32 CPU restatement of the reference's DRSformer-guided networks
(/root/reference/models/archs/network_drsformer_guided_arch.py ``DRSformerRefFusion`` :679-1123 and
network_drsformer_guided_arch_200L_SPA.py ``DRSformer200L_SPA_RefFusion``), functional over a ``state_dict``.  MASA guidance
and U-Net wiring are shared with ``oracle/restormer.py`` (textually identical code in the reference files); this file
restates the sparse transformer block (TKSA :260-332, MSFN :216-256) and the MEFC ``subnet`` (:371-549).  Pinned to the
unmodified reference modules by ``tests/golden/guided_drsformer*_128.npz``.  Never imported by the product.
"""
import torch
import torch.nn.functional as F

from .restormer import _conv, _ln, downsample, masa_encoder, masa_warp, pad_to, upsample


def tksa(sd, p, x, heads):
    b, c, h, w = x.shape
    qkv = _conv(sd, p + ".qkv", x)
    qkv = _conv(sd, p + ".qkv_dwconv", qkv, padding=1, groups=qkv.shape[1])
    q, k, v = [t.reshape(b, heads, c // heads, h * w) for t in qkv.chunk(3, dim=1)]
    q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    attn = (q @ k.transpose(-2, -1)) * sd[p + ".temperature"]
    C = c // heads
    out = 0
    for i, kk in enumerate((int(C / 2), int(C * 2 / 3), int(C * 3 / 4), int(C * 4 / 5))):
        idx = torch.topk(attn, k=kk, dim=-1, largest=True)[1]
        mask = torch.zeros_like(attn).scatter_(-1, idx, 1.)
        a = torch.where(mask > 0, attn, torch.full_like(attn, float("-inf"))).softmax(dim=-1)
        out = out + (a @ v) * sd[f"{p}.attn{i + 1}"]
    return _conv(sd, p + ".project_out", out.reshape(b, c, h, w))


def msfn(sd, p, x):
    x = _conv(sd, p + ".project_in", x)
    a3 = F.relu(_conv(sd, p + ".dwconv3x3", x, padding=1, groups=x.shape[1]))
    a5 = F.relu(_conv(sd, p + ".dwconv5x5", x, padding=2, groups=x.shape[1]))
    x1_3, x2_3 = a3.chunk(2, dim=1)
    x1_5, x2_5 = a5.chunk(2, dim=1)
    x1, x2 = torch.cat([x1_3, x1_5], 1), torch.cat([x2_3, x2_5], 1)
    g = x1.shape[1] // 2
    y1 = F.relu(_conv(sd, p + ".dwconv3x3_1", x1, padding=1, groups=g))
    y2 = F.relu(_conv(sd, p + ".dwconv5x5_1", x2, padding=2, groups=g))
    return _conv(sd, p + ".project_out", torch.cat([y1, y2], 1))


def block(sd, p, x, heads):
    x = x + tksa(sd, p + ".attn", _ln(sd, p + ".norm1", x), heads)
    return x + msfn(sd, p + ".ffn", _ln(sd, p + ".norm2", x))


def fusion_block(sd, p, x, heads):
    return block(sd, p, x, heads) * sd[p + ".alpha"] + x


def _stack(sd, p, x, heads, fn=block):
    i = 0
    while f"{p}.{i}.norm1.body.weight" in sd:
        x = fn(sd, f"{p}.{i}", x, heads)
        i += 1
    return x


def _dwk(sd, key, x, k, dil=1):
    return F.conv2d(x, sd[key], padding=dil * (k - 1) // 2, dilation=dil, groups=x.shape[1])


def subnet(sd, p, x):
    """MEFC (:522-549): layers.0 = OALayer, layers.1 = GroupOLs."""
    b = x.shape[0]
    y = x.mean(dim=(-2, -1))
    y = F.linear(F.relu(F.linear(y, sd[p + ".layers.0.ca_fc.0.weight"], sd[p + ".layers.0.ca_fc.0.bias"])),
                 sd[p + ".layers.0.ca_fc.2.weight"], sd[p + ".layers.0.ca_fc.2.bias"])
    weights = F.softmax(y.view(b, 4, 8), dim=-1)
    g = p + ".layers.1"
    s0 = F.relu(F.conv2d(x, sd[g + ".preprocess.op.0.weight"]))
    for i in range(4):
        o = f"{g}._ops.{i}"
        states = []
        for k, ks in enumerate((1, 3, 5, 7)):
            e = f"{o}._ops.{k}.op"
            t = F.conv2d(_dwk(sd, e + ".0.weight", s0, ks), sd[e + ".1.weight"])
            t = F.conv2d(_dwk(sd, e + ".3.weight", F.relu(t), ks), sd[e + ".4.weight"])
            states.append(t)
        for k, ks in enumerate((3, 5, 7)):
            e = f"{o}._ops.{4 + k}.op"
            states.append(F.conv2d(_dwk(sd, e + ".0.weight", s0, ks, dil=2), sd[e + ".1.weight"]))
        states.append(F.avg_pool2d(s0, 3, stride=1, padding=1, count_include_pad=False))
        states = [s * weights[:, i, k].view(-1, 1, 1, 1) for k, s in enumerate(states)]
        s0 = F.relu(F.relu(F.conv2d(torch.cat(states, 1), sd[o + "._out.0.weight"])) + s0)
    return s0


def drsformer_ref_fusion_forward(sd, inp_img, ref_img, heads=(1, 2, 4, 8), lr_block_size=8, ref_down_block_size=1.5,
                                 dilations=(1, 2, 3)):
    """Both variants: the MEFC stages run when their parameters are in the state_dict."""
    padder = 8
    _, _, oh, ow = inp_img.shape
    inp_img = pad_to(inp_img, padder * lr_block_size)
    ref_img = pad_to(ref_img, padder * lr_block_size)
    _, _, h, w = inp_img.shape
    _, _, hr, wr = ref_img.shape
    f_lq, f_ref = masa_encoder(sd, inp_img), masa_encoder(sd, ref_img)
    warps = masa_warp(f_lq[-1], f_ref, padder, lr_block_size, ref_down_block_size, dilations, h, w, hr, wr)
    mefc = "encoder_level0.layers.0.ca_fc.0.weight" in sd

    def fuse(x, warp, name, hd):
        return _stack(sd, name, torch.cat([x, warp], 1), hd, fn=fusion_block)[:, : x.shape[1]]

    x1 = _conv(sd, "patch_embed.proj", inp_img, padding=1)
    if mefc:
        x1 = fuse(subnet(sd, "encoder_level0", x1), warps[0], "masa_blk_enc_level1", heads[0])
    # 200L_SPA variant (:968-975 of its file): the level-1 fusion result is assigned to ``inp_enc_level0`` and never used --
    # ``encoder_level1`` consumes the patch embedding itself
    e1 = _stack(sd, "encoder_level1", x1, heads[0])
    e2 = _stack(sd, "encoder_level2", fuse(downsample(sd, "down1_2", e1), warps[1], "masa_blk_enc_level2", heads[1]), heads[1])
    e3 = _stack(sd, "encoder_level3", fuse(downsample(sd, "down2_3", e2), warps[2], "masa_blk_enc_level3", heads[2]), heads[2])
    lat = _stack(sd, "latent", fuse(downsample(sd, "down3_4", e3), warps[3], "masa_blk_enc_level4", heads[3]), heads[3])
    d3 = _stack(sd, "decoder_level3", _conv(sd, "reduce_chan_level3", torch.cat([upsample(sd, "up4_3", lat), e3], 1)), heads[2])
    d2 = _stack(sd, "decoder_level2", _conv(sd, "reduce_chan_level2", torch.cat([upsample(sd, "up3_2", d3), e2], 1)), heads[1])
    d1 = _stack(sd, "decoder_level1", torch.cat([upsample(sd, "up2_1", d2), e1], 1), heads[0])
    if mefc:
        d1 = subnet(sd, "refinement", d1)
    return (_conv(sd, "output", d1, padding=1) + inp_img)[:, :, :oh, :ow]
