"""TEST INFRASTRUCTURE ONLY -- fp32 CPU restatement of Restormer / RestormerRefFusion.

Functional over a ``state_dict`` (reference key names).  The MASA
guidance path is restated in closed form (direct gathers, no unfold/fold), i.e. NOT a
transliteration of the reference; ``tests/test_oracle_vs_reference.py`` and the fixtures
written by ``oracle/make_golden.py`` pin it against the unmodified reference modules.

Reference: /root/reference/models/archs/network_restormer_guided_arch.py
  LayerNorm :172-218   FeedForward :223-241   Attention :246-277   TransformerBlock :318-331
  TransformerResFusionBlock :334-353   Down/Upsample :372-391   Restormer.forward :463-501
  Encoder :100-134   ResidualBlock :34-49   search :674-696   search_org :654-672
  transfer :698-715   window placement :793-815   RestormerRefFusion.forward :747-964
"""
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- blocks
def layernorm_c(x, weight, bias=None, eps=1e-5):
    """Per-pixel LayerNorm over the channel dim of NCHW (:172-218).

    BiasFree (bias is None) divides x itself (mean NOT subtracted) by sqrt(var+eps) (:184-186).
    """
    mu = x.mean(dim=1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=1, keepdim=True)
    w = weight.view(1, -1, 1, 1)
    if bias is None:
        return x / torch.sqrt(var + eps) * w
    return (x - mu) / torch.sqrt(var + eps) * w + bias.view(1, -1, 1, 1)


def _ln(sd, p, x):
    return layernorm_c(x, sd[p + ".body.weight"], sd.get(p + ".body.bias"))


def _conv(sd, p, x, **kw):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), **kw)


def mdta(sd, p, x, heads):
    """Multi-DConv head transposed attention (:246-277)."""
    b, c, h, w = x.shape
    qkv = _conv(sd, p + ".qkv", x)
    qkv = _conv(sd, p + ".qkv_dwconv", qkv, padding=1, groups=3 * c)
    q, k, v = qkv.view(b, 3, heads, c // heads, h * w).unbind(1)
    q = q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    k = k / k.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    attn = torch.softmax((q @ k.transpose(-2, -1)) * sd[p + ".temperature"].view(1, heads, 1, 1), dim=-1)
    out = (attn @ v).reshape(b, c, h, w)
    return _conv(sd, p + ".project_out", out)


def gdfn(sd, p, x):
    """Gated-Dconv feed-forward (:223-241); exact (erf) GELU."""
    y = _conv(sd, p + ".project_in", x)
    y = _conv(sd, p + ".dwconv", y, padding=1, groups=y.shape[1])
    y1, y2 = y.chunk(2, dim=1)
    return _conv(sd, p + ".project_out", F.gelu(y1) * y2)


def transformer_block(sd, p, x, heads):
    x = x + mdta(sd, p + ".attn", _ln(sd, p + ".norm1", x), heads)
    return x + gdfn(sd, p + ".ffn", _ln(sd, p + ".norm2", x))


def res_fusion_block(sd, p, x, heads):
    """TransformerResFusionBlock (:334-353): block(x) * alpha + x."""
    return transformer_block(sd, p, x, heads) * sd[p + ".alpha"] + x


def _stack(sd, p, x, heads, fn=transformer_block):
    i = 0
    while f"{p}.{i}.norm1.body.weight" in sd:
        x = fn(sd, f"{p}.{i}", x, heads)
        i += 1
    return x


def downsample(sd, p, x):
    return F.pixel_unshuffle(F.conv2d(x, sd[p + ".body.0.weight"], padding=1), 2)


def upsample(sd, p, x):
    return F.pixel_shuffle(F.conv2d(x, sd[p + ".body.0.weight"], padding=1), 2)


def _decoder(sd, heads, latent, enc1, enc2, enc3, inp_img, enc_in1=None):
    d3 = torch.cat([upsample(sd, "up4_3", latent), enc3], 1)
    d3 = _stack(sd, "decoder_level3", _conv(sd, "reduce_chan_level3", d3), heads[2])
    d2 = torch.cat([upsample(sd, "up3_2", d3), enc2], 1)
    d2 = _stack(sd, "decoder_level2", _conv(sd, "reduce_chan_level2", d2), heads[1])
    d1 = torch.cat([upsample(sd, "up2_1", d2), enc1], 1)
    d1 = _stack(sd, "decoder_level1", d1, heads[0])
    d1 = _stack(sd, "refinement", d1, heads[0])
    if "skip_conv.weight" in sd:                    # dual_pixel_task (:494-496)
        return _conv(sd, "output", d1 + _conv(sd, "skip_conv", enc_in1), padding=1)
    return _conv(sd, "output", d1, padding=1) + inp_img


def restormer_forward(sd, inp_img, heads=(1, 2, 4, 8)):
    """Restormer.forward (:463-501)."""
    x1 = _conv(sd, "patch_embed.proj", inp_img, padding=1)
    e1 = _stack(sd, "encoder_level1", x1, heads[0])
    e2 = _stack(sd, "encoder_level2", downsample(sd, "down1_2", e1), heads[1])
    e3 = _stack(sd, "encoder_level3", downsample(sd, "down2_3", e2), heads[2])
    lat = _stack(sd, "latent", downsample(sd, "down3_4", e3), heads[3])
    return _decoder(sd, heads, lat, e1, e2, e3, inp_img, x1)


# ----------------------------------------------------------------------------- MASA
def masa_encoder(sd, x, p="masa_enc"):
    """Encoder (:100-134): conv(+bias)+ReLU then residual blocks, stride-2 conv between levels."""
    feats = []
    lvl = 1
    while f"{p}.conv_L{lvl}.weight" in sd:
        x = F.relu(_conv(sd, f"{p}.conv_L{lvl}", x, stride=1 if lvl == 1 else 2, padding=1))
        i = 0
        while f"{p}.blk_L{lvl}.{i}.conv1.weight" in sd:
            q = f"{p}.blk_L{lvl}.{i}"
            x = _conv(sd, q + ".conv2", F.relu(_conv(sd, q + ".conv1", x, padding=1)), padding=1) + x
            i += 1
        feats.append(x)
        lvl += 1
    return feats


def _l2n(t, dim):
    return t / t.norm(dim=dim, keepdim=True).clamp_min(1e-12)


def lq_blocks(f_lq, k_y, k_x):
    """[N,C,H,W] -> [N, py*px, C, k_y+2, k_x+2] tiles with a replicated 1-px halo (:785-787)."""
    n, c, h, w = f_lq.shape
    py, px = h // k_y, w // k_x
    fp = F.pad(f_lq, (1, 1, 1, 1), mode="replicate")
    rows = []
    for by in range(py):
        for bx in range(px):
            rows.append(fp[:, :, by * k_y: by * k_y + k_y + 2, bx * k_x: bx * k_x + k_x + 2])
    return torch.stack(rows, 1)


def coarse_search(blocks, f_ref, dilations):
    """search (:674-696): sum over dilations of cosine(block centre 3x3(dil), ref 3x3(dil)); returns
    (score [N, p2, Hr*Wr], argmax [N, p2])."""
    n, p2, c, ky, kx = blocks.shape
    _, _, hr, wr = f_ref.shape
    cy, cx = ky // 2, kx // 2
    score = 0
    for d in dilations:
        u = blocks[:, :, :, cy - d: cy + d + 1: d, cx - d: cx + d + 1: d].reshape(n, p2, c * 9)
        u = _l2n(u, 2)
        fp = F.pad(f_ref, (d, d, d, d))
        v = torch.stack([fp[:, :, ty * d: ty * d + hr, tx * d: tx * d + wr]
                         for ty in range(3) for tx in range(3)], 2)          # [N,C,9,Hr,Wr]
        v = _l2n(v.reshape(n, c * 9, hr * wr), 1)
        score = score + u @ v
    return score, score.argmax(-1)


def window_origin(idx, wr, hr, dx, dy):
    """Window placement (:793-815).  idx [N,p2] -> (y1, x1) top-left of the (d+2)^2 window."""
    ix, iy = idx % wr, idx // wr

    def place(i, d, lim):
        a = i - d // 2 - 1
        b = i + d // 2 + 1
        neg = a < 0
        a = torch.where(neg, torch.zeros_like(a), a)
        b = torch.where(neg, torch.full_like(b, d + 1), b)
        over = b > lim - 1
        b = torch.where(over, torch.full_like(b, lim - 1), b)
        a = torch.where(over, b - (d + 1), a)
        return a

    return place(iy, dy, hr), place(ix, dx, wr)


def crop_windows(f, y1, x1, size_y, size_x, s):
    """W_s(b)[c,u,v] = f[n, c, y1*s+u, x1*s+v]  (:717-734, :835-852).  -> [N*p2, C, size_y*s, size_x*s]"""
    n, c = f.shape[:2]
    out = []
    y1l, x1l = y1.tolist(), x1.tolist()            # one host read (the reference loops per block, :717-734)
    for i in range(n):
        for b in range(y1.shape[1]):
            yy, xx = int(y1l[i][b]) * s, int(x1l[i][b]) * s
            out.append(f[i, :, yy: yy + size_y * s, xx: xx + size_x * s])
    return torch.stack(out, 0)


def fine_search(blk, win, return_corr=False):
    """search_org (:654-672): blk [M,C,k+2,k+2], win [M,C,d+2,d+2] -> (att [M,k,k], index [M,k,k])
    (+ the whole correlation [M, k*k, d*d] when return_corr: the tests grade arg-max mismatches by the gap they jump)."""
    m, c, kh, kw = blk.shape
    _, _, wh, ww = win.shape
    a = torch.stack([blk[:, :, ty: ty + kh - 2, tx: tx + kw - 2] for ty in range(3) for tx in range(3)], 2)
    r = torch.stack([win[:, :, ty: ty + wh - 2, tx: tx + ww - 2] for ty in range(3) for tx in range(3)], 2)
    a = _l2n(a.reshape(m, c * 9, (kh - 2) * (kw - 2)), 1)
    r = _l2n(r.reshape(m, c * 9, (wh - 2) * (ww - 2)), 1)
    corr = a.transpose(1, 2) @ r                                             # [M, k*k, d*d]
    att, idx = corr.max(-1)
    if return_corr:
        return att.view(m, kh - 2, kw - 2), idx.view(m, kh - 2, kw - 2), corr
    return att.view(m, kh - 2, kw - 2), idx.view(m, kh - 2, kw - 2)


def transfer(win_s, index, att, s, d_x):
    """transfer (:698-715) in closed form.

    win_s [M, C, (d+2)s, (d+2)s]; index/att [M,k,k].  Output [M, C, k*s, k*s]: every pixel is the mean of
    the 4..9 overlapping matched (3s x 3s) ref patches, times the bilinearly up-sampled confidence."""
    m, c = win_s.shape[:2]
    k_y, k_x = index.shape[1:]
    jy, jx = index // d_x, index % d_x
    dev = win_s.device
    ys = torch.arange(k_y * s, device=dev)
    xs = torch.arange(k_x * s, device=dev)
    acc = torch.zeros(m, c, k_y * s, k_x * s, device=dev, dtype=win_s.dtype)
    cnt = torch.zeros(1, 1, k_y * s, k_x * s, device=dev)
    mi = torch.arange(m, device=dev).view(m, 1, 1)
    for oy in (-1, 0, 1):
        by = ys // s + oy
        vy = (by >= 0) & (by < k_y)
        byc = by.clamp(0, k_y - 1)
        for ox in (-1, 0, 1):
            bx = xs // s + ox
            vx = (bx >= 0) & (bx < k_x)
            bxc = bx.clamp(0, k_x - 1)
            sy = jy[:, byc][:, :, bxc] * s + (ys - byc * s + s).view(1, -1, 1)     # [M, Ys, Xs]
            sx = jx[:, byc][:, :, bxc] * s + (xs - bxc * s + s).view(1, 1, -1)
            valid = (vy.view(-1, 1) & vx.view(1, -1)).float()
            sy = sy.clamp(0, win_s.shape[2] - 1)
            sx = sx.clamp(0, win_s.shape[3] - 1)
            g = win_s[mi, :, sy, sx].permute(0, 3, 1, 2)                            # [M,C,Ys,Xs]
            acc = acc + g * valid
            cnt = cnt + valid
    att_up = F.interpolate(att.unsqueeze(1), size=(k_y * s, k_x * s), mode="bilinear", align_corners=False)
    return acc / cnt * att_up


def retile(t, n, py, px):
    """[N*py*px, C, a, b] -> [N, C, py*a, px*b]  (:877-891)."""
    _, c, a, b = t.shape
    return t.view(n, py, px, c, a, b).permute(0, 3, 1, 4, 2, 5).reshape(n, c, py * a, px * b)


def masa_warp(feat_lq_deep, feat_ref, padder_size, lr_block_size, ref_down_block_size, dilations,
              h, w, hr, wr, return_aux=False):
    """Everything between masa_enc and the fusion blocks (:753-900).

    feat_ref: list deepest-last (scale 1 = deepest).  Returns [warp at finest ... warp at deepest]."""
    n = feat_lq_deep.shape[0]
    px = w // padder_size // lr_block_size
    py = h // padder_size // lr_block_size
    k_x = w // padder_size // px
    k_y = h // padder_size // py
    d_x = 2 * int(wr // padder_size // (2 * px) * ref_down_block_size) + 1
    d_y = 2 * int(hr // padder_size // (2 * py) * ref_down_block_size) + 1
    f_ref_deep = feat_ref[-1]
    _, c, hr_d, wr_d = f_ref_deep.shape
    blocks = lq_blocks(feat_lq_deep, k_y, k_x)
    score, idx = coarse_search(blocks, f_ref_deep, dilations)
    y1, x1 = window_origin(idx, wr_d, hr_d, d_x, d_y)
    win1 = crop_windows(f_ref_deep, y1, x1, d_y + 2, d_x + 2, 1)
    att, index, corr = fine_search(blocks.reshape(n * py * px, c, k_y + 2, k_x + 2), win1, return_corr=True)
    warps = []
    nlev = len(feat_ref)
    for lev in range(nlev):                       # lev 0 = finest level, scale 2**(nlev-1)
        s = 2 ** (nlev - 1 - lev)
        win = crop_windows(feat_ref[lev], y1, x1, d_y + 2, d_x + 2, s)
        warps.append(retile(transfer(win, index, att, s, d_x), n, py, px))
    if return_aux:
        return warps, dict(score=score, idx=idx, y1=y1, x1=x1, att=att, index=index, corr=corr, d=(d_y, d_x),
                           k=(k_y, k_x))
    return warps


def pad_to(x, mult):
    _, _, h, w = x.shape
    return F.pad(x, (0, (mult - w % mult) % mult, 0, (mult - h % mult) % mult))


def restormer_ref_fusion_forward(sd, inp_img, ref_img, heads=(1, 2, 4, 8), lr_block_size=8,
                                 ref_down_block_size=1.5, dilations=(1, 2, 3), return_aux=False):
    """RestormerRefFusion.forward (:747-964) with the B1 index shift (deepest = 1/8 scale)."""
    padder = 8
    _, _, oh, ow = inp_img.shape
    inp_img = pad_to(inp_img, padder * lr_block_size)
    ref_img = pad_to(ref_img, padder * lr_block_size)
    _, _, h, w = inp_img.shape
    _, _, hr, wr = ref_img.shape
    f_lq = masa_encoder(sd, inp_img)
    f_ref = masa_encoder(sd, ref_img)
    res = masa_warp(f_lq[-1], f_ref, padder, lr_block_size, ref_down_block_size, dilations, h, w, hr, wr,
                    return_aux=return_aux)
    warps, aux = res if return_aux else (res, None)

    def fuse(x, warp, name, hd):
        cat = torch.cat([x, warp], 1)
        return _stack(sd, name, cat, hd, fn=res_fusion_block)[:, : x.shape[1]]

    x1 = _conv(sd, "patch_embed.proj", inp_img, padding=1)
    x1 = fuse(x1, warps[0], "masa_blk_enc_level1", heads[0])
    e1 = _stack(sd, "encoder_level1", x1, heads[0])
    x2 = fuse(downsample(sd, "down1_2", e1), warps[1], "masa_blk_enc_level2", heads[1])
    e2 = _stack(sd, "encoder_level2", x2, heads[1])
    x3 = fuse(downsample(sd, "down2_3", e2), warps[2], "masa_blk_enc_level3", heads[2])
    e3 = _stack(sd, "encoder_level3", x3, heads[2])
    x4 = fuse(downsample(sd, "down3_4", e3), warps[3], "masa_blk_enc_level4", heads[3])
    lat = _stack(sd, "latent", x4, heads[3])
    out = _decoder(sd, heads, lat, e1, e2, e3, inp_img, x1)[:, :, :oh, :ow]
    if return_aux:
        aux["warps"] = warps
        aux["feat_lq"] = f_lq
        aux["feat_ref"] = f_ref
        return out, aux
    return out
