"""MASA guidance pieces shared by the guided Restormer and guided NAFNet modules (not an ``*_arch.py`` file, so the
registry does not scan it): the feature ``Encoder`` (parameter holder), its kernel schedule, and the match-and-transfer
schedule (reference network_restormer_guided_arch.py:100-134, :642-734, :753-900 and the identical copy in
network_nafnet_guided_arch.py:110-143, :483-707)."""
import torch
import torch.nn as nn

from .. import ops

F32, BF16 = torch.float32, torch.bfloat16


def _f(t):
    return None if t is None else t.detach().float().contiguous()


def prep_conv(conv: nn.Conv2d):
    return dict(w=ops.pack_conv_weight(conv.weight), b=_f(conv.bias), Co=conv.out_channels, Ci=conv.in_channels,
                stride=conv.stride[0], k=conv.kernel_size[0], pad=conv.padding[0])


def conv3x3(x16, pc, **kw):
    return ops.conv_gemm(x16, pc["w"], pc["Co"], k=3, stride=pc["stride"], pad=1, bias=pc["b"], **kw)


class ResidualBlock(nn.Module):
    def __init__(self, nf):
        super().__init__()
        self.conv1 = nn.Conv2d(nf, nf, 3, 1, 1)
        self.conv2 = nn.Conv2d(nf, nf, 3, 1, 1)


class Encoder(nn.Module):
    """MASA feature extractor: ``levels`` scales nf, 2nf, 4nf, ...; n_blks[2] is reused for every level >= 3 (as in
    the reference)."""

    def __init__(self, in_chl, nf, n_blks=(1, 1, 1), levels=4):
        super().__init__()
        prev = in_chl
        for i in range(1, levels + 1):
            c = nf * 2 ** (i - 1)
            n = n_blks[min(i - 1, 2)]
            setattr(self, f"conv_L{i}", nn.Conv2d(prev, c, 3, 1 if i == 1 else 2, 1, bias=True))
            setattr(self, f"blk_L{i}", nn.Sequential(*[ResidualBlock(c) for _ in range(n)]))
            prev = c
        self.levels = levels
        self.nf = nf


class MasaMixin:
    """Needs: self.masa_enc (Encoder), self.padder_size, self.lr_block_size, self.ref_down_block_size, self.dilations."""

    def prepare_masa_enc(self):
        enc = {}
        for i in range(1, self.masa_enc.levels + 1):
            c = getattr(self.masa_enc, f"conv_L{i}")
            enc[f"conv_L{i}"] = dict(w=_f(c.weight), b=_f(c.bias)) if i == 1 else prep_conv(c)
            enc[f"blk_L{i}"] = [(prep_conv(b.conv1), prep_conv(b.conv2)) for b in getattr(self.masa_enc, f"blk_L{i}")]
        return enc

    # ---- MASA encoder (:100-134) on a batch of images ---------------------------------------------
    def _masa_encode(self, E, img32):
        feats = []
        B, H, W, _ = img32.shape
        x = torch.empty((B, H, W, self.masa_enc.nf), dtype=BF16, device=img32.device)
        ops.conv3x3_small_ci(img32, E["conv_L1"]["w"], E["conv_L1"]["b"], relu=True, out_bf16=x)
        for lvl in range(1, self.masa_enc.levels + 1):
            if lvl > 1:
                _, x = conv3x3(x, E[f"conv_L{lvl}"], relu=True)
            for c1, c2 in E[f"blk_L{lvl}"]:
                _, t = conv3x3(x, c1, relu=True)
                _, x = conv3x3(t, c2, res2=x)
            feats.append(x)
        return feats

    # ---- MASA search + transfer (:753-900) ----------------------------------------------------------
    def _masa_warp(self, f_lq_deep, f_ref, h, w, hr, wr, targets):
        """targets[lev] = fp32 NHWC view receiving warp at level lev (0 = finest).  Returns aux tensors."""
        ps, lb = self.padder_size, self.lr_block_size
        px, py = w // ps // lb, h // ps // lb
        k_x, k_y = w // ps // px, h // ps // py
        d_x = 2 * int(wr // ps // (2 * px) * self.ref_down_block_size) + 1
        d_y = 2 * int(hr // ps // (2 * py) * self.ref_down_block_size) + 1
        fr = f_ref[-1]
        B, Hr, Wr, Cd = fr.shape
        if Wr < d_x + 2 or Hr < d_y + 2:
            raise ValueError(f"reference image too small for the MASA search window ({d_y + 2}x{d_x + 2} at 1/8 scale)")
        nblk = py * px
        co_pad = ops.round_up(nblk, 8)
        dils = self.dilations
        # coarse search: 3 dilated 3x3 "convs" of the ref feature with the normalised lq block descriptors
        n2 = ops.sqnorm_rows(fr)
        inv = ops.masa_ref_invnorm(n2, dils)
        wc = ops.masa_coarse_filters(f_lq_deep, k_y, k_x, dils, co_pad)
        score = torch.empty((B, Hr, Wr, co_pad), dtype=F32, device=fr.device)
        for i, dl in enumerate(dils):
            ops.conv_gemm(fr, wc[i], co_pad, k=3, pad=dl, dil=dl, rowscale=inv[i], res2=score if i else None,
                          out_f32=score, w_batched=True)
        idx, origin = ops.masa_coarse_argmax(score, nblk, d_y, d_x)
        # fine search inside each (d+2)^2 window
        wf = ops.masa_fine_filters(f_lq_deep, k_y, k_x)
        winv = ops.masa_win_invnorm(n2, origin, d_y, d_x)
        corr, _ = ops.conv_gemm(fr, wf, k_y * k_x, k=3, pad=0, rowscale=winv, want="f32", w_batched=True,
                                origin=origin, window=(d_y + 2, d_x + 2))
        index, att = ops.masa_fine_argmax(corr)
        nlev = len(f_ref)
        for lev in range(nlev):
            s = 2 ** (nlev - 1 - lev)
            ops.masa_transfer(f_ref[lev], origin, index, att, py, px, k_y, k_x, d_x, s, out32=targets[lev])
        return dict(idx=idx, origin=origin, index=index, att=att, score=score, corr=corr,
                    geom=dict(py=py, px=px, k_y=k_y, k_x=k_x, d_x=d_x, d_y=d_y))



# =============================================================================================== training (tape + backward)
class MasaTrainMixin:
    """Training forward (keeps every activation of the feature encoder and the match aux tensors) and explicit backward
    of the MASA guidance path.  The matches (coarse / fine arg-max) are constants; gradients flow through the gathered
    reference features (R:698-715) and through the confidence map ``soft_att`` (R:661-670)."""

    def _prep_masa_train(self, E):
        if E.get("_train"):
            return E
        for i in range(1, self.masa_enc.levels + 1):
            c = getattr(self.masa_enc, f"conv_L{i}")
            if i > 1:
                E[f"conv_L{i}"]["wT"] = ops.pack_conv(c.weight, fwd=False, dgrad=True)[1]
            for (c1, c2), b in zip(E[f"blk_L{i}"], getattr(self.masa_enc, f"blk_L{i}")):
                c1["wT"] = ops.pack_conv(b.conv1.weight, fwd=False, dgrad=True)[1]
                c2["wT"] = ops.pack_conv(b.conv2.weight, fwd=False, dgrad=True)[1]
        E["_train"] = True
        return E

    def _masa_encode_train(self, E, img32, img16):
        """Like _masa_encode, keeping every intermediate activation.  Returns (feats, tape)."""
        B, H, W, _ = img32.shape
        x = torch.empty((B, H, W, self.masa_enc.nf), dtype=BF16, device=img32.device)
        ops.conv3x3_small_ci(img32, E["conv_L1"]["w"], E["conv_L1"]["b"], relu=True, out_bf16=x)
        tape = dict(img16=img16, levels=[])
        feats = []
        for lvl in range(1, self.masa_enc.levels + 1):
            x_prev = x
            if lvl > 1:
                _, x = conv3x3(x, E[f"conv_L{lvl}"], relu=True)
            blocks = []
            y0 = x
            for c1, c2 in E[f"blk_L{lvl}"]:
                _, t = conv3x3(x, c1, relu=True)
                _, xn = conv3x3(t, c2, res2=x)
                blocks.append((x, t))
                x = xn
            tape["levels"].append(dict(x_prev=x_prev, y0=y0, blocks=blocks))
            feats.append(x)
        return feats, tape

    def _masa_encode_bwd(self, E, tape, dfeats, G):
        """dfeats[lev]: fp32 NHWC gradient w.r.t. the level-lev feature (same batch as the forward).  Consumed in place."""
        enc = self.masa_enc
        d = None
        for lvl in range(enc.levels, 0, -1):
            T = tape["levels"][lvl - 1]
            d = dfeats[lvl - 1] if d is None else d          # deeper levels already added their part (res2 below)
            mods = getattr(enc, f"blk_L{lvl}")
            for (c1, c2), m, (x_in, t) in reversed(list(zip(E[f"blk_L{lvl}"], mods, T["blocks"]))):
                d16 = ops.rownorm(d, 0)
                ops.wgrad(d16, t, G(m.conv2.weight), k=3, pad=1)
                ops.colsum(d16, G(m.conv2.bias))
                _, dt = ops.conv_gemm(d16, c2["wT"], c2["Ci"], k=3, pad=1)
                dt = ops.relu_mask(t, dt, out=dt)
                ops.wgrad(dt, x_in, G(m.conv1.weight), k=3, pad=1)
                ops.colsum(dt, G(m.conv1.bias))
                ops.conv_gemm(dt, c1["wT"], c1["Ci"], k=3, pad=1, res2=d, out_f32=d)
            conv = getattr(enc, f"conv_L{lvl}")
            dy = ops.relu_mask(T["y0"], ops.rownorm(d, 0))
            ops.colsum(dy, G(conv.bias))
            if lvl > 1:
                xp = T["x_prev"]
                ops.wgrad(dy, xp, G(conv.weight), k=3, stride=2, pad=1)
                dyd = ops.dilate2(dy, xp.shape[1], xp.shape[2])
                nxt = dfeats[lvl - 2]
                ops.conv_gemm(dyd, E[f"conv_L{lvl}"]["wT"], conv.in_channels, k=3, pad=1, res2=nxt, out_f32=nxt)
                d = nxt
            else:
                cin = conv.in_channels
                m = torch.full((8,), -1, dtype=torch.int32, device=dy.device)
                m[:cin] = torch.arange(cin, dtype=torch.int32, device=dy.device)
                ops.wgrad(dy, tape["img16"], G(conv.weight), k=3, pad=1, ci_map=m)

    def _masa_warp_bwd(self, aux, f_lq_deep, f_ref, dwarps, dfeat_lq_deep, dfeat_ref):
        """dwarps[lev]: fp32 NHWC gradient w.r.t. the warped features at level lev (views allowed).  Accumulates into
        dfeat_ref[lev] (fp32, zero-initialised, same shapes as f_ref[lev]) and dfeat_lq_deep."""
        g = aux["geom"]
        nlev = len(f_ref)
        datt = torch.zeros_like(aux["att"])
        for lev in range(nlev):
            s = 2 ** (nlev - 1 - lev)
            ops.masa_transfer_bwd(dwarps[lev], f_ref[lev], aux["origin"], aux["index"], aux["att"], g["py"], g["px"],
                                  g["k_y"], g["k_x"], g["d_x"], s, dfeat_ref[lev], datt)
        ops.masa_fine_bwd(f_lq_deep, f_ref[-1], aux["origin"], aux["index"], datt, g["k_y"], g["k_x"], g["d_x"],
                          dfeat_lq_deep, dfeat_ref[-1])
