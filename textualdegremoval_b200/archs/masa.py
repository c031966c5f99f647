"""MASA guidance pieces shared by the guided Restormer and guided NAFNet modules (not an ``*_arch.py`` file, so the
registry does not scan it): the feature ``Encoder`` (parameter holder), its kernel schedule, and the match-and-transfer
schedule (reference network_restormer_guided_arch.py:100-134, :642-734, :753-900 and the identical copy in
network_nafnet_guided_arch.py:110-143, :483-707)."""
import torch
import torch.nn as nn

from .. import ops

F32, BF16, F16 = torch.float32, torch.bfloat16, torch.float16


def _f(t):
    return None if t is None else t.detach().float().contiguous()


def prep_conv(conv: nn.Conv2d, dt=BF16):
    return dict(w=ops.pack_conv_weight(conv.weight, dt=dt), b=_f(conv.bias), Co=conv.out_channels, Ci=conv.in_channels,
                stride=conv.stride[0], k=conv.kernel_size[0], pad=conv.padding[0])


def conv3x3(x16, pc, **kw):
    bias = kw.pop("bias", pc["b"])
    return ops.conv_gemm(x16, pc["w"], pc["Co"], k=3, stride=pc["stride"], pad=1, bias=bias, **kw)


def prep_conv_f16(conv: nn.Conv2d):
    """Feature-encoder convs: IEEE fp16 operands (see MasaMixin._masa_encode)."""
    return prep_conv(conv, dt=F16)


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
        nlev = self.masa_enc.levels
        for i in range(1, nlev + 1):
            c = getattr(self.masa_enc, f"conv_L{i}")
            enc[f"conv_L{i}"] = dict(w=_f(c.weight), b=_f(c.bias), w16=ops.pack_conv(c.weight, Ci_p=8, dt=F16)[0],
                                     Co=c.out_channels) if i == 1 else prep_conv_f16(c)
            enc[f"blk_L{i}"] = [(prep_conv_f16(b.conv1), prep_conv_f16(b.conv2))
                                for b in getattr(self.masa_enc, f"blk_L{i}")]
        # biases grouped by the level scale they are expressed in (see _masa_encode): group L = the residual blocks of
        # level L and the stride-2 conv that LEAVES it; one flat fp32 buffer per group, slices 16 B-aligned
        groups = []
        for i in range(1, nlev + 1):
            convs = [c for pair in enc[f"blk_L{i}"] for c in pair] + ([enc[f"conv_L{i + 1}"]] if i < nlev else [])
            if not convs:
                groups.append((None, []))
                continue
            flat = torch.cat([c["b"].reshape(-1) for c in convs]).contiguous()
            offs, o = [], 0
            for c in convs:
                offs.append((c, o, o + c["b"].numel()))
                o += c["b"].numel()
            groups.append((flat, offs))
        enc["bias_groups"] = groups
        dev = self.masa_enc.conv_L1.weight.device
        st = torch.zeros((nlev, 4), dtype=F32, device=dev)
        st[0, :3] = 1.0
        enc["state_init"] = st
        return enc

    # ---- MASA encoder (:100-134) on a batch of images ---------------------------------------------
    # Precision plan (tools/precision_study.py; the features feed two top-1 searches whose near-ties flip on bf16 noise):
    # the residual stream of every level is fp32 (as the transformer blocks' is); the GEMM operands are IEEE fp16 -- 11
    # significand bits instead of bf16's 8 at the same tcgen05 rate -- and the deepest level is handed to the searches in
    # fp32.  The bf16 copies of the level outputs serve the transfer / backward kernels.
    # Range plan: 17-21 ReLU convs with residual adds and NO normalisation can grow geometrically (x10 per level with
    # variance-preserving weights: 1.7e5 at level 5, beyond fp16's 65504).  Conv, bias, ReLU and the residual add are
    # positively homogeneous, so level L runs in units of a power-of-two scale s_L picked on the device from the level's
    # first activation (tdr_masa_level_scale: no host sync; scaling by 2^k changes no mantissa, so the result does not
    # depend on the choice).  Biases enter as b / s_L; the bf16 level outputs are multiplied back; the searches are
    # scale-invariant (cosine similarities), so the deepest fp32 features stay in scaled units.
    def _masa_encode(self, E, img32, tape=None, img16=None):
        """Returns (feats, deep32): bf16 NHWC features per level (finest first, unscaled) and the fp32 deepest-level
        stream (in units of its level scale).  tape: optional dict that receives every saved activation (scaled units)
        and the level-scale states ``S`` (training forward)."""
        feats = []
        B, H, W, _ = img32.shape
        dev = img32.device
        nlev = self.masa_enc.levels
        S = E["state_init"].clone()              # S[l] = {s, 1/s, 1/r, max}; level 1 is unscaled
        x32 = torch.empty((B, H, W, self.masa_enc.nf), dtype=F32, device=dev)
        xh = torch.empty((B, H, W, self.masa_enc.nf), dtype=F16, device=dev) if img16 is None else None
        if img16 is not None:       # tensor-core path: 8-channel fp16 image rows, K = 16 per tap, then one cast pass
            ops.conv_gemm(img16, E["conv_L1"]["w16"], E["conv_L1"]["Co"], Ci=8, k=3, pad=1, bias=E["conv_L1"]["b"], relu=True,
                          out_f32=x32)
            _, xh = ops.cast_rows(x32, want_bf16=False, want_fp16=True)
        else:
            ops.conv3x3_small_ci(img32, E["conv_L1"]["w"], E["conv_L1"]["b"], relu=True, out_f32=x32, out_bf16=xh)
        xb = None
        bias_next = None                         # conv_L{lvl+1}'s bias in units of s_lvl
        for lvl in range(1, nlev + 1):
            x_prev = xh
            blks = E[f"blk_L{lvl}"]
            st = S[lvl - 1]
            if lvl > 1:
                x32, _ = conv3x3(xh, E[f"conv_L{lvl}"], relu=True, want="f32", bias=bias_next)     # in units of s_{lvl-1}
                ops.masa_level_scale(x32, S[lvl - 2], st)
                xb, xh = ops.cast_rows(x32, want_bf16=not blks, want_fp16=bool(blks) or lvl < nlev, scale16=st[2:3],
                                       scale_bf16=S[lvl - 2][0:1], rescale_in=True)                # now in units of s_lvl
            elif not blks:
                xb, _ = ops.cast_rows(x32)
            flat, offs = E["bias_groups"][lvl - 1]
            bias = {}
            if flat is not None:
                fs = ops.scale_vec(flat, st[1:2])
                bias = {id(c): fs[a:b] for c, a, b in offs}
            bias_next = bias.get(id(E[f"conv_L{lvl + 1}"])) if lvl < nlev else None
            y0 = xh
            saved = []
            for j, (c1, c2) in enumerate(blks):
                last = j == len(blks) - 1
                _, t = conv3x3(xh, c1, relu=True, out_fp16=True, bias=bias[id(c1)])
                conv3x3(t, c2, res2=x32, out_f32=x32, bias=bias[id(c2)])                          # fp32 stream, in place
                saved.append((xh, t))
                xb, xh = ops.cast_rows(x32, want_bf16=last, want_fp16=(not last) or lvl < nlev, scale_bf16=st[0:1])
            if tape is not None:
                tape["levels"].append(dict(x_prev=x_prev, y0=y0, blocks=saved))
            feats.append(xb)
        if tape is not None:
            tape["S"] = S
        self._masa_last_scale = S
        return feats, x32

    # ---- MASA search + transfer (:753-900) ----------------------------------------------------------
    def _masa_warp(self, lq_deep32, ref_deep32, f_ref, h, w, hr, wr, targets):
        """lq_deep32 / ref_deep32: fp32 deepest-level features (dense NHWC); f_ref: bf16 reference features per level.
        targets[lev] = fp32 NHWC view receiving warp at level lev (0 = finest).  Returns aux tensors."""
        ps, lb = self.padder_size, self.lr_block_size
        px, py = w // ps // lb, h // ps // lb
        k_x, k_y = w // ps // px, h // ps // py
        d_x = 2 * int(wr // ps // (2 * px) * self.ref_down_block_size) + 1
        d_y = 2 * int(hr // ps // (2 * py) * self.ref_down_block_size) + 1
        B, Hr, Wr, Cd = ref_deep32.shape
        if Wr < d_x + 2 or Hr < d_y + 2:
            raise ValueError(f"reference image too small for the MASA search window ({d_y + 2}x{d_x + 2} at 1/8 scale)")
        nblk = py * px
        co_pad = ops.round_up(nblk, 8)
        dils = self.dilations
        # coarse search: 3 dilated 3x3 "convs" of the ref feature with the normalised lq block descriptors
        # Both correlations are ONE bf16 GEMM over 3C channels each: reference [hi | lo | hi] x descriptors [hi | hi | lo]
        # (16 significand bits, fp32 accumulation); norms come from the fp32 features.
        n2 = ops.sqnorm_rows(ref_deep32)
        inv = ops.masa_ref_invnorm(n2, dils)
        fr = ops.masa_split3(ref_deep32)
        wc = ops.masa_coarse_filters(lq_deep32, k_y, k_x, dils, co_pad)
        score = torch.empty((B, Hr, Wr, co_pad), dtype=F32, device=fr.device)
        for i, dl in enumerate(dils):
            ops.conv_gemm(fr, wc[i], co_pad, k=3, pad=dl, dil=dl, rowscale=inv[i], res2=score if i else None,
                          out_f32=score, w_batched=True)
        idx, origin = ops.masa_coarse_argmax(score, nblk, d_y, d_x)
        # fine search inside each (d+2)^2 window
        wf = ops.masa_fine_filters(lq_deep32, k_y, k_x)
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
        """Like _masa_encode, keeping every intermediate activation (fp16: conv inputs and the post-ReLU tensors).
        Returns (feats, deep32, tape)."""
        tape = dict(img16=img16, levels=[])
        feats, deep32 = self._masa_encode(E, img32, tape=tape)
        return feats, deep32, tape

    def _masa_encode_bwd(self, E, tape, dfeats, G):
        """dfeats[lev]: fp32 NHWC gradient w.r.t. the level-lev feature (same batch as the forward).  Consumed in place."""
        enc = self.masa_enc
        d = None
        S = tape["S"]                                        # tape activations of level l are in units of S[l-1][0]
        for lvl in range(enc.levels, 0, -1):
            T = tape["levels"][lvl - 1]
            sc = S[lvl - 1][0:1]
            d = dfeats[lvl - 1] if d is None else d          # deeper levels already added their part (res2 below)
            mods = getattr(enc, f"blk_L{lvl}")
            for (c1, c2), m, (x_in, t) in reversed(list(zip(E[f"blk_L{lvl}"], mods, T["blocks"]))):
                d16 = ops.rownorm(d, 0)
                ops.wgrad(d16, t, G(m.conv2.weight), k=3, pad=1, scale_ptr=sc)
                ops.colsum(d16, G(m.conv2.bias))
                _, dt = ops.conv_gemm(d16, c2["wT"], c2["Ci"], k=3, pad=1)
                dt = ops.relu_mask(t, dt, out=dt)
                ops.wgrad(dt, x_in, G(m.conv1.weight), k=3, pad=1, scale_ptr=sc)
                ops.colsum(dt, G(m.conv1.bias))
                ops.conv_gemm(dt, c1["wT"], c1["Ci"], k=3, pad=1, res2=d, out_f32=d)
            conv = getattr(enc, f"conv_L{lvl}")
            dy = ops.relu_mask(T["y0"], ops.rownorm(d, 0))
            ops.colsum(dy, G(conv.bias))
            if lvl > 1:
                xp = T["x_prev"]
                ops.wgrad(dy, xp, G(conv.weight), k=3, stride=2, pad=1, scale_ptr=S[lvl - 2][0:1])
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
