"""Training-step schedules (forward with a tape + explicit backward) for ``NAFNet`` / ``NAFNetRefFusion``.

Same construction as ``restormer_train`` (the reference differentiates stock PyTorch ops,
/root/reference/models/image_restoration_ref_model.py:251-284): data gradients are ``tdr_conv_gemm`` /
``tdr_dwconv3x3`` with transposed + flipped weights, weight gradients ``tdr_wgrad`` / ``tdr_dwconv3x3_wgrad`` /
``tdr_colsum``; LayerNorm2d (nafnet_arch_utils.py:264-300) uses ``tdr_rownorm_bwd``; SimpleGate ``tdr_gate_bwd``.

The forward folds ``beta`` / ``gamma`` and the simplified channel attention into the conv3 / conv5 weights
(``tdr_naf_sca_fold``); the backward un-folds them from the RAW weight-gradient matrices
(``tdr_naf_scaled_conv_bwd`` / ``tdr_naf_sca_bwd``), so that zero-initialised ``beta`` / ``gamma`` still receive their
gradient.
"""
import torch

from .. import ops
from .restormer_train import Grads, _flip_T, _head_map

F32, BF16 = torch.float32, torch.bfloat16


def prep_naf_train(blk, p):
    if "w1_T" in p:
        return p
    gamma = blk.gamma.detach().float().reshape(-1)
    p["mod"] = blk
    p["w1_T"] = ops.pack_conv_weight(_flip_T(blk.conv1.weight))
    p["w2_f"] = ops.pack_dw_weight(blk.conv2.weight.detach().flip(2, 3))
    p["w4_T"] = ops.pack_conv_weight(_flip_T(blk.conv4.weight))
    p["w5_T"] = ops.pack_conv_weight(_flip_T(blk.conv5.weight.detach() * gamma.view(-1, 1, 1, 1)))
    p["w5_raw"] = blk.conv5.weight.detach().float().reshape(blk.conv5.out_channels, -1).contiguous()
    p["b3"] = blk.conv3.bias.detach().float().contiguous() if blk.conv3.bias is not None else None
    p["b5"] = blk.conv5.bias.detach().float().contiguous() if blk.conv5.bias is not None else None
    p["gamma"] = gamma.contiguous()
    return p


def run_naf_block_train(x32, p, tape):
    """Training forward of one NAFBlock / NAFResFuseBlock (N:178-238, :241-302): not in place."""
    c = p["C"]
    sv = dict(p=p, x0=x32)
    xn = ops.rownorm(x32, 1, p["n1_w"], p["n1_b"], p["eps"])
    sv["xn1"] = xn
    _, t1 = ops.conv_gemm(xn, p["w1"], p["dw"], bias=p["b1"])
    g, sv["ydw"] = ops.dwconv3x3_gated_train(t1, p["w2"], p["b2"], 2)
    w3eff = ops.naf_sca_fold(g, p["w_sca"], p["b_sca"], p["w3"], rowscale=p["beta"], save=sv)
    y, _ = ops.conv_gemm(g, w3eff, c, bias=p["b3_beta"], res2=x32, want="f32", w_batched=True)
    xn = ops.rownorm(y, 1, p["n2_w"], p["n2_b"], p["eps"])
    sv["xn2"] = xn
    _, t2 = ops.conv_gemm(xn, p["w4"], p["ffn"], bias=p["b4"])
    g2 = ops.gate_mul(t2)
    out, _ = ops.conv_gemm(g2, p["w5"], c, bias=p["b5_gamma"], res2=y, want="f32")
    sv.update(t1=t1, g=g, y=y, t2=t2, g2=g2)
    tape.append(sv)
    return out


def run_naf_block_bwd(dout, sv, G):
    """dout: fp32 NHWC gradient w.r.t. the block output (overwritten).  Returns the gradient w.r.t. the block input."""
    p = sv["p"]
    blk = p["mod"]
    c, dw, ffn = p["C"], p["dw"], p["ffn"]
    x0, t1, g, y, t2, g2 = sv["x0"], sv["t1"], sv["g"], sv["y"], sv["t2"], sv["g2"]
    B, H, W, _ = x0.shape
    dev = dout.device
    # ---- out = y + gamma * (conv5(SG(conv4(LN2(y)))) + b5)
    d16 = ops.rownorm(dout, 0)
    raw5 = torch.empty((1, c, ffn // 2), dtype=F32, device=dev)
    ops.wgrad(d16, g2, raw5, Co=c, Ci=ffn // 2, strides=(0, ffn // 2, 1, 0), accumulate=False)
    cs = torch.empty(c, dtype=F32, device=dev)
    ops.colsum(d16, cs, accumulate=False)
    ops.naf_scaled_conv_bwd(raw5, p["w5_raw"], p["b5"], p["gamma"], cs, None, G(blk.conv5.weight), G(blk.conv5.bias),
                            G(blk.gamma))
    _, dg2 = ops.conv_gemm(d16, p["w5_T"], ffn // 2, Ci=c)
    dt2 = ops.gate_bwd(t2, dg2, 2)
    ops.wgrad(dt2, sv["xn2"], G(blk.conv4.weight))
    ops.colsum(dt2, G(blk.conv4.bias))
    _, dxn2 = ops.conv_gemm(dt2, p["w4_T"], c, Ci=ffn)
    dy, dy16 = ops.rownorm_bwd(y, dxn2, 1, p["n2_w"], p["eps"], add=dout, out=dout, dweight=G(blk.norm2.weight),
                               dbias=G(blk.norm2.bias), want16=True)
    # ---- y = x + beta * (conv3(g * sca(g)) + b3),  g = SG(dw3x3(conv1(LN1(x))))
    raw3 = torch.empty((B, c, dw // 2), dtype=F32, device=dev)
    ops.wgrad(dy16, g, raw3, Co=c, Ci=dw // 2, per_sample=True, strides=(c * (dw // 2), dw // 2, 1, 0), accumulate=False)
    ops.colsum(dy16, cs, accumulate=False)
    ops.naf_scaled_conv_bwd(raw3, p["w3"], p["b3"], p["beta"], cs, sv["s"], G(blk.conv3.weight), G(blk.conv3.bias),
                            G(blk.beta))
    dg_add = ops.naf_sca_bwd(raw3, p["w3"], p["beta"], sv["mean"], p["w_sca"], H * W, G(blk.sca[1].weight),
                             G(blk.sca[1].bias))
    _, dg = ops.conv_gemm(dy16, sv["weff_t"], dw // 2, Ci=c, w_batched=True)
    dyd = ops.gate_bwd(sv["ydw"], dg, 2, dg_add=dg_add)
    ops.dwconv3x3_wgrad(dyd, t1, G(blk.conv2.weight), G(blk.conv2.bias))
    dt1 = ops.dwconv3x3(dyd, p["w2_f"], None)
    ops.wgrad(dt1, sv["xn1"], G(blk.conv1.weight))
    ops.colsum(dt1, G(blk.conv1.bias))
    _, dxn1 = ops.conv_gemm(dt1, p["w1_T"], c, Ci=dw)
    return ops.rownorm_bwd(x0, dxn1, 1, p["n1_w"], p["eps"], add=dy, out=dy, dweight=G(blk.norm1.weight),
                           dbias=G(blk.norm1.bias))


def run_naf_stack_train(x32, preps, mods, tape):
    n0 = len(tape)
    for p, m in zip(preps, mods):
        x32 = run_naf_block_train(x32, prep_naf_train(m, p), tape)
    return x32, (n0, len(tape))


def run_naf_stack_bwd(d, tape, span, G):
    for i in range(span[1] - 1, span[0] - 1, -1):
        d = run_naf_block_bwd(d, tape[i], G)
        tape[i] = None
    return d


class NAFTrainMixin:
    """Training forward / backward of the NAFNet U-Net body (N:356-379) and, with ``fuse`` stages, of NAFNetRefFusion."""

    @staticmethod
    def _image16(img, pad_h, pad_w):
        t = torch.zeros((img.shape[0], pad_h, pad_w, 8), dtype=BF16, device=img.device)
        ops.nchw_to_nhwc_into(img, pad_h, pad_w, dst16=t)
        return t

    def _prep_train(self, P):
        if P.get("_train"):
            return P
        for pd, d in zip(P["downs"], self.downs):
            # 2x2 stride-2 conv: its data gradient is a 1x1 conv Co -> 4*Ci followed by PixelShuffle(2):
            # row (ci*4 + ky*2 + kx) of the dgrad weight = W[:, ci, ky, kx]
            w = d.weight.detach()                                       # [Co, Ci, 2, 2]
            pd["wT"] = ops.pack_conv_weight(w.permute(1, 2, 3, 0).reshape(w.shape[1] * 4, w.shape[0], 1, 1))
            pd["mod"] = d
        for pu, u in zip(P["ups"], self.ups):
            pu["wT"] = ops.pack_conv_weight(_flip_T(u[0].weight))
            pu["mod"] = u[0]
        ew = self.ending.weight
        w8 = torch.zeros(8, ew.shape[1], 3, 3, dtype=ew.dtype, device=ew.device)
        w8[:ew.shape[0]] = ew.detach()
        P["ending"]["wT"] = ops.pack_conv_weight(_flip_T(w8))
        P["_train"] = True
        return P

    # ---- U-Net ----------------------------------------------------------------------------------------------------
    def _unet_train(self, P, x32, tape, T, fbuf=None):
        """fbuf: per-stage fusion buffers [x || warp] (guided net); the stream of stage i is fbuf[i][..., :C]."""
        dev = x32.device
        n_enc = len(P["encoders"])
        encs = []
        T["downs"], T["ups"] = [], []
        for i in range(n_enc + 1):
            if fbuf is not None:
                yf, T[f"s_fuse{i}"] = run_naf_stack_train(fbuf[i], P["fuse"][i],
                                                          self.masa_blk_enc[i] if i < n_enc else self.masa_blk_middle[0], tape)
                x32 = yf[..., :yf.shape[3] // 2]
            if i == n_enc:
                break
            x32, T[f"s_enc{i}"] = run_naf_stack_train(x32, P["encoders"][i], self.encoders[i], tape)
            encs.append(x32)
            B, H, W, c = x32.shape
            nxt = fbuf[i + 1][..., :2 * c] if fbuf is not None else torch.empty((B, H // 2, W // 2, 2 * c), dtype=F32, device=dev)
            pd = P["downs"][i]
            ops.conv_gemm(ops.rownorm(x32, 0), pd["w"], pd["Co"], k=2, stride=2, pad=0, bias=pd["b"], out_f32=nxt)
            T["downs"].append(x32)
            x32 = nxt
        x32, T["s_mid"] = run_naf_stack_train(x32, P["middle"], self.middle_blks, tape)
        for i, skip in enumerate(encs[::-1]):
            B, H, W, c = x32.shape
            up = torch.empty((B, H * 2, W * 2, c // 2), dtype=F32, device=dev)
            ops.conv_gemm(ops.rownorm(x32, 0), P["ups"][i]["w"], P["ups"][i]["Co"], out_f32=up, res2=skip, store_mode=2)
            T["ups"].append(x32)
            x32, T[f"s_dec{i}"] = run_naf_stack_train(up, P["decoders"][i], self.decoders[i], tape)
        T["last"] = x32
        o8, _ = ops.conv_gemm(ops.rownorm(x32, 0), P["ending"]["w"], 8, k=3, pad=1, bias=P["ending"]["b"], want="f32")
        return o8[..., :P["ending"]["Co"]]

    def _unet_bwd(self, P, dout_nchw, h, w, tape, T, G, guided=False):
        """Returns (dx0, dwarps): gradient w.r.t. the intro-conv output and (guided) the warped-reference gradients."""
        dev = dout_nchw.device
        B = dout_nchw.shape[0]
        n_enc = len(P["encoders"])
        co = P["ending"]["Co"]
        do8 = torch.zeros((B, h, w, 8), dtype=BF16, device=dev)
        ops.nchw_to_nhwc_into(dout_nchw, h, w, dst16=do8)
        m = _head_map(8, co, dev)
        ops.wgrad(do8, ops.rownorm(T["last"], 0), G(self.ending.weight), k=3, pad=1, co_map=m)
        ops.colsum(do8, G(self.ending.bias), c_map=m)
        d, _ = ops.conv_gemm(do8, P["ending"]["wT"], self.width, Ci=8, k=3, pad=1, want="f32")
        dskips = []
        for i in range(n_enc - 1, -1, -1):
            d = run_naf_stack_bwd(d, tape, T[f"s_dec{i}"], G)
            dskips.append(d)                                              # `up + skip`: the skip gets d as is
            conv = P["ups"][i]["mod"]
            dconv = ops.pixel_shuffle(ops.rownorm(d, 0), 1)
            ops.wgrad(dconv, ops.rownorm(T["ups"][i], 0), G(conv.weight))
            d, _ = ops.conv_gemm(dconv, P["ups"][i]["wT"], conv.in_channels, want="f32")
        dskips = dskips[::-1]                                             # dskips[j] belongs to encoder stage n_enc-1-j
        d = run_naf_stack_bwd(d, tape, T["s_mid"], G)
        dwarps = [None] * (n_enc + 1)
        for i in range(n_enc, -1, -1):
            if i < n_enc:
                # down conv i: x_{i+1} = conv2x2s2(enc_i) + b
                pd = P["downs"][i]
                conv = pd["mod"]
                d16 = ops.rownorm(d, 0)
                ops.wgrad(d16, ops.rownorm(T["downs"][i], 0), G(conv.weight), k=2, stride=2, pad=0)
                ops.colsum(d16, G(conv.bias))
                dx = torch.empty(T["downs"][i].shape, dtype=F32, device=dev)
                ops.conv_gemm(d16, pd["wT"], 4 * conv.in_channels, out_f32=dx, res2=dskips[n_enc - 1 - i], store_mode=2)
                d = run_naf_stack_bwd(dx, tape, T[f"s_enc{i}"], G)
            if guided:
                C_ = d.shape[3]
                dfo = torch.zeros(d.shape[:3] + (2 * C_,), dtype=F32, device=dev)
                ops.copy_rows(d, dst32=dfo[..., :C_])
                dfb = run_naf_stack_bwd(dfo, tape, T[f"s_fuse{i}"], G)
                dwarps[i] = dfb[..., C_:]
                d = dfb[..., :C_]
        return d, dwarps

    def _intro_bwd(self, dx0, img16, G):
        dx16 = ops.rownorm(dx0, 0)
        ops.wgrad(dx16, img16, G(self.intro.weight), k=3, pad=1, ci_map=_head_map(8, self.img_channel, dx0.device))
        ops.colsum(dx16, G(self.intro.bias))

    # ---- plain NAFNet ----------------------------------------------------------------------------------------------
    def _forward_train(self, inp):
        self._check(inp)
        P = self._prep_train(self.prepared(train=True))
        B, _, H, W = inp.shape
        h, w = ops.round_up(H, self.padder_size), ops.round_up(W, self.padder_size)
        inp32 = ops.nchw_to_nhwc(inp, h, w)
        tape, T = [], dict(hw=(h, w), inp16=self._image16(inp, h, w))
        x = torch.empty((B, h, w, self.width), dtype=F32, device=inp.device)
        ops.conv3x3_small_ci(inp32, P["intro"]["w"], P["intro"]["b"], out_f32=x)
        out = self._unet_train(P, x, tape, T)
        return ops.nhwc_to_nchw(out, H, W, res=inp32), (P, tape, T)

    def _backward(self, state, dout, G=None):
        P, tape, T = state
        G = Grads() if G is None else G
        h, w = T["hw"]
        dx0, _ = self._unet_bwd(P, dout.contiguous().float(), h, w, tape, T, G)
        self._intro_bwd(dx0, T["inp16"], G)
        return G


class GuidedNAFTrainMixin(NAFTrainMixin):
    """NAFNetRefFusion (N:587-740): MASA encoder + match/transfer, NAFResFuseBlock stages, U-Net."""

    def _forward_train(self, inp, ref):
        self._check(inp, ref)
        P = self._prep_train(self.prepared(train=True))
        E = self._prep_masa_train(P["masa_enc"])
        dev = inp.device
        B, _, oh, ow = inp.shape
        mult = self.padder_size * self.lr_block_size
        h, w = ops.round_up(oh, mult), ops.round_up(ow, mult)
        hr, wr = ops.round_up(ref.shape[2], mult), ops.round_up(ref.shape[3], mult)
        if (h, w) == (hr, wr):           # lq and ref share one batch buffer (no concatenation copy)
            both32 = torch.empty((2 * B, h, w, inp.shape[1]), dtype=F32, device=dev)
            both16 = torch.zeros((2 * B, h, w, 8), dtype=BF16, device=dev)
            lq32, ref32, lq16, ref16 = both32[:B], both32[B:], both16[:B], both16[B:]
            ops.nchw_to_nhwc_into(inp, h, w, dst32=lq32, dst16=lq16)
            ops.nchw_to_nhwc_into(ref, hr, wr, dst32=ref32, dst16=ref16)
        else:
            lq32, ref32 = ops.nchw_to_nhwc(inp, h, w), ops.nchw_to_nhwc(ref, hr, wr)
            lq16, ref16 = self._image16(inp, h, w), self._image16(ref, hr, wr)
        tape, T = [], dict(hw=(h, w), B=B, inp16=lq16)
        if (h, w) == (hr, wr):
            fb, d32, et = self._masa_encode_train(E, both32, both16)
            f_lq, f_ref, lq_d32, ref_d32 = [t[:B] for t in fb], [t[B:] for t in fb], d32[:B], d32[B:]
            T["enc"] = [(et, fb)]
        else:
            f_lq, lq_d32, et_l = self._masa_encode_train(E, lq32, lq16)
            f_ref, ref_d32, et_r = self._masa_encode_train(E, ref32, ref16)
            T["enc"] = [(et_l, f_lq), (et_r, f_ref)]
        nlev = len(f_ref)
        chans = [self.width * 2 ** i for i in range(nlev)]
        fbuf = [torch.empty((B, h >> i, w >> i, 2 * chans[i]), dtype=F32, device=dev) for i in range(nlev)]
        aux = self._masa_warp(lq_d32, ref_d32, f_ref, h, w, hr, wr, [fbuf[i][..., chans[i]:] for i in range(nlev)])
        T["aux"], T["f_lq_deep"], T["f_ref"] = aux, f_lq[-1], f_ref
        ops.conv3x3_small_ci(lq32, P["intro"]["w"], P["intro"]["b"], out_f32=fbuf[0][..., :chans[0]])
        out = self._unet_train(P, fbuf[0][..., :chans[0]], tape, T, fbuf=fbuf)
        return ops.nhwc_to_nchw(out, oh, ow, res=lq32), (P, tape, T)

    def _backward(self, state, dout, G=None):
        P, tape, T = state
        E = P["masa_enc"]
        G = Grads() if G is None else G
        h, w = T["hw"]
        B = T["B"]
        dev = dout.device
        dx0, dwarps = self._unet_bwd(P, dout.contiguous().float(), h, w, tape, T, G, guided=True)
        self._intro_bwd(dx0, T["inp16"], G)
        f_ref = T["f_ref"]
        if len(T["enc"]) == 1:
            et, fb = T["enc"][0]
            dfeat = [torch.zeros(t.shape, dtype=F32, device=dev) for t in fb]
            self._masa_warp_bwd(T["aux"], T["f_lq_deep"], f_ref, dwarps, dfeat[-1][:B], [t[B:] for t in dfeat])
            self._masa_encode_bwd(E, et, dfeat, G)
        else:
            (et_l, f_lq), (et_r, _) = T["enc"]
            dlq = [torch.zeros(t.shape, dtype=F32, device=dev) for t in f_lq]
            dref = [torch.zeros(t.shape, dtype=F32, device=dev) for t in f_ref]
            self._masa_warp_bwd(T["aux"], T["f_lq_deep"], f_ref, dwarps, dlq[-1], dref)
            self._masa_encode_bwd(E, et_l, dlq, G)
            self._masa_encode_bwd(E, et_r, dref, G)
        return G
