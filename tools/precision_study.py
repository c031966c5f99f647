"""Which operand precision does the MASA search path need?  CPU study on the oracle (test infrastructure): round the
operands of the feature encoder's convs / the correlation descriptors to a given format, count arg-max flips against
the fp32 oracle.  Usage: python -m tools.precision_study [size]"""
import sys

import torch
import torch.nn.functional as F

from oracle import restormer as O, weights as W


def rnd(t, mode):
    if mode == "f32":
        return t
    if mode == "bf16":
        return t.bfloat16().float()
    if mode == "f16":
        return t.half().float()
    if mode == "tf32":                       # 10-bit mantissa, round to nearest
        i = t.view(torch.int32)
        i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32)
    if mode == "bf16x2":                     # hi + lo
        hi = t.bfloat16().float()
        return hi + (t - hi).bfloat16().float()
    raise ValueError(mode)


def encoder(sd, x, op, stream, p="masa_enc"):
    feats = []
    lvl = 1
    cv = lambda name, x, **kw: F.conv2d(rnd(x, op), rnd(sd[name + ".weight"], op), sd[name + ".bias"], **kw)
    while f"{p}.conv_L{lvl}.weight" in sd:
        if lvl == 1:
            x = F.relu(F.conv2d(x, sd[f"{p}.conv_L1.weight"], sd[f"{p}.conv_L1.bias"], padding=1))
        else:
            x = F.relu(cv(f"{p}.conv_L{lvl}", x, stride=2, padding=1))
        x = rnd(x, stream)
        i = 0
        while f"{p}.blk_L{lvl}.{i}.conv1.weight" in sd:
            q = f"{p}.blk_L{lvl}.{i}"
            x = rnd(cv(q + ".conv2", F.relu(cv(q + ".conv1", x, padding=1)), padding=1) + x, stream)
            i += 1
        feats.append(x)
        lvl += 1
    return feats


def run(sd, lq, ref, op, stream, desc):
    f_lq, f_ref = encoder(sd, lq, op, stream), encoder(sd, ref, op, stream)
    fl, fr = rnd(f_lq[-1], desc), [rnd(t, desc) if i == len(f_ref) - 1 else t for i, t in enumerate(f_ref)]
    h, w = lq.shape[2:]
    hr, wr = ref.shape[2:]
    warps, aux = O.masa_warp(fl, fr, 8, 8, 1.5, (1, 2, 3), h, w, hr, wr, return_aux=True)
    aux["warps"] = warps
    # oracle-side correlation of every fine candidate, to grade flips by the gap they jump
    n, c = fl.shape[:2]
    blocks = O.lq_blocks(fl, 8, 8)
    win1 = O.crop_windows(fr[-1], aux["y1"], aux["x1"], aux["d"][0] + 2, aux["d"][1] + 2, 1)
    blk = blocks.reshape(-1, c, 10, 10)
    a = torch.stack([blk[:, :, ty: ty + 8, tx: tx + 8] for ty in range(3) for tx in range(3)], 2)
    r = torch.stack([win1[:, :, ty: ty + 13, tx: tx + 13] for ty in range(3) for tx in range(3)], 2)
    a = O._l2n(a.reshape(blk.shape[0], c * 9, 64), 1); r = O._l2n(r.reshape(blk.shape[0], c * 9, 169), 1)
    aux["corr"] = a.transpose(1, 2) @ r
    return aux, f_lq, f_ref


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    torch.set_grad_enabled(False)
    shapes = {}
    nf = 48
    prev = 3
    for i in range(1, 5):
        c = nf * 2 ** (i - 1)
        shapes[f"masa_enc.conv_L{i}.weight"] = (c, prev, 3, 3); shapes[f"masa_enc.conv_L{i}.bias"] = (c,)
        for j in range(4):
            for k in (1, 2):
                shapes[f"masa_enc.blk_L{i}.{j}.conv{k}.weight"] = (c, c, 3, 3)
                shapes[f"masa_enc.blk_L{i}.{j}.conv{k}.bias"] = (c,)
        prev = c
    sd = W.seeded_state_dict(shapes, 7)
    lq = W.seeded_image("lq", (1, 3, size, size), 7)
    lq = F.avg_pool2d(F.pad(lq, (2, 2, 2, 2), mode="reflect"), 5, 1)          # some spatial structure
    ref = torch.roll(lq, (5, -7), (2, 3)) + 0.02 * (W.seeded_image("n", lq.shape, 7) - 0.5)
    base, bl, br = run(sd, lq, ref, "f32", "f32", "f32")
    n = base["index"].numel()
    for op, stream, desc in [("bf16", "bf16", "bf16"), ("bf16", "f32", "bf16"), ("bf16", "f32", "f32"),
                             ("f32", "f32", "bf16"), ("f32", "f32", "bf16x2"), ("tf32", "f32", "f32"), ("f16", "f32", "f32"),
                             ("f16", "f32", "bf16x2"), ("bf16x2", "f32", "bf16x2"), ("bf16x2", "f32", "bf16")]:
        aux, fl, fr = run(sd, lq, ref, op, stream, desc)
        same_c = (aux["idx"] == base["idx"]).float().mean().item()
        ok = (aux["idx"] == base["idx"]).view(-1, 1, 1).expand_as(aux["index"])
        same_f = ((aux["index"] == base["index"]) & ok).float().mean().item()
        e3 = (fl[-1] - bl[-1]).abs().max().item() / bl[-1].abs().max().item()
        bc = base["corr"]
        m = aux["index"].reshape(bc.shape[0], 64)
        gap = bc.max(-1).values - bc.gather(2, m.unsqueeze(-1)).squeeze(-1)          # oracle-corr lost by our choice
        okb = (aux["idx"] == base["idx"]).view(-1, 1).expand_as(gap)
        big = ((gap > 1e-4) & okb).sum().item()
        werr = max((a_ - b_).abs().max().item() for a_, b_ in zip(aux["warps"], base["warps"]))
        wmean = sum((a_ - b_).abs().mean().item() for a_, b_ in zip(aux["warps"], base["warps"])) / 4
        print(f"    flips with oracle gap > 1e-4: {big}; max gap {gap[okb].max().item():.2e}; warp max err {werr:.3e} mean {wmean:.3e}")
        print(f"op={op:7s} stream={stream:5s} desc={desc:7s} coarse agree {same_c:.4f}  fine agree {same_f:.4f} "
              f"({int(round((1 - same_f) * n))}/{n} flips)  deep-feature rel err {e3:.2e}", flush=True)


if __name__ == "__main__":
    main()
