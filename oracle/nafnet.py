"""TEST INFRASTRUCTURE ONLY -- fp32 CPU restatement of NAFNet / NAFNetRefFusion.

Reference: /root/reference/models/archs/network_nafnet_guided_arch.py
  SimpleGate :170-175   NAFBlock :178-238   NAFResFuseBlock :241-302 (same math)   NAFNet :305-386
  NAFNetRefFusion :389-740 (5-level MASA encoder :110-143, padder 16)
and LayerNorm2d nafnet_arch_utils.py:264-300 (eps 1e-6).  The MASA path is shared with oracle/restormer.py.
Pinned against the unmodified reference by oracle/make_golden.py (fixtures nafnet_*.npz).
"""
import torch
import torch.nn.functional as F

from .restormer import _conv, masa_encoder, masa_warp, pad_to


def layernorm2d(x, w, b, eps=1e-6):
    mu = x.mean(1, keepdim=True)
    var = (x - mu).pow(2).mean(1, keepdim=True)
    return (x - mu) / (var + eps).sqrt() * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def naf_block(sd, p, inp):
    x = layernorm2d(inp, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    x = _conv(sd, p + ".conv1", x)
    x = _conv(sd, p + ".conv2", x, padding=1, groups=x.shape[1])
    x1, x2 = x.chunk(2, 1)
    x = x1 * x2
    x = x * _conv(sd, p + ".sca.1", x.mean((2, 3), keepdim=True))
    x = _conv(sd, p + ".conv3", x)
    y = inp + x * sd[p + ".beta"]
    x = _conv(sd, p + ".conv4", layernorm2d(y, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"]))
    x1, x2 = x.chunk(2, 1)
    x = _conv(sd, p + ".conv5", x1 * x2)
    return y + x * sd[p + ".gamma"]


def _stack(sd, p, x):
    i = 0
    while f"{p}.{i}.conv1.weight" in sd:
        x = naf_block(sd, f"{p}.{i}", x)
        i += 1
    return x


def _count(sd, prefix):
    n = 0
    while any(k.startswith(f"{prefix}.{n}.") for k in sd):
        n += 1
    return n


def _unet(sd, x, inp, warps=None):
    n_enc = _count(sd, "encoders")
    encs = []
    for i in range(n_enc):
        if warps is not None:
            c = x.shape[1]
            x = _stack(sd, f"masa_blk_enc.{i}", torch.cat([x, warps[i]], 1))[:, :c]
        x = _stack(sd, f"encoders.{i}", x)
        encs.append(x)
        x = _conv(sd, f"downs.{i}", x, stride=2)
    if warps is not None:
        c = x.shape[1]
        x = _stack(sd, "masa_blk_middle.0", torch.cat([x, warps[-1]], 1))[:, :c]
    x = _stack(sd, "middle_blks", x)
    for i, skip in enumerate(encs[::-1]):
        x = F.pixel_shuffle(F.conv2d(x, sd[f"ups.{i}.0.weight"]), 2) + skip
        x = _stack(sd, f"decoders.{i}", x)
    return _conv(sd, "ending", x, padding=1) + inp


def nafnet_forward(sd, inp):
    """NAFNet.forward (:356-379)."""
    _, _, h, w = inp.shape
    x0 = pad_to(inp, 2 ** _count(sd, "encoders"))
    return _unet(sd, _conv(sd, "intro", x0, padding=1), x0)[:, :, :h, :w]


def nafnet_ref_fusion_forward(sd, inp, ref, lr_block_size=8, ref_down_block_size=1.5, dilations=(1, 2, 3),
                              return_aux=False):
    """NAFNetRefFusion.forward (:587-740)."""
    _, _, oh, ow = inp.shape
    padder = 2 ** _count(sd, "encoders")
    inp = pad_to(inp, padder * lr_block_size)
    ref = pad_to(ref, padder * lr_block_size)
    _, _, h, w = inp.shape
    _, _, hr, wr = ref.shape
    f_lq, f_ref = masa_encoder(sd, inp), masa_encoder(sd, ref)
    res = masa_warp(f_lq[-1], f_ref, padder, lr_block_size, ref_down_block_size, dilations, h, w, hr, wr,
                    return_aux=return_aux)
    warps, aux = res if return_aux else (res, None)
    out = _unet(sd, _conv(sd, "intro", inp, padding=1), inp, warps)[:, :, :oh, :ow]
    return (out, aux) if return_aux else out
