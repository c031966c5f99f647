"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Every fixture holds: the ctor kwargs (json), the seed (weights and inputs are regenerated from
``oracle.weights``), and the reference module's fp32 outputs.  The reference modules are imported
through ``oracle.ref_loader`` (stubbed package inits, B1 index shim -- see its docstring).
"""
import json
import os

import numpy as np
import torch

from . import ref_loader as R
from . import weights as W

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

RESTORMER_CASES = {
    "restormer_withbias": dict(cfg=dict(dim=16, num_blocks=[1, 2, 1, 1], num_refinement_blocks=1,
                                        heads=[1, 2, 4, 8], LayerNorm_type="WithBias", bias=False),
                               seed=11, shape=(2, 3, 64, 64)),
    "restormer_biasfree": dict(cfg=dict(dim=24, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1,
                                        heads=[1, 2, 4, 8], LayerNorm_type="BiasFree", bias=False),
                               seed=12, shape=(1, 3, 64, 96)),
    "restormer_gray_bias": dict(cfg=dict(inp_channels=1, out_channels=1, dim=16, num_blocks=[1, 1, 1, 1],
                                         num_refinement_blocks=1, heads=[1, 2, 4, 8], LayerNorm_type="BiasFree",
                                         bias=True),
                                seed=13, shape=(1, 1, 64, 64)),
}
GUIDED_CASES = {
    "guided_restormer_128": dict(cfg=dict(dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1,
                                          heads=[1, 2, 4, 8], nf=16, ext_n_blocks=[1, 1, 1, 1],
                                          reffusion_n_blocks=[1, 2, 1, 1], LayerNorm_type="WithBias"),
                                 seed=21, lq=(2, 3, 128, 128), ref=(2, 3, 128, 128)),
    "guided_restormer_ragged": dict(cfg=dict(dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1,
                                             heads=[1, 2, 4, 8], nf=16, ext_n_blocks=[2, 1, 1, 1],
                                             reffusion_n_blocks=[1, 1, 1, 1], LayerNorm_type="WithBias"),
                                    seed=22, lq=(1, 3, 120, 130), ref=(1, 3, 128, 192)),
}


NAFNET_CASES = {
    # BASELINE.json configs[0]: NAFNet-tiny Gaussian gray denoise sigma=15, one 64x64 tile
    "nafnet_tiny_gray64": dict(cfg=dict(img_channel=1, width=16, middle_blk_num=1, enc_blk_nums=[1, 1, 1, 1],
                                        dec_blk_nums=[1, 1, 1, 1]), seed=31, shape=(1, 1, 64, 64), sigma=15),
    "nafnet_rgb_ragged": dict(cfg=dict(img_channel=3, width=16, middle_blk_num=2, enc_blk_nums=[1, 2],
                                       dec_blk_nums=[1, 1]), seed=32, shape=(2, 3, 50, 70), sigma=25),
}
NAF_GUIDED_CASES = {
    "guided_nafnet_256": dict(cfg=dict(img_channel=3, width=16, middle_blk_num=1, enc_blk_nums=[1, 1, 1, 2],
                                       dec_blk_nums=[1, 1, 1, 1], nf=16, ext_n_blocks=[1, 1, 1, 1],
                                       reffusion_n_blocks=[1, 1, 1, 1, 1]),
                              seed=41, lq=(1, 3, 256, 256), ref=(1, 3, 256, 256)),
}


def denoise_inputs(case):
    """The reference's deterministic test-noise recipe (data/restoration_dataset.py:479-480): np.random.seed(0) then
    N(0, (sigma/255)^2) added to a seeded clean tile."""
    gt = W.seeded_image("gt", case["shape"], case["seed"])
    np.random.seed(0)
    noise = np.random.normal(0, case["sigma"] / 255.0, tuple(case["shape"])).astype(np.float32)
    return gt + torch.from_numpy(noise), gt


def guided_inputs(case):
    """lq = seeded image; ref = a shifted/noised copy when shapes agree (so matching is non-degenerate)."""
    lq = W.seeded_image("lq", case["lq"], case["seed"])
    if tuple(case["ref"]) == tuple(case["lq"]):
        ref = torch.roll(lq, shifts=(5, -7), dims=(2, 3)) + 0.02 * (W.seeded_image("n", case["ref"], case["seed"]) - 0.5)
    else:
        ref = W.seeded_image("ref", case["ref"], case["seed"])
    return lq, ref


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    for name, case in RESTORMER_CASES.items():
        net = R.restormer(**case["cfg"])
        W.load_seeded(net, case["seed"])
        x = W.seeded_image("x", case["shape"], case["seed"])
        y = net(x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(case), out=y.numpy())
        print(name, tuple(y.shape), float(y.abs().max()))
    for name, case in GUIDED_CASES.items():
        net = R.restormer_ref_fusion(**case["cfg"])
        W.load_seeded(net, case["seed"])
        lq, ref = guided_inputs(case)
        y = net(lq, ref)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(case), out=y.numpy())
        print(name, tuple(y.shape), float(y.abs().max()))


def main_nafnet():
    torch.set_grad_enabled(False)
    for name, case in NAFNET_CASES.items():
        net = R.nafnet(**case["cfg"])
        W.load_seeded(net, case["seed"])
        lq, _ = denoise_inputs(case)
        y = net(lq)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(case), out=y.numpy())
        print(name, tuple(y.shape), float(y.abs().max()))
    for name, case in NAF_GUIDED_CASES.items():
        net = R.nafnet_ref_fusion(**case["cfg"])
        W.load_seeded(net, case["seed"])
        lq, ref = guided_inputs(case)
        y = net(lq, ref)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(case), out=y.numpy())
        print(name, tuple(y.shape), float(y.abs().max()))


if __name__ == "__main__":
    main()
    main_nafnet()
