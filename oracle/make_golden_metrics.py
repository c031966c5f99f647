"""Golden vectors for the validation PSNR: runs the UNMODIFIED reference tensor2img + calculate_psnr (this container
only; `skimage`, imported but unused by metrics/psnr_ssim.py, is stubbed) and stores inputs as seeds plus the reference
doubles in tests/golden/psnr.npz.

    python -m oracle.make_golden_metrics
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "psnr.npz")
CASES = [dict(c=3, h=24, w=40, crop=0, kind="noisy"), dict(c=3, h=33, w=17, crop=4, kind="noisy"),
         dict(c=1, h=20, w=20, crop=0, kind="noisy"), dict(c=3, h=16, w=16, crop=0, kind="equal"),
         dict(c=3, h=16, w=16, crop=2, kind="dark"), dict(c=3, h=32, w=32, crop=0, kind="overshoot"),
         dict(c=3, h=19, w=23, crop=1, kind="ties")]


def make_pair(case, seed):
    """(result, gt) fp32 CHW; regenerated from the seed by the tests."""
    g = torch.Generator().manual_seed(seed)
    c, h, w = case["c"], case["h"], case["w"]
    gt = torch.rand(c, h, w, generator=g)
    if case["kind"] == "equal":
        return gt.clone(), gt
    if case["kind"] == "dark":                      # img1.max() <= 1 in uint8 -> max_value 1 (psnr_ssim.py:58)
        return torch.rand(c, h, w, generator=g) * (1.4 / 255), torch.rand(c, h, w, generator=g) * (3.0 / 255)
    if case["kind"] == "overshoot":                 # values outside [0, 1] are clamped by tensor2img
        return gt + torch.randn(c, h, w, generator=g) * 0.5, gt * 1.5 - 0.2
    if case["kind"] == "ties":                      # exact .5 after * 255: round half to even
        k = torch.randint(0, 255, (c, h, w), generator=g).float()
        return (k + 0.5) / 255.0, k / 255.0
    return gt + torch.randn(c, h, w, generator=g) * 0.03, gt


def main():
    sys.path.insert(0, ref_loader.REF_ROOT)
    sk = types.ModuleType("skimage"); skm = types.ModuleType("skimage.metrics"); sk.metrics = skm
    sys.modules.setdefault("skimage", sk); sys.modules.setdefault("skimage.metrics", skm)
    for pkg in ("metrics", "utils"):                # bypass the package __init__ files (niqe / logger imports)
        m = types.ModuleType(pkg); m.__path__ = [os.path.join(ref_loader.REF_ROOT, pkg)]
        sys.modules[pkg] = m
    import importlib
    U = importlib.import_module("utils.utils_image")
    P = importlib.import_module("metrics.psnr_ssim")
    vals = []
    for i, case in enumerate(CASES):
        res, gt = make_pair(case, 500 + i)
        sr_img = U.tensor2img([res.unsqueeze(0)], rgb2bgr=True)
        gt_img = U.tensor2img([gt.unsqueeze(0)], rgb2bgr=True)
        vals.append(P.calculate_psnr(sr_img, gt_img, crop_border=case["crop"], test_y_channel=False))
    np.savez(OUT, psnr=np.array(vals, dtype=np.float64))
    print("wrote", OUT, vals)


if __name__ == "__main__":
    main()
