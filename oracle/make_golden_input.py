"""Golden vectors for the input pipeline: runs the UNMODIFIED reference functions (this container only) and stores
frames, the random decisions and the reference tensors in tests/golden/input_pipeline.npz.

    python -m oracle.make_golden_input
"""
import importlib.util
import os
import random
import sys
import types

import numpy as np

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "input_pipeline.npz")


def _load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ref_loader.REF_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    try:
        import torchvision  # noqa: F401
    except Exception:                       # utils_image only imports make_grid for visualisation helpers
        tv = types.ModuleType("torchvision"); tvu = types.ModuleType("torchvision.utils")
        tvu.make_grid = None; tv.utils = tvu
        sys.modules.setdefault("torchvision", tv); sys.modules.setdefault("torchvision.utils", tvu)
    T = _load("data/transforms.py", "ref_transforms")
    U = _load("utils/utils_image.py", "ref_utils_image")
    import torch
    rng = np.random.RandomState(0)
    cases = [dict(h=24, w=30, size=16), dict(h=17, w=23, size=16), dict(h=10, w=26, size=16),      # h < size: padding
             dict(h=18, w=7, size=16), dict(h=20, w=20, size=20), dict(h=5, w=4, size=16)]         # tiny: multi-reflection
    store = {}
    k = 0
    for ci, c in enumerate(cases):
        for rep in range(4):
            gt = rng.randint(0, 256, (c["h"], c["w"], 3)).astype(np.uint8)
            lq = rng.randint(0, 256, (c["h"], c["w"], 3)).astype(np.uint8)
            seed = 1000 + k
            # the reference's own draws, replayed to record the decisions
            random.seed(seed)
            ph, pw = max(c["h"], c["size"]), max(c["w"], c["size"])
            top = random.randint(0, ph - c["size"]); left = random.randint(0, pw - c["size"]); mode = random.randint(0, 7)
            random.seed(seed)
            img_gt = gt.astype(np.float32) / 255.          # imfrombytes(float32=True) after cv2.imdecode
            img_lq = lq.astype(np.float32) / 255.
            img_gt, img_lq = U.padding(img_gt, img_lq, c["size"])
            img_gt, img_lq = T.paired_random_crop(img_gt, img_lq, c["size"], 1, "gt_path")
            img_gt, img_lq = T.random_augmentation(img_gt, img_lq)
            t_gt, t_lq = U.img2tensor([img_gt, img_lq], bgr2rgb=True, float32=True)
            if rep == 3:                                   # normalize branch (:240-244)
                mean, std = [0.5, 0.4, 0.3], [0.25, 0.5, 0.2]
                from torchvision.transforms.functional import normalize
                normalize(t_gt, mean, std, inplace=True); normalize(t_lq, mean, std, inplace=True)
            else:
                mean = std = None
            store[f"s{k}_gt_frame"] = gt; store[f"s{k}_lq_frame"] = lq
            store[f"s{k}_dec"] = np.array([top, left, mode, c["size"], 1 if mean else 0], dtype=np.int32)
            store[f"s{k}_gt"] = t_gt.numpy(); store[f"s{k}_lq"] = t_lq.numpy()
            k += 1
    store["n"] = np.array(k)
    store["mean"] = np.array([0.5, 0.4, 0.3], dtype=np.float32); store["std"] = np.array([0.25, 0.5, 0.2], dtype=np.float32)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, k, "samples", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
