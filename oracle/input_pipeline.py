"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the per-sample tensor preparation of the reference's paired dataset.

Reference (Dataset_PairedImageWithRef.__getitem__, /root/reference/data/restoration_dataset.py:194-253):
  imfrombytes(float32=True)   utils/utils_image.py:194-218     uint8 -> float32 / 255.
  padding                     utils/utils_image.py:243-254     cv2.copyMakeBorder(0, h_pad, 0, w_pad, BORDER_REFLECT)
  paired_random_crop          data/transforms.py:24-83         top / left from random.randint, scale 1
  random_augmentation         data/transforms.py:223-275       mode = random.randint(0, 7): np.rot90 / np.flipud
  img2tensor(bgr2rgb=True)    utils/utils_image.py:102-126     BGR -> RGB, HWC -> CHW
  normalize(mean, std)        restoration_dataset.py:240-244   (x - mean) / std, float32

Pinned to the unmodified reference functions by tests/golden/input_pipeline.npz (oracle/make_golden_input.py).
"""
import numpy as np


def reflect_pad(img, size_h, size_w):
    """cv2.BORDER_REFLECT (fedcba|abcdefgh|hgfedcb) at the bottom / right up to (size_h, size_w) (:243-254)."""
    h, w = img.shape[:2]
    ph, pw = max(h, size_h), max(w, size_w)

    def idx(n_out, n):
        i = np.arange(n_out) % (2 * n)
        return np.where(i < n, i, 2 * n - 1 - i)

    return img[idx(ph, h)][:, idx(pw, w)]


def augment(img, mode):
    """data_augmentation (data/transforms.py:223-268)."""
    if mode == 0:
        return img
    if mode == 1:
        return np.flipud(img)
    out = np.rot90(img, k={2: 1, 3: 1, 4: 2, 5: 2, 6: 3, 7: 3}[mode])
    return np.flipud(out) if mode in (3, 5, 7) else out


def prepare_patch(frame_u8, top, left, mode, size, bgr2rgb=True, mean=None, std=None):
    """One sample: uint8 HWC (BGR) frame -> float32 CHW patch."""
    size_h, size_w = (size, size) if isinstance(size, int) else size
    img = frame_u8.astype(np.float32) / 255.
    img = reflect_pad(img, size_h, size_w)
    img = img[top: top + size_h, left: left + size_w]
    assert img.shape[:2] == (size_h, size_w), "crop outside the (padded) frame"
    img = augment(img, mode)
    if img.shape[2] == 3 and bgr2rgb:
        img = img[:, :, ::-1]
    out = np.ascontiguousarray(img.transpose(2, 0, 1))
    if mean is not None:
        m = np.asarray(mean, dtype=np.float32).reshape(-1, 1, 1)
        s = np.asarray(std, dtype=np.float32).reshape(-1, 1, 1)
        out = (out - m) / s
    return out.astype(np.float32)
