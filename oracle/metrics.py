"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's validation PSNR (use_image: true).

Reference:
  tensor2img        /root/reference/utils/utils_image.py:129-191   clamp [0, 1], (x - 0) / (1 - 0), * 255.0, round, uint8
  calculate_psnr    /root/reference/metrics/psnr_ssim.py:9-63      crop_border, float64 mse, max_value 1 or 255
Pinned to the unmodified reference functions by tests/golden/psnr.npz (oracle/make_golden_metrics.py).
"""
import numpy as np


def to_u8(img_chw):
    """tensor2img on one fp32 CHW image (channel order is irrelevant for PSNR): uint8 HWC."""
    x = np.clip(img_chw.astype(np.float32), 0.0, 1.0)
    x = (x - np.float32(0)) / np.float32(1)
    return (x.transpose(1, 2, 0) * 255.0).round().astype(np.uint8)


def psnr_sums(a_chw, b_chw, crop_border=0):
    """(sum of squared uint8 differences as a python int, max of the first image) over the cropped window."""
    a, b = to_u8(a_chw).astype(np.int64), to_u8(b_chw).astype(np.int64)
    if crop_border:
        a = a[crop_border:-crop_border, crop_border:-crop_border]
        b = b[crop_border:-crop_border, crop_border:-crop_border]
    return int(((a - b) ** 2).sum()), int(a.max()), a.size


def psnr(a_chw, b_chw, crop_border=0):
    """calculate_psnr(tensor2img(a), tensor2img(b), crop_border, test_y_channel=False)."""
    a, b = to_u8(a_chw).astype(np.float64), to_u8(b_chw).astype(np.float64)
    if crop_border:
        a = a[crop_border:-crop_border, crop_border:-crop_border]
        b = b[crop_border:-crop_border, crop_border:-crop_border]
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return float("inf")
    max_value = 1. if a.max() <= 1 else 255.
    return float(20. * np.log10(max_value / np.sqrt(mse)))
