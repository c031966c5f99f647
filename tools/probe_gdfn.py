"""Quick CUDA-event timing of the fused GDFN tail vs the two-kernel path at the benchmark's shapes (B200 only).
    python tools/probe_gdfn.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from textualdegremoval_b200 import ops
DEV = "cuda"
F16 = torch.float16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1e3


SHAPES = ((4, 512, 512, 256, 96), (4, 256, 256, 256, 96), (4, 512, 512, 128, 48), (4, 128, 128, 512, 192))
if len(sys.argv) > 1 and sys.argv[1] == "one":          # single fused launch for ncu
    B, H, W, hp, C = SHAPES[0]
    hid = torch.randn(B, H, W, 2 * hp, device=DEV).to(F16)
    w9 = torch.randn(9, 2 * hp, device=DEV) * 0.3
    wo = (torch.randn(1, C, hp, device=DEV) / hp ** 0.5).to(F16)
    x = torch.randn(B, H, W, C, device=DEV)
    for _ in range(3):
        ops.gdfn_tail(hid, w9, None, wo, C, res2=x, out=x)
    torch.cuda.synchronize()
    sys.exit(0)
for (B, H, W, hp, C) in SHAPES:
    hid = torch.randn(B, H, W, 2 * hp, device=DEV).to(F16)
    w9 = torch.randn(9, 2 * hp, device=DEV) * 0.3
    wo = (torch.randn(1, C, hp, device=DEV) / hp ** 0.5).to(F16)
    x = torch.randn(B, H, W, C, device=DEV)
    t_f = timeit(lambda: ops.gdfn_tail(hid, w9, None, wo, C, res2=x, out=x))

    def two():
        g = ops.dwconv3x3(hid, w9, None, gate=1)
        ops.conv_gemm(g, wo, C, res2=x, out_f32=x)
    t_2 = timeit(two)
    nbytes = B * H * W * (2 * hp * 2 + C * 8)
    print(f"hp{hp} C{C} {H}x{W}: fused {t_f:8.1f} us ({nbytes / t_f / 1e3:7.1f} GB/s)   two kernels {t_2:8.1f} us", flush=True)
