"""Time tdr_vit_attention against the materialised-score schedule on the towers' shapes (CUDA events, GPU box only)."""
import torch

from textualdegremoval_b200 import ops
from textualdegremoval_b200.archs import vit_b200 as VB


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


for name, (B, N, heads, hd) in dict(dino_518=(8, 1370, 12, 64), clip_224=(32, 257, 16, 80), dino_224=(8, 257, 12, 64)).items():
    D = heads * hd
    buf = torch.randn(B * N + 8, 3 * D, device="cuda").to(torch.bfloat16)
    qkv = buf[: B * N].view(B, 1, N, 3 * D)
    fused = timeit(lambda: ops.vit_attention(qkv, heads, hd, hd ** -0.5))
    VB._MATERIALISED = True
    mat = timeit(lambda: VB.attention(qkv, heads, D), n=5)
    VB._MATERIALISED = False
    fl = 4.0 * B * heads * N * N * hd
    print(f"{name}: fused {fused:8.1f} us ({fl / fused * 1e-6:6.1f} TFLOP/s)   materialised {mat:8.1f} us", flush=True)
