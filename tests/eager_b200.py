"""PyTorch-eager-on-B200 baseline (SURVEY.md §8(d) "Also report"): the fp32 restatement of the reference modules
(``oracle/``, pinned to the unmodified reference by tests/golden) run with stock cuDNN / cuBLAS kernels on ``cuda:0``
-- the only "GPU implementation" the reference has, since it ships no kernels of its own.  Test infrastructure: it
is a measurement of the checker, never part of the product path.

    python -m tests.eager_b200 [--batch 4] [--size 512] [--steps 5] [--train]

Prints one JSON line per mode (fp32, tf32, bf16 autocast; forward, and forward+backward with --train).  The oracle's
MASA path is the closed form (direct gathers), i.e. cheaper than the reference's unfold / fold, so these figures are an
upper bound on what the stock reference reaches on this GPU.
"""
import argparse
import json
import sys

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--train", action="store_true")
    ap.add_argument("--skip-forward", action="store_true")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "eager", "unavailable": "no CUDA device"}))
        return 0
    sys.path.insert(0, ".")
    import bench
    from oracle import restormer as O, weights as W
    from textualdegremoval_b200.archs.restormer_b200_arch import RestormerRefFusion

    dev = torch.device("cuda:0")
    shapes = {k: v.shape for k, v in RestormerRefFusion(**bench.OPTION_003).state_dict().items()}
    sd = {k: v.to(dev) for k, v in W.seeded_state_dict(shapes, 0).items()}
    lq, ref, gt = (t.to(dev) for t in bench.synth_inputs(args.batch, args.size, 100))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def fwd():
        with torch.no_grad():
            return O.restormer_ref_fusion_forward(sd, lq, ref)

    def fwd_bwd():
        for v in sd.values():
            v.grad = None
        out = O.restormer_ref_fusion_forward(sd, lq, ref)
        (out - gt).abs().mean().backward()
        return out

    def timed(fn, label, mode):
        try:
            _timed(fn, label, mode)
        except torch.OutOfMemoryError:
            for v in sd.values():
                v.grad = None
            torch.cuda.empty_cache()
            print(json.dumps({"impl": "eager", "what": label, "mode": mode, "batch": args.batch, "size": args.size,
                              "oom": True, "device_mem_gb": torch.cuda.get_device_properties(0).total_memory / 2 ** 30}),
                  flush=True)

    def _timed(fn, label, mode):
        for _ in range(args.warmup):
            fn()
        ms = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        med = ms[len(ms) // 2]
        print(json.dumps({"impl": "eager", "what": label, "mode": mode, "batch": args.batch, "size": args.size,
                          "ms_per_step": med, "img_per_s": args.batch * 1000.0 / med, "steps": args.steps,
                          "torch": torch.__version__, "device": torch.cuda.get_device_name(0),
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)

    modes = [("fp32", False, None), ("tf32", True, None), ("bf16-autocast", True, torch.bfloat16)]
    for name, tf32, amp in ([] if args.skip_forward else modes):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        ctx = torch.autocast("cuda", dtype=amp) if amp is not None else torch.autocast("cuda", enabled=False)
        with ctx:
            timed(fwd, "forward", name)
    if args.train:
        for v in sd.values():
            v.requires_grad_(True)
        for name, tf32, amp in modes[1:]:
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            ctx = torch.autocast("cuda", dtype=amp) if amp is not None else torch.autocast("cuda", enabled=False)
            with ctx:
                timed(fwd_bwd, "forward+backward (L1)", name)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
