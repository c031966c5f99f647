"""CPU suite: sharded validation (images are independent units; eval shards them across ranks and reduces
PSNR sums).  World size 2 over gloo must give the same dataset mean as one process scoring every image, and the
single-process mean must equal the reference's sequential accumulation of calculate_psnr(tensor2img(.), tensor2img(.))."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import metrics as M


def _dataset(n=7):
    g = torch.Generator().manual_seed(4)
    items = []
    for i in range(n):
        h, w = 16 + 2 * i, 24 - i
        gt = torch.rand(3, h, w, generator=g)
        items.append(dict(gt=gt, lq=gt + torch.randn(3, h, w, generator=g) * 0.05, ref=gt.flip(-1)))
    return items


class _Net(torch.nn.Module):                      # stands in for net_g(lq, ref): a fixed "restoration"
    def forward(self, lq, ref):
        return [lq * 0.9 + 0.05, (lq * 0.98 + 0.01).clamp(-0.2, 1.3)]       # list: last = output


def _psnr_cpu(result, gt, crop_border):           # the oracle stands in for the CUDA op on the CPU
    return [M.psnr(result[0].numpy(), gt[0].numpy(), crop_border)]


def _expected(crop):
    net, tot = _Net(), 0.0
    items = _dataset()
    for it in items:                              # the reference's loop: metric_results[name] += calculate_psnr(...)
        out = net(it["lq"][None], it["ref"][None])[-1]
        tot += M.psnr(out[0].numpy(), it["gt"].numpy(), crop)
    return tot / len(items)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from textualdegremoval_b200.validation import shard_indices, validate_psnr
    assert shard_indices(7, rank, world) == list(range(rank, 7, world))
    mean, n = validate_psnr(_Net(), _dataset(), crop_border=2, psnr_fn=_psnr_cpu)
    out[rank] = (mean, n)
    dist.destroy_process_group()


def test_sharded_validation_world2_matches_single_process():
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    want = _expected(2)
    for rank in (0, 1):                           # every rank holds the dataset mean (the reference: rank 0 only)
        mean, n = out[rank]
        assert n == 7
        assert abs(mean - want) <= 1e-12 * abs(want)


def test_single_process_validation_equals_reference_accumulation():
    from textualdegremoval_b200.validation import validate_psnr
    mean, n = validate_psnr(_Net(), _dataset(), crop_border=0, psnr_fn=_psnr_cpu)
    assert n == 7 and mean == _expected(0)        # same order of float64 additions -> identical double
    assert np.isfinite(mean)
