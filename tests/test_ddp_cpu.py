"""CPU suite, world_size 2 over gloo (127.0.0.1): the N > 1 host logic of the DDP step -- flat parameter groups split on
"masa" (image_restoration_ref_model.py:149-158), bucketed all-reduce, and the equivalence the reference's DDP provides:
the averaged gradient equals the gradient of the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class _Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.body = nn.Conv2d(3, 8, 3, padding=1)
        self.masa_enc = nn.Conv2d(3, 8, 3, padding=1)
        self.head = nn.Conv2d(8, 3, 1)

    def forward(self, x):
        return self.head(torch.relu(self.body(x)) + self.masa_enc(x))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from textualdegremoval_b200.ddp import DDPStep
    from textualdegremoval_b200.lib import TdrError
    torch.manual_seed(0)
    net = _Tiny()
    eng = DDPStep(net.named_parameters(), lr=2e-4, ref_lr=1e-4, bucket_bytes=256)      # tiny buckets -> many of them
    assert len(eng.groups) == 2 and eng.groups[1].lr == 1e-4
    assert sum(g.n_params for g in eng.groups) == sum(p.numel() for p in net.parameters())
    assert all(p.data_ptr() % 64 == 0 and p.grad.data_ptr() % 64 == 0 for p in net.parameters())
    assert len(eng.buckets()) > 4
    g = torch.Generator().manual_seed(1)
    x, y = torch.rand(4, 3, 8, 8, generator=g), torch.rand(4, 3, 8, 8, generator=g)
    xs, ys = x[rank * 2:(rank + 1) * 2], y[rank * 2:(rank + 1) * 2]
    eng.zero_grad()
    loss = (net(xs) - ys).abs().mean()
    loss.backward()
    for p in net.parameters():                 # autograd accumulated INTO the flat views
        assert p.grad.data_ptr() >= min(gr.grad.data_ptr() for gr in eng.groups)
    for w in eng.all_reduce_gradients():
        w.wait()
    eng.reduce_loss_async(loss)
    avg = torch.cat([p.grad.reshape(-1) for gr in eng.groups for p in gr.params]) / world
    if rank == 0:
        torch.manual_seed(0)
        ref = _Tiny()
        (ref(x) - y).abs().mean().backward()
        named = dict(ref.named_parameters())
        order = [n for n in named if "masa" not in n] + [n for n in named if "masa" in n]
        full = torch.cat([named[n].grad.reshape(-1) for n in order])
        out["grad_err"] = (avg - full).abs().max().item()
        out["loss"] = eng.read_loss()
        out["loss_ref"] = (ref(x) - y).abs().mean().item()
        try:
            eng.step()
            out["cpu_step"] = "ran"
        except TdrError:
            out["cpu_step"] = "refused"
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equivalence():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["grad_err"] < 1e-6
    assert abs(out["loss"] - out["loss_ref"]) < 1e-6
    assert out["cpu_step"] == "refused"          # no CPU fallback for the fused optimizer tail
