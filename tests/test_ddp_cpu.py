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
    torch.manual_seed(100 + rank)              # the reference seeds every rank differently (manual_seed + rank)
    net = _Tiny()
    net.register_buffer("running", torch.full((3,), float(rank)))
    eng = DDPStep(net.named_parameters(), lr=2e-4, ref_lr=1e-4, bucket_bytes=256,      # tiny buckets -> many of them
                  buffers=list(net.buffers()))
    # construction broadcast rank 0's parameters and buffers (what DistributedDataParallel does, base_model.py:76-82)
    flat = torch.cat([g.flat for g in eng.groups])
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    out[f"param_spread_{rank}"] = (both[0] - both[1]).abs().max().item()
    out[f"buffer_{rank}"] = net.running.tolist()
    torch.manual_seed(100)
    init = _Tiny()
    out[f"is_rank0_init_{rank}"] = all(torch.equal(a, b) for a, b in zip(net.parameters(), init.parameters()))
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
    # overlap bookkeeping: modules reported in backward order (head -> masa_enc -> body); slices go out as the finished
    # tail of each flat buffer grows, the rest in all_reduce_gradients(); every element is reduced exactly once
    eng.begin_backward()
    eng.mark_done(net.head)
    n_early_head = len(eng._early)
    eng.mark_done(net.masa_enc)
    eng.mark_done(net.body.bias)                 # a single parameter
    n_early = len(eng._early)
    works = eng.all_reduce_gradients()
    out[f"early_{rank}"] = (n_early_head, n_early, len(works))
    for w in works:
        w.wait()
    eng.reduce_loss_async(loss)
    avg = torch.cat([p.grad.reshape(-1) for gr in eng.groups for p in gr.params]) / world
    if rank == 0:
        torch.manual_seed(100)
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
    for r in (0, 1):
        e_head, e_all, total = out[f"early_{r}"]
        # the head alone is smaller than a bucket (nothing sent yet); with masa_enc done the ref group's tail goes out
        assert e_head == 0 and e_all >= 1 and total > e_all, out[f"early_{r}"]
        assert out[f"param_spread_{r}"] == 0.0 and out[f"is_rank0_init_{r}"], "ranks must start from rank 0's parameters"
        assert out[f"buffer_{r}"] == [0.0, 0.0, 0.0], "buffers must be broadcast from rank 0"


def test_optimizer_face_drives_lr_and_checkpoints():
    """FlatAdamW is a torch.optim.Optimizer: torch LR schedulers accept it and write the lr the fused step uses; its
    state_dict round-trips the Adam moments, per-group step counts and the EMA copy."""
    from textualdegremoval_b200.ddp import DDPStep, FlatAdamW
    net = _Tiny()
    eng = DDPStep(net.named_parameters(), lr=2e-4, ref_lr=1e-4, ema_decay=0.999)
    opt = FlatAdamW(eng)
    assert [pg["lr"] for pg in opt.param_groups] == [2e-4, 1e-4] and [pg["tag"] for pg in opt.param_groups] == ["normal", "ref"]
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10)
    for pg in opt.param_groups:
        pg["lr"] *= 0.5                             # what update_learning_rate / a scheduler step does
    opt.sync_lr()
    assert [g.lr for g in eng.groups] == [1e-4, 5e-5]
    eng.groups[0].m.fill_(0.25); eng.groups[1].v.fill_(0.5); eng.groups[0].steps = 7; eng.step_count = 9
    sd = opt.state_dict()
    net2 = _Tiny()
    eng2 = DDPStep(net2.named_parameters(), lr=1.0, ref_lr=1.0, ema_decay=0.999)
    opt2 = FlatAdamW(eng2)
    opt2.load_state_dict(sd)
    assert eng2.groups[0].steps == 7 and eng2.step_count == 9 and [g.lr for g in eng2.groups] == [1e-4, 5e-5]
    assert torch.equal(eng2.groups[0].m, eng.groups[0].m) and torch.equal(eng2.groups[1].v, eng.groups[1].v)
    assert torch.equal(eng2.ema[0], eng.ema[0])
    assert sched is not None
