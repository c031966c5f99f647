"""DDP training-step tail: flat gradient buffers, bucketed all-reduce, global-norm clip, AdamW over
the reference's two LR groups, EMA, asynchronous loss reduction.

Reference behaviour being replaced (all in /root/reference):
  * ``models/base_model.py:76-82``  -- ``DistributedDataParallel`` (c10d 25 MB buckets, mean all-reduce of fp32 grads);
  * ``models/image_restoration_ref_model.py:149-181`` -- two optimizer groups split on the substring ``"masa"`` in the
    parameter name (``lr`` / ``ref_lr``), AdamW with ``weight_decay`` and ``betas``;
  * ``:276-281`` -- ``clip_grad_norm_(net_g.parameters(), 0.01)`` when ``use_grad_clip``, ``optimizer.step()``,
    ``reduce_loss_dict`` (``base_model.py:353-378``: ``dist.reduce`` to rank 0 + ``.item()`` EVERY step);
  * ``base_model.py:54-62`` -- EMA of the parameters.

Here: every parameter (and its ``.grad``) of a group is a view into ONE contiguous fp32 buffer, so
  - the gradient exchange is ``torch.distributed.all_reduce`` (NCCL over NVLink on the B200 box, gloo in the CPU tests)
    over fixed 25 MB slices of that buffer, issued asynchronously on the process group's stream;
  - clip + AdamW (+ EMA) are three flat CUDA kernels (csrc/tdr_optim.cu) that read the clip coefficient from device
    memory -- the step never synchronises with the host; the loss is reduced with an async ``dist.reduce`` and only
    read when the caller asks (``print_freq``).
The 1/world averaging is folded into the optimizer kernel (``grad_scale``) instead of a separate pass.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import lib

F32 = torch.float32


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class FlatGroup:
    """One optimizer group whose parameters / gradients are views into flat buffers."""

    ALIGN = 16     # floats: every parameter starts on a 64 B boundary (the kernels read weights / biases as float4)

    def __init__(self, params, lr):
        self.params = list(params)
        self.lr = lr
        self.tag = "normal"
        self.steps = 0                                   # updates applied to THIS group (AdamW bias correction)
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.n = n                                       # padded length (padding stays zero: p = g = m = v = 0)
        self.n_params = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(max(n, 4), dtype=F32, device=dev)
        self.grad = torch.zeros(max(n, 4), dtype=F32, device=dev)
        self.offsets = offs
        for p, off in zip(self.params, offs):
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.grad[off:off + k].view_as(p)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)

    def views(self, flat):
        """Per-parameter views of a flat buffer laid out like ``self.flat`` (e.g. the EMA copy)."""
        return [flat[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)]


def split_param_groups(named_params, lr, ref_lr):
    """The reference's grouping rule (image_restoration_ref_model.py:149-158): names containing ``masa`` -> ref_lr."""
    normal, ref = [], []
    for name, p in named_params:
        (ref if "masa" in name else normal).append(p)
    groups = []
    for tag, ps, rate in (("normal", normal, lr), ("ref", ref, ref_lr)):
        if ps:
            g = FlatGroup(ps, rate)
            g.tag = tag
            groups.append(g)
    return groups


class DDPStep:
    """all-reduce -> clip -> AdamW -> (EMA).  ``process_group=None`` and world size 1 make it a single-GPU step."""

    def __init__(self, named_params, lr, ref_lr, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=0.01,
                 use_grad_clip=True, bucket_bytes=25 * 1024 * 1024, ema_decay=0.0, process_group=None, buffers=()):
        """buffers: the module's buffers (``net_g.buffers()``); with world > 1 they and every parameter are broadcast
        from rank 0 at construction -- what ``DistributedDataParallel`` does when the reference wraps ``net_g``
        (models/base_model.py:76-82); the reference seeds each rank with ``manual_seed + rank``
        (main_train_restoration_with_ref_input.py:55), so without it every rank would train different weights."""
        self.groups = split_param_groups(list(named_params), lr, ref_lr)
        self.weight_decay, self.betas, self.eps = weight_decay, betas, eps
        self.max_grad_norm, self.use_grad_clip = max_grad_norm, use_grad_clip
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.step_count = 0
        self.ema_decay = ema_decay
        if self.world > 1:
            for g in self.groups:
                dist.broadcast(g.flat, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)
            for b in buffers:
                dist.broadcast(b, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)
        self.ema = [g.flat.clone() for g in self.groups] if ema_decay > 0 else None
        self._loss = None
        self._loss_work = None
        dev = self.groups[0].flat.device
        self._on_cuda = dev.type == "cuda"
        if self._on_cuda:
            nb = lib.load().tdr_sumsq_partial_count()
            self._partials = torch.zeros(nb * len(self.groups), dtype=F32, device=dev)
            self._clip = torch.ones(2, dtype=F32, device=dev)

    # ---- gradient exchange ---------------------------------------------------------------------------------------
    def buckets(self):
        """(group index, start, end) slices of the flat gradient buffers, <= bucket_bytes each, in REVERSE order (the
        last layers' gradients are produced first by backward)."""
        out = []
        for gi, g in enumerate(self.groups):
            for s in range(0, g.n, self.bucket_elems):
                out.append((gi, s, min(g.n, s + self.bucket_elems)))
        return out[::-1]

    # Overlap with the backward pass.  The explicit backward schedule finishes the parameters in (roughly) reverse
    # registration order, i.e. from the END of each flat buffer towards its start.  The schedule reports finished modules /
    # parameters (``mark_done``); whenever the finished TAIL of a group's buffer has grown by a bucket, that slice is
    # all-reduced at once.  ``dist.all_reduce(async_op=True)`` orders the collective after the kernels already queued on
    # the current stream and runs it on the process group's own stream, so it overlaps the rest of the backward; what is
    # left (and everything, when nothing was reported) goes out in ``all_reduce_gradients``.
    def begin_backward(self):
        self._done = set()
        self._tail = [len(g.params) - 1 for g in self.groups]      # index of the last parameter not yet sent
        self._hi = [g.n for g in self.groups]                      # everything in [hi, n) has been sent
        self._early = []

    def mark_done(self, obj):
        """obj: an nn.Module or a parameter whose gradient is final for this step."""
        if self.world == 1 or not hasattr(self, "_done"):
            return
        params = [obj] if isinstance(obj, torch.Tensor) else list(obj.parameters())
        self._done.update(id(p) for p in params)
        for gi, g in enumerate(self.groups):
            t = self._tail[gi]
            while t >= 0 and id(g.params[t]) in self._done:
                t -= 1
            lo = g.offsets[t + 1] if t + 1 < len(g.params) else g.n
            if self._hi[gi] - lo >= self.bucket_elems or (t < 0 and lo < self._hi[gi]):
                self._tail[gi] = t
                self._send(gi, lo, self._hi[gi], self._early)
                self._hi[gi] = lo

    def _send(self, gi, lo, hi, works):
        g = self.groups[gi]
        for s in range(hi, lo, -self.bucket_elems):
            a = max(lo, s - self.bucket_elems)
            works.append(dist.all_reduce(g.grad[a:s], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def all_reduce_gradients(self):
        """SUM all-reduce of every bucket not sent yet (async); averaging is folded into the optimizer kernel.  Returns
        all outstanding works of this step, the ones issued during the backward included."""
        if self.world == 1:
            return []
        if not hasattr(self, "_done"):
            self.begin_backward()
        works = self._early
        for gi in range(len(self.groups)):
            if self._hi[gi] > 0:
                self._send(gi, 0, self._hi[gi], works)
                self._hi[gi] = 0
        del self._done
        return works

    # ---- optimizer tail ------------------------------------------------------------------------------------------
    def step(self, works=(), frozen=()):
        """frozen: indices of groups whose parameters are not updated this step (the reference's ``fix_iterations``
        warm-up sets requires_grad=False on the ``masa`` parameters, image_restoration_ref_model.py:203-209: they then
        take no part in the clipped norm and receive no update)."""
        for w in works:
            w.wait()                      # stream-level wait on NCCL; does not block the host for CUDA tensors
        self.step_count += 1
        scale = 1.0 / self.world
        if not self._on_cuda:
            raise lib.TdrError("DDPStep.step: the fused clip/AdamW kernels run on CUDA tensors only")
        clip_ptr = None
        if self.use_grad_clip:
            nb = lib.load().tdr_sumsq_partial_count()
            active = [gi for gi in range(len(self.groups)) if gi not in frozen]
            for k, gi in enumerate(active):
                g = self.groups[gi]
                lib.call("tdr_sumsq_partial", C.c_void_p(g.grad.data_ptr()), g.n,
                         C.c_void_p(self._partials.data_ptr() + k * nb * 4), _stream())
            lib.call("tdr_clip_coef", C.c_void_p(self._partials.data_ptr()), nb * len(active), self.max_grad_norm,
                     scale, C.c_void_p(self._clip.data_ptr()), _stream())
            clip_ptr = C.c_void_p(self._clip.data_ptr())
        for gi, g in enumerate(self.groups):
            if gi in frozen:
                continue
            g.steps += 1          # per-group step, as torch.optim.AdamW's per-parameter state['step']: a group that was
            #                       frozen for fix_iterations starts its bias correction at 1 when it is first updated
            lib.call("tdr_adamw_step", C.c_void_p(g.flat.data_ptr()), C.c_void_p(g.grad.data_ptr()),
                     C.c_void_p(g.m.data_ptr()), C.c_void_p(g.v.data_ptr()), g.n, g.lr, self.betas[0], self.betas[1],
                     self.eps, self.weight_decay, g.steps, scale, clip_ptr, _stream())
        if self.ema is not None:
            for g, e in zip(self.groups, self.ema):
                lib.call("tdr_ema_update", C.c_void_p(e.data_ptr()), C.c_void_p(g.flat.data_ptr()), g.n, self.ema_decay,
                         _stream())
        # the kernels wrote the parameters behind autograd's back: bump the version counters so that weight caches
        # keyed on (data_ptr, _version) -- the packed bf16 GEMM operands of the arch modules -- are rebuilt
        for g in self.groups:
            for p in g.params:
                torch.autograd.graph.increment_version(p)

    def zero_grad(self):
        for g in self.groups:
            g.grad.zero_()

    # ---- loss logging (base_model.py:353-378 without the per-step .item()) ------------------------------------------
    def reduce_loss_async(self, loss):
        self._loss = loss.detach().clone().reshape(1)
        self._loss_work = dist.reduce(self._loss, dst=0, group=self.pg, async_op=True) if self.world > 1 else None

    def read_loss(self):
        """Host read (synchronises): call every print_freq iterations, not every step.  The rank-averaged value is
        meaningful on rank 0 only (``dist.reduce`` to dst 0, as base_model.py:369); other ranks get local / world."""
        if self._loss is None:
            raise lib.TdrError("DDPStep.read_loss: no step has been taken yet")
        if self._loss_work is not None:
            self._loss_work.wait()
        return float(self._loss.item()) / self.world

    # ---- checkpointing (models/base_model.py:311-351 save_training_state / resume_training) -----------------------------
    def state_dict(self):
        """Optimizer state of the flat groups: Adam moments, per-group step counts, the EMA copy."""
        return dict(step_count=self.step_count,
                    groups=[dict(tag=g.tag, lr=g.lr, steps=g.steps, m=g.m.detach().cpu().clone(), v=g.v.detach().cpu().clone())
                            for g in self.groups],
                    ema=None if self.ema is None else [e.detach().cpu().clone() for e in self.ema])

    def load_state_dict(self, sd):
        if len(sd["groups"]) != len(self.groups):
            raise lib.TdrError("DDPStep.load_state_dict: group count mismatch")
        self.step_count = int(sd["step_count"])
        for g, st in zip(self.groups, sd["groups"]):
            if st["m"].numel() != g.m.numel() or st["tag"] != g.tag:
                raise lib.TdrError(f"DDPStep.load_state_dict: group '{g.tag}' does not match the checkpoint")
            g.lr, g.steps = float(st["lr"]), int(st["steps"])
            g.m.copy_(st["m"])
            g.v.copy_(st["v"])
        if self.ema is not None and sd.get("ema") is not None:
            for e, t in zip(self.ema, sd["ema"]):
                e.copy_(t)

    def grad_norm(self):
        return float(self._clip[1].item())


class FlatAdamW(torch.optim.Optimizer):
    """``torch.optim.Optimizer`` face of a ``DDPStep`` so that the reference's model wrapper keeps working unchanged: it is
    what goes into ``self.optimizers`` (image_restoration_ref_model.py:170-181), so ``setup_schedulers`` /
    ``update_learning_rate`` (base_model.py:101-205: ``CosineAnnealingRestartCyclicLR`` + warm-up write
    ``param_groups[i]['lr']``) and ``save_training_state`` / ``resume_training`` (:311-351) act on the fused step.
    ``step()`` = gradient all-reduce + clip + AdamW (+ EMA) on the flat buffers, with the lr of each ``param_groups``
    entry; ``state_dict()`` carries the Adam moments, step counts and the EMA copy."""

    def __init__(self, engine: "DDPStep"):
        self.engine = engine
        groups = [dict(params=g.params, lr=g.lr, initial_lr=g.lr, betas=engine.betas, eps=engine.eps,
                       weight_decay=engine.weight_decay, tag=g.tag) for g in engine.groups]
        super().__init__(groups, dict(lr=engine.groups[0].lr, betas=engine.betas, eps=engine.eps,
                                      weight_decay=engine.weight_decay))

    def sync_lr(self):
        for g, pg in zip(self.engine.groups, self.param_groups):
            g.lr = float(pg["lr"])

    @torch.no_grad()
    def step(self, closure=None, frozen=()):
        self.sync_lr()
        self.engine.step(self.engine.all_reduce_gradients(), frozen=frozen)

    def zero_grad(self, set_to_none=False):
        self.engine.zero_grad()

    def state_dict(self):
        return dict(param_groups=[{k: v for k, v in pg.items() if k != "params"} for pg in self.param_groups],
                    flat=self.engine.state_dict())

    def load_state_dict(self, sd):
        for pg, st in zip(self.param_groups, sd["param_groups"]):
            pg.update({k: v for k, v in st.items() if k != "params"})
        self.engine.load_state_dict(sd["flat"])
        for pg, g in zip(self.param_groups, self.engine.groups):
            pg["lr"] = g.lr if "lr" not in pg else pg["lr"]
        self.sync_lr()


class RefGuidedTrainer:
    """The training half of the reference's ``RefGuidedImageCleanModel`` (models/image_restoration_ref_model.py) for one
    process per GPU: ``feed_train_data`` :185-213 (without the dataset / DINO crop selection, which the caller does) and
    ``optimize_parameters`` :248-284 -- zero_grad, ``net_g(lq, ref_in)``, L1 ``cri_pix``, backward, DDP gradient
    all-reduce (``base_model.py:76-82``), ``clip_grad_norm_(…, 0.01)``, AdamW over the ``masa`` / non-``masa`` LR groups,
    loss reduction, EMA.

    ``train_opt`` takes the reference's option keys: ``optim_g: {type: AdamW, lr, ref_lr, weight_decay, betas}``,
    ``use_grad_clip``, ``pixel_opt: {type: L1Loss, loss_weight}``, ``ema_decay``.  The network runs its explicit
    forward-with-tape / backward schedule (archs/restormer_train.py) and accumulates gradients straight into the flat
    buffers the all-reduce and the fused optimizer kernels work on; nothing in the step synchronises with the host.
    """

    def __init__(self, net_g, train_opt, process_group=None, net_ext=None):
        """net_ext: optional frozen DINOv2 ViT (``archs.vit_b200.vit_base``): when given and the batch carries the FULL
        reference image under ``'ref'``, the h x h reference crop most similar to ``lq`` is selected on the device every
        step (image_restoration_ref_model.py:215-247), exactly where the reference does it."""
        og = dict(train_opt.get("optim_g", {}))
        if og.get("type", "AdamW") != "AdamW":
            raise lib.TdrError(f"RefGuidedTrainer: optimizer {og.get('type')} not implemented (AdamW only, as in the "
                               "shipped option files)")
        pix = dict(train_opt.get("pixel_opt", {}))
        if pix.get("type", "L1Loss") != "L1Loss" or pix.get("reduction", "mean") != "mean":
            raise lib.TdrError("RefGuidedTrainer: only the L1Loss(mean) pixel criterion is implemented")
        self.loss_weight = float(pix.get("loss_weight", 1.0))
        self.net_g = net_g
        self.engine = DDPStep(net_g.named_parameters(), lr=og.get("lr", 3e-4), ref_lr=og.get("ref_lr", og.get("lr", 3e-4)),
                              weight_decay=og.get("weight_decay", 1e-4), betas=tuple(og.get("betas", (0.9, 0.999))),
                              use_grad_clip=bool(train_opt.get("use_grad_clip", True)),
                              ema_decay=float(train_opt.get("ema_decay", 0.0)), process_group=process_group,
                              buffers=list(net_g.buffers()))
        # the torch.optim.Optimizer face: append it to the reference wrapper's ``self.optimizers`` so
        # that its LR schedulers and training-state checkpoints drive / persist this step
        self.optimizer_g = FlatAdamW(self.engine)
        self.optimizers = [self.optimizer_g]
        net_g.grad_direct = True
        self.net_ext = net_ext
        self.fix_iterations = train_opt.get("fix_iterations")
        self._ref_group = tuple(i for i, g in enumerate(self.engine.groups) if g.tag == "ref")
        dev = self.engine.groups[0].flat.device
        self._loss = torch.zeros(1, dtype=F32, device=dev)
        self._partial = torch.zeros(lib.load().tdr_sumsq_partial_count(), dtype=F32, device=dev)
        self.lq = self.gt = self.ref_in = self.ref = self.output = None
        self.log_dict = {}

    def feed_train_data(self, data):
        dev = self.engine.groups[0].flat.device
        self.lq = data["lq"].to(dev, non_blocking=True)
        self.gt = data["gt"].to(dev, non_blocking=True) if "gt" in data else None
        self.ref_in = data["ref_in"].to(dev, non_blocking=True) if "ref_in" in data else None
        self.ref = data["ref"].to(dev, non_blocking=True) if "ref" in data else None
        if self.ref_in is None and self.ref is not None and self.net_ext is None:
            self.ref_in = self.ref             # no selector: the reference image is used as is

    def optimize_parameters(self, current_iter=0):
        from .archs.restormer_train import Grads
        net = self.net_g
        if self.net_ext is not None and self.ref is not None:
            from .archs.vit_b200 import select_reference_crop
            with torch.no_grad():
                self.ref_in, _, _ = select_reference_crop(self.net_ext, self.lq, self.ref)
        # fix_iterations: the ``masa`` group takes no update (and no part in the clipped norm) for the first iterations.
        # The reference intends this (image_restoration_ref_model.py:203-212) but never triggers it: its option files spell
        # the key ``param_fix_iterations``, so with the shipped options every parameter always trains --
        # the same happens here (the key is absent => frozen = ()).  When set, the group's AdamW state starts at its first
        # real update (per-group step counter), as torch.optim.AdamW would for parameters whose grad was None until then.
        frozen = self._ref_group if (self.fix_iterations is not None and current_iter < self.fix_iterations) else ()
        if self.gt is None:
            raise lib.TdrError("RefGuidedTrainer.optimize_parameters: feed_train_data() got no 'gt'")
        self.optimizer_g.sync_lr()
        self.engine.zero_grad()
        inputs = (self.lq,) if self.ref_in is None else (self.lq, self.ref_in)
        out, state = net._forward_train(*inputs)
        self.output = out
        dout = torch.empty_like(out)
        gt = self.gt.contiguous().float()
        if tuple(gt.shape) != tuple(out.shape):
            raise lib.TdrError(f"RefGuidedTrainer: gt shape {tuple(gt.shape)} != network output {tuple(out.shape)}")
        lib.call("tdr_l1_loss_grad", C.c_void_p(out.data_ptr()), C.c_void_p(gt.data_ptr()), out.numel(), self.loss_weight,
                 C.c_void_p(dout.data_ptr()), C.c_void_p(self._loss.data_ptr()), C.c_void_p(self._partial.data_ptr()),
                 _stream())
        self.engine.begin_backward()
        net._backward(state, dout, Grads(direct=True, on_done=self.engine.mark_done))
        works = self.engine.all_reduce_gradients()
        self.engine.step(works, frozen=frozen)
        self.engine.reduce_loss_async(self._loss)
        return self._loss

    def ema_state_dict(self):
        """``net_g_ema`` parameters (models/base_model.py:54-62, saved under 'params_ema' :213-250) as a name -> tensor
        dict of views into the flat EMA buffers; None when ema_decay == 0.  Buffers of the module are passed through."""
        if self.engine.ema is None:
            return None
        by_id = {}
        for g, e in zip(self.engine.groups, self.engine.ema):
            for p, v in zip(g.params, g.views(e)):
                by_id[id(p)] = v
        sd = {n: by_id[id(p)] for n, p in self.net_g.named_parameters()}
        for n, b in self.net_g.named_buffers():
            sd[n] = b
        return sd

    def current_loss(self):
        """Host read of the (rank-averaged) pixel loss: call every print_freq iterations (base_model.py:353-378)."""
        self.log_dict = {"l_pix": self.engine.read_loss()}
        return self.log_dict["l_pix"]
