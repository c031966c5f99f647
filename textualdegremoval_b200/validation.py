"""Sharded validation PSNR -- the eval side of the data-parallel path.

The reference validates on rank 0 alone (``dist_validation`` returns 0. on every other rank,
models/image_restoration_ref_model.py:319-323) and averages per-image ``calculate_psnr(tensor2img(result),
tensor2img(gt))`` (:352-395).  Images are independent units, so here every rank takes the images ``rank, rank + world,
...`` of the validation set, computes each PSNR with ``ops.psnr_u8`` (identical double to the reference's, 12 bytes per
image leave the device), and one all-reduce of ``[sum of PSNRs, count]`` in float64 gives every rank the dataset mean.
No activation or image crosses ranks.  With world size 1 the accumulation order is the reference's (sequential
``+=``), so the mean is identical too; for world > 1 the float64 sum is re-associated across ranks (last-bit effect).
"""
import torch
import torch.distributed as dist

from . import ops


def shard_indices(n_items, rank, world):
    """Validation items of this rank: the strided slice the reference's EnlargedSampler uses for training
    (data/data_sampler.py:30-43), without padding -- every image is scored exactly once."""
    return list(range(rank, n_items, world))


def reduce_mean(psnr_sum, count, process_group=None, device="cpu"):
    """All-reduce (sum) of [sum, count] in float64 -> (dataset mean, total count).  inf (identical images) propagates
    as in the reference's ``+=``."""
    t = torch.tensor([psnr_sum, float(count)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=process_group)
    total = int(t[1].item())
    return (float(t[0].item()) / total if total else 0.0), total


@torch.no_grad()
def validate_psnr(net_g, dataset, crop_border=0, process_group=None, psnr_fn=None, prepare=None):
    """Mean validation PSNR of ``net_g`` over ``dataset`` (indexable; items are dicts with 'lq', 'gt', 'ref' tensors
    [C, H, W] as the reference's datasets return them), sharded over the ranks of ``process_group``.

    psnr_fn(result[1,C,H,W], gt[1,C,H,W], crop_border) -> [float]; defaults to ``ops.psnr_u8`` (CUDA).  ``prepare`` maps
    an item to (lq, ref, gt) batched tensors on the right device; the default moves them to the net's device.
    Returns (mean PSNR over the whole set, number of images)."""
    rank = dist.get_rank(process_group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
    psnr_fn = psnr_fn or ops.psnr_u8
    try:
        dev = next(net_g.parameters()).device
    except (StopIteration, AttributeError):
        dev = torch.device("cpu")
    if prepare is None:
        def prepare(item):
            return (item["lq"].unsqueeze(0).to(dev), item["ref"].unsqueeze(0).to(dev), item["gt"].unsqueeze(0).to(dev))
    total, cnt = 0.0, 0
    for i in shard_indices(len(dataset), rank, world):
        lq, ref, gt = prepare(dataset[i])
        out = net_g(lq, ref)
        if isinstance(out, list):                      # a list of tensors is accepted, last = output (:252-256)
            out = out[-1]
        total += psnr_fn(out.float(), gt.float(), crop_border)[0]
        cnt += 1
    return reduce_mean(total, cnt, process_group, device=dev if dev.type == "cuda" else "cpu")
