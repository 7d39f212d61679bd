"""Multi-GPU plumbing: one process per GPU (torch.distributed), where the reference uses single-process nn.DataParallel
(models/utils.py:93, flow_models/flow_model.py:109).

  * sampling / likelihood evaluation shard by image: every rank owns a contiguous slice of the global batch and runs its own
    CUDA graph; there is no data-path collective.  The Langevin corrector's batch-mean norms
    (sampling.py:286-288) can optionally be made global (`sampling.global_langevin_norms`): `allreduce_langevin_sums_` sums the
    three floats (sum |s_n|, sum |z_n|, N) over ranks, on the sampling stream, inside the sampler's CUDA graph.
  * training is data parallel: gradients accumulated by the engine into one flat buffer are summed with ONE all-reduce (NCCL over
    NVLink on GPUs, gloo on CPU in the tests) and divided by the world size before the global-norm clip (losses.FusedAdamW.step).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total, rank=None, world_size=None):
    """[start, stop) of this rank's contiguous slice of a global batch of `total` items; earlier ranks take the remainder"""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(total), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(x, rank=None, world_size=None):
    a, b = shard_range(x.shape[0], rank, world_size)
    return x[a:b]


def allreduce_mean_(flat, group=None):
    """in-place mean over ranks of a flat gradient buffer (the data-parallel exchange step)"""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(ws)
    return flat


def allreduce_langevin_sums_(sums3, group=None):
    """in place: `sums3` = [sum_n |s_n|, sum_n |z_n|, N] of this rank's shard (written by `indm_langevin_norm_sums`) -> the same
    three sums over all ranks; `indm_langevin_update_global` divides the first two by the third, which makes a sharded Langevin
    trajectory identical to the single-batch one.  Called by the sampler (indm_b200/sampling.py) on the sampling stream — NCCL
    collectives are capturable, so the exchange lives inside the per-step CUDA graph."""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(sums3, op=dist.ReduceOp.SUM, group=group)
    return sums3


def gather_cat(x, group=None):
    """concatenate per-rank tensors along dim 0 on every rank (logging / evaluation only)"""
    _, ws = world()
    if ws == 1:
        return x
    sizes = [torch.zeros(1, dtype=torch.int64, device=x.device) for _ in range(ws)]
    dist.all_gather(sizes, torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device), group=group)
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[:x.shape[0]] = x
    outs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:int(s)] for o, s in zip(outs, sizes)], dim=0)
