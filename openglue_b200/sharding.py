"""Multi-GPU plumbing for the matching core: image pairs are independent, so a batch is cut into
contiguous shards, one per rank (one process per GPU), with NO collective on the data path.  The
only exchange is the reduction of per-rank match statistics - what the reference's
``self.log(..., sync_dist=True)`` does (models/matching_module.py:102-103, 127-128)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def shard_range(total_pairs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split: rank r gets pairs [start, start+count); the first (total % world) ranks get one extra."""
    if not 0 <= rank < world_size:
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, extra = divmod(total_pairs, world_size)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def shard_pairs(data: Dict[str, object], rank: int, world_size: int) -> Dict[str, object]:
    """Slice every batched tensor of a ``data`` dict (superglue.py:30-38 keys) to this rank's pairs."""
    total = data['keypoints0'].shape[0]
    start, count = shard_range(total, rank, world_size)
    return {k: (v[start:start + count] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == total else v)
            for k, v in data.items()}


def match_statistics(matches0: torch.Tensor, matching_scores0: torch.Tensor) -> torch.Tensor:
    """[#pairs, #matches, sum of match confidences] of a shard (float64, on the tensors' device)."""
    ok = matches0 >= 0
    # no boolean-mask indexing, no host scalars moved to the device: nothing here synchronises with the host
    conf = torch.where(ok, matching_scores0.double(), matching_scores0.new_zeros((), dtype=torch.float64)).sum()
    pairs = torch.full((), float(matches0.shape[0]), device=matches0.device, dtype=torch.float64)
    return torch.stack([pairs, ok.sum().double(), conf])


def all_reduce_statistics(stats: torch.Tensor, group=None) -> Dict[str, float]:
    """Sum the per-rank statistics over the process group (NCCL on GPUs, gloo on CPU) and derive the means."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    pairs, nmatch, conf = (float(x) for x in stats.tolist())
    return {'pairs': pairs, 'matches_per_pair': nmatch / max(pairs, 1.0), 'mean_confidence': conf / max(nmatch, 1.0)}


def all_reduce_loss(loss: torch.Tensor, group=None) -> torch.Tensor:
    """Mean over ranks of a per-rank loss vector (e.g. stack([loss, metric_loss])): what the reference's
    ``self.log(..., sync_dist=True)`` does with the training losses (models/matching_module.py:102-103).  Each rank's
    criterion already divides by its own pair count (utils/losses.py:51), so for equal shards this is the loss of the
    whole batch.  In place on ``loss``' device (NCCL on GPUs, gloo on CPU); returns the reduced tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
        loss /= dist.get_world_size(group)
    return loss


def all_reduce_gradients(parameters, group=None, bucket_bytes: int = 32 << 20):
    """Data-parallel training step, the exchange: average ``.grad`` of the parameters over the ranks - what
    ``DistributedDataParallel`` does for the reference's ``MatchingTrainingModule`` (Lightning ``strategy='ddp'``, train.py).
    The gradients of this path appear all at once, at the end of the explicit backward schedule (``openglue_b200/training.py``),
    so there is nothing to overlap a per-layer hook with; they are packed into flat fp32 buckets (47.8 MB of parameters at the
    default config -> two 32 MB buckets), every bucket's all-reduce is issued asynchronously (NCCL: on the communicator's own stream,
    back to back, while the next bucket is being packed), then the means are scattered back in place.  Parameters without a
    gradient are skipped (all ranks run the same graph, so the set is the same everywhere).  Returns the number of buckets."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    grads = [p.grad for p in parameters if p.grad is not None]
    if world == 1 or not grads:
        return 0
    buckets, cur, cur_bytes = [], [], 0
    for g in grads:
        nbytes = g.numel() * 4
        if cur and cur_bytes + nbytes > bucket_bytes:
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(g)
        cur_bytes += nbytes
    buckets.append(cur)
    pending = []
    for b in buckets:
        flat = torch.cat([g.detach().reshape(-1).float() for g in b])
        pending.append((b, flat, dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)))
    for b, flat, work in pending:
        work.wait()
        flat /= world
        off = 0
        for g in b:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
    return len(buckets)
