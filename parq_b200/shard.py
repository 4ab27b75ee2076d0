"""Clip-level data parallelism: the only parallel axis of the PARQ decoder.

No state crosses clips inside the decoder (GroupNorm(1,C) and self-attention couple the
queries of ONE clip only), so clips are partitioned across ranks, weights are replicated and
the hot path contains no collective.  The single exchange is the gather of the
fixed-size per-clip detection tensors for evaluation: the reference's F1 calculator fuses
snippets per scene in dataset order (utils/f1_eval.py:293-352), so detections are returned in
global clip order.
"""
import torch
import torch.distributed as dist

DETECTION_KEYS = ("center_unnormalized", "size_unnormalized", "ortho6d", "sem_cls_prob")


def clip_range(n_clips, rank, world_size):
    """Contiguous block of clips owned by ``rank``: sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(n_clips, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(batch, rank, world_size):
    """Slice every tensor / wrapper of a clip-major batch dict to this rank's clips."""
    n = next(iter(batch.values())).shape[0]
    lo, hi = clip_range(n, rank, world_size)
    return {k: v[lo:hi] for k, v in batch.items()}


def gather_detections(last_iter, n_clips, group=None):
    """All-gather the last-iteration detections of every rank into global clip order.

    ``last_iter``: dict with DETECTION_KEYS, each (local_clips, Nq, n) -- plus, when present, the (local_clips, Nq)
    ``pred_mask`` of parse_pred (gathered as bytes).  Ranks may own different numbers of clips (block partition of
    ``n_clips``); tensors are padded to the largest block for the fixed-size collective and trimmed afterwards."""
    keys = DETECTION_KEYS + (("pred_mask",) if "pred_mask" in last_iter else ())
    if not dist.is_available() or not dist.is_initialized():
        return {k: last_iter[k] for k in keys}
    world = dist.get_world_size(group)
    sizes = [clip_range(n_clips, r, world)[1] - clip_range(n_clips, r, world)[0] for r in range(world)]
    cap = max(sizes)
    out = {}
    for k in keys:
        t = last_iter[k].contiguous()
        is_bool = t.dtype == torch.bool
        if is_bool:
            t = t.to(torch.uint8)
        pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        out[k] = torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)
        if is_bool:
            out[k] = out[k].bool()
    return out
