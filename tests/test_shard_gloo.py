"""Clip sharding and the detection gather (the only collective), world_size 2 over gloo on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parq_b200.shard import DETECTION_KEYS, clip_range, gather_detections, shard_batch


def test_clip_range_partitions_everything():
    for n in (0, 1, 5, 16, 17, 1024):
        for world in (1, 2, 3, 8):
            spans = [clip_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    full = {k: torch.randn(n_clips, 4, n, generator=g) for k, n in zip(DETECTION_KEYS, (3, 3, 6, 10))}
    full["pred_mask"] = torch.rand(n_clips, 4, generator=g) > 0.5           # parse_pred's mask travels with the boxes (as bytes)
    local = shard_batch(full, rank, world)
    lo, hi = clip_range(n_clips, rank, world)
    assert local["ortho6d"].shape[0] == hi - lo
    got = gather_detections(local, n_clips)
    ok = all(torch.equal(got[k], full[k]) for k in DETECTION_KEYS + ("pred_mask",)) and got["pred_mask"].dtype == torch.bool
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_gather_detections_world2_gloo():
    world, n_clips = 2, 5            # uneven split: 3 + 2 clips
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret[r] for r in range(world))


def test_gather_without_process_group_is_identity():
    d = {k: torch.zeros(2, 4, 3) for k in DETECTION_KEYS}
    out = gather_detections(d, 2)
    assert all(out[k] is d[k] for k in DETECTION_KEYS)
