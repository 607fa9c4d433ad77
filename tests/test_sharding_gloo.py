"""Multi-rank host logic on the CPU: world_size-2 (and 3) `gloo` process groups shard a batch by
image (tlod_b200.sharding), run the per-rank work with the CPU oracle, exchange the results and
check them against the unsharded computation.  The path has no data-path collective; the
all_gather here is the test's own plumbing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from oracle.synth import synth_rois
from util import features


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tlod_b200 import sharding
        C, H, W, R = 6, 13, 17, 41
        feat = features(B, C, H, W, 5)
        rois = synth_rois(R, B, 6, im_h=H * 16, im_w=W * 16)  # unsorted image indices
        lo, hi = sharding.image_range(B, rank, world)
        my_feat = sharding.shard_batch(feat, rank, world)
        my_rois, index = sharding.shard_rois(rois, B, rank, world)
        assert my_feat.shape[0] == hi - lo
        assert int(my_rois[:, 0].min()) >= 0 and int(my_rois[:, 0].max()) < hi - lo
        out = torch.from_numpy(orc.roi_align_forward(my_feat.numpy(), my_rois.numpy(), 8, 8, 1 / 16))
        # exchange: sizes differ per rank, so pad to R rows
        pad = torch.zeros(R, C, 8, 8)
        pad[:out.shape[0]] = out
        idx = torch.full((R,), -1, dtype=torch.long)
        idx[:index.numel()] = index
        outs = [torch.zeros_like(pad) for _ in range(world)]
        idxs = [torch.zeros_like(idx) for _ in range(world)]
        dist.all_gather(outs, pad)
        dist.all_gather(idxs, idx)
        counts = [int((i >= 0).sum()) for i in idxs]
        assert sum(counts) == R  # a partition: every RoI on exactly one rank
        full = sharding.unshard_rows([o[:c] for o, c in zip(outs, counts)], [i[:c] for i, c in zip(idxs, counts)], R)
        ref = torch.from_numpy(orc.roi_align_forward(feat.numpy(), rois.numpy(), 8, 8, 1 / 16))
        ok = bool(torch.equal(full, ref))
        # weak-scaling bookkeeping the bench uses: whole-job units = sum over ranks, time = max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t.item()) == float(world)
        if rank == 0:
            ret.put(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 8), (2, 5), (3, 4)])
def test_image_sharding_over_gloo_ranks(world, B):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) is True


def test_image_range_is_a_balanced_partition():
    from tlod_b200 import sharding
    for n in (0, 1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [sharding.image_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.image_range(8, 2, 2)
