"""Video sharding for multi-GPU inference: host-side plan + halo logic, checked on CPU
with a world_size-2 gloo group.  Each rank derives the window ROWS of the clips it owns from its
local (halo-extended) index; gathered over ranks they must equal the global reference walk bit for
bit.  No data-path collective exists in the product; the all_gather here is only the test's check."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tmr_oracle as orc
from tmrnet_b200.infer import VideoShard, shard_videos

SEQ, L = 10, 30
LENGTHS = [57, 12, 140, 9, 33, 210, 45, 95, 11, 64, 18, 300]


def _local_rows(shard):
    idx = shard.build_index()
    own = shard.own_local_starts()
    f2r = idx.frame2row_host.astype(np.int64)
    rows = orc.window_rows_closed_form(own, f2r, L) + shard.row_lo        # back to global row ids
    return own + shard.frame_lo, rows


def test_shard_plan_is_contiguous_and_balanced():
    for world in (1, 2, 3, 4, 8):
        plan = shard_videos(LENGTHS, world)
        assert plan[0][0] == 0 and plan[-1][1] == len(LENGTHS)
        assert all(plan[i][1] == plan[i + 1][0] for i in range(world - 1))
    frames = [sum(LENGTHS[a:b]) for a, b in shard_videos(LENGTHS, 2)]
    assert max(frames) < 0.75 * sum(LENGTHS)


def test_halo_reproduces_global_windows_single_process():
    starts = orc.get_useful_start_idx(SEQ, LENGTHS)
    glob = orc.window_rows(starts, orc.build_start_dict(starts), L)
    for world in (2, 3, 5):
        got_s, got_r = [], []
        for lo, hi in shard_videos(LENGTHS, world):
            sh = VideoShard(LENGTHS, SEQ, L, lo, hi)
            s, r = _local_rows(sh)
            got_s.append(s)
            got_r.append(r)
        assert np.array_equal(np.concatenate(got_s), np.asarray(starts))
        assert np.array_equal(np.concatenate(got_r), glob)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_videos(LENGTHS, world)[rank]
    sh = VideoShard(LENGTHS, SEQ, L, lo, hi)
    s, r = _local_rows(sh)
    n = torch.tensor([len(s)])
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, n)
    mx = int(max(c.item() for c in counts))
    pad = torch.full((mx, L + 1), -1, dtype=torch.int64)
    pad[:len(s), 0] = torch.from_numpy(s)
    pad[:len(s), 1:] = torch.from_numpy(r)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    if rank == 0:
        cat = torch.cat([o[:int(c.item())] for o, c in zip(out, counts)])
        q.put(cat.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shards_equal_global():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    starts = orc.get_useful_start_idx(SEQ, LENGTHS)
    glob = orc.window_rows(starts, orc.build_start_dict(starts), L)
    assert np.array_equal(got[:, 0], np.asarray(starts))
    assert np.array_equal(got[:, 1:], glob)


# ---- training collective plumbing: flat gradient bucket + SUM all-reduce (gloo on CPU) ----
def _bucket_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tmrnet_b200.train import FlatBuffer
    shapes = [(4, 3), None, (7,), (2, 2, 5)]
    fbuf = FlatBuffer(shapes, "cpu")
    for i, v in enumerate(fbuf.views):
        if v is not None:
            v.copy_(torch.full(v.shape, float(rank + 1) * (i + 1)))
    dist.all_reduce(fbuf.flat, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put([None if v is None else v.clone().numpy() for v in fbuf.views])
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_bucket_allreduce_sum_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[1] is None
    for i in (0, 2, 3):
        assert np.all(got[i] == 3.0 * (i + 1))          # (1 + 2) * (i + 1)
    assert got[3].shape == (2, 2, 5)
