#!/usr/bin/env python
"""H2D-only probe for the end-to-end leg of bench.py: every rank copies the pinned fp32 features of ITS shard of the
40-video job to its GPU, in the same 9 472-clip pieces run_host uses, with NO kernels - the copy engine and the
host-memory / PCIe-root path alone.  Run it like the bench:

    python scripts/h2d_probe.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/h2d_probe.py

Prints one JSON line on rank 0: per-rank and aggregate GB/s (max-over-ranks device time), for the sharded job
(strong scaling: total bytes fixed) and for the round-1 configuration (every rank copies a whole 40-video set)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tmrnet_b200 import synth  # noqa: E402
from tmrnet_b200.infer import VideoShard, shard_videos  # noqa: E402

SEQ, L, V, SEED = 10, 30, 40, 1234
PIECE_FRAMES = 9472 + SEQ - 1


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lengths = synth.video_lengths(V, seed=SEED)
    n_all = int(sum(lengths))
    if world > 1:
        v_lo, v_hi = shard_videos(lengths, world)[rank]
        sh = VideoShard(lengths, SEQ, L, v_lo, v_hi)
        n_mine = sh.frame_hi - sh.frame_lo
    else:
        n_mine = n_all
    host = torch.empty((n_all, 2048), dtype=torch.float32).pin_memory()
    host.normal_()
    stage = [torch.empty((PIECE_FRAMES, 2048), dtype=torch.float32, device=dev) for _ in range(3)]
    stream = torch.cuda.Stream(device=dev)

    def copy(n_frames):
        with torch.cuda.stream(stream):
            i = 0
            for lo in range(0, n_frames, PIECE_FRAMES):
                hi = min(n_frames, lo + PIECE_FRAMES)
                stage[i % 3][:hi - lo].copy_(host[lo:hi], non_blocking=True)
                i += 1

    def timed(n_frames, reps=10):
        for _ in range(2):
            copy(n_frames)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            copy(n_frames)
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        t = torch.tensor([ms, n_frames * 8192.0], device=dev, dtype=torch.float64)
        if world > 1:
            mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = t.clone(); dist.all_reduce(sm)
            return float(mx[0]), float(sm[1]), float(t[0])
        return ms, float(t[1]), ms

    ms_s, bytes_s, mine_s = timed(n_mine)
    ms_w, bytes_w, mine_w = timed(n_all)
    if rank == 0:
        print(json.dumps({
            "probe": "h2d_only", "n_gpus": world, "piece_bytes": PIECE_FRAMES * 8192,
            "sharded_job": {"bytes_total": bytes_s, "ms_max_over_ranks": ms_s, "aggregate_gbs": bytes_s / ms_s / 1e6,
                            "rank0_gbs": n_mine * 8192 / mine_s / 1e6},
            "one_job_per_rank": {"bytes_total": bytes_w, "ms_max_over_ranks": ms_w, "aggregate_gbs": bytes_w / ms_w / 1e6,
                                 "per_rank_gbs": bytes_w / world / ms_w / 1e6},
            "host_cpus": os.cpu_count()}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
