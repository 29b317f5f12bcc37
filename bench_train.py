#!/usr/bin/env python
"""Secondary bench (BASELINE.json configs[4]): head training step with the NCCL-over-NVLink gradient
all-reduce at N GPUs on a synthetic M2CAI-shaped job (C=8).  One process per GPU:

    python bench_train.py                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_train.py

Each rank trains on its own B clips per step (data parallel over clips); the ONLY collective is one
all-reduce (SUM) of the flat fp32 head gradient.  Also checks, on the first step, that the reduced
gradient equals the single-GPU gradient of the union batch (dropout off)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import tmrnet_b200 as tb  # noqa: E402
from tmrnet_b200 import synth  # noqa: E402
from tmrnet_b200.train import HeadTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=40, help="clips per rank per step (reference -t 400 = 40 clips)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    run(args.batch, args.steps, args.warmup)


def run(batch=40, steps=20, warmup=3):
    """Also reachable as `bench.py --train [--train-batch B] [--steps K] [--warmup W]` (same launch line as the
    inference bench, so the driver's torchrun command works unchanged)."""
    args = argparse.Namespace(batch=batch, steps=steps, warmup=warmup)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    C, seq, L, B = 8, 10, 30, args.batch
    sd = synth.head_state_dict(num_class=C, seed=1234)
    model = tb.resnet_lstm(num_class=C, sequence_length=seq)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.to(dev)
    # union batch of world*B clips; rank r owns rows [r*B, (r+1)*B)
    x_all = torch.from_numpy(synth.features(world * B * seq, seed=7).reshape(world * B, seq, 2048)).to(dev)
    lf_all = torch.from_numpy(synth.bank(world * B * L, seed=8).reshape(world * B, L, 512)).to(dev)
    y_all = torch.from_numpy(np.random.default_rng(9).integers(0, C, world * B)).to(dev)
    sl = slice(rank * B, (rank + 1) * B)
    x, lf, y = x_all[sl].contiguous(), lf_all[sl].contiguous(), y_all[sl].contiguous()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(mode):
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        tr = HeadTrainer(model, lr=5e-7, class_weight=np.ones(C, np.float32), seed=1, math_mode=mode)
        # correctness: all-reduced gradient == single-GPU gradient of the union batch
        tr.forward_backward(x, lf, y, dropout=False)
        tr.allreduce_grads()
        reduced = tr.grads.flat.clone()
        tr.forward_backward(x_all, lf_all, y_all, dropout=False)
        union = tr.grads.flat.clone()
        grad_err = float((reduced - union).abs().max() / union.abs().max())
        for _ in range(args.warmup):
            tr.step(x, lf, y)
        sync()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t_fb = t_ar = t_sgd = 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            e[0].record(); tr.forward_backward(x, lf, y)
            e[1].record(); tr.allreduce_grads()
            e[2].record(); tr.sgd_update()
            e[3].record()
            torch.cuda.synchronize()
            t_fb += e[0].elapsed_time(e[1]); t_ar += e[1].elapsed_time(e[2]); t_sgd += e[2].elapsed_time(e[3])
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, t_fb, t_ar, t_sgd, grad_err], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_fb, t_ar, t_sgd, grad_err = (float(v) for v in t)
        nbytes = tr.num_grad_elements * 4
        return {"value": world * B * args.steps / (ms / 1e3), "ms_per_step": ms / args.steps, "ms_fwd_bwd": t_fb / args.steps,
                "ms_allreduce": t_ar / args.steps, "ms_sgd": t_sgd / args.steps, "allreduce_bytes": nbytes,
                "allreduce_busbw_gbs": (2 * (world - 1) / world * nbytes / (t_ar / args.steps / 1e3) / 1e9) if world > 1 else None,
                "grad_elements": tr.num_grad_elements, "allreduced_vs_union_batch_grad_rel_err": grad_err}

    f16 = measure("f16")
    fp32 = measure("fp32")
    if rank == 0:
        line = {"metric": "TMRNet head training clips/sec (fwd+bwd+allreduce+SGD)", "unit": "clips/s", "n_gpus": world,
                "steps": args.steps}
        line.update(f16)
        line["fp32_math"] = fp32
        line["config"] = {"workload": "head training step, synthetic M2CAI-shaped (C=8), L=30, seq=10", "clips_per_gpu": B,
                          "math": "headline: fp16-rounded GEMM operands on tcgen05, fp32 accumulate (TMR_MATH_F16); fp32_math: every "
                                  "GEMM in fp32 on CUDA cores", "collective": "one NCCL all-reduce(SUM) of the flat head gradient"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
