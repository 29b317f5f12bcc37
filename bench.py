#!/usr/bin/env python
"""bench.py — TMRNet head frames/s at L=30, seq=10 (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--math fp32|f16]

Workload (BASELINE.json configs[1]): train_non-local_mutiConv_resnet.py head — LSTM(2048->512, 10
frames) + multi-scale TimeConv + non-local block + FCs, L=30 — over a synthetic 40-video
Cholec80-shaped bank (~80 k frames, ~79.6 k clips; one clip = one predicted frame).  A "step" is
one pass of the head over every clip of the rank's bank.  N>1: one process per GPU, each rank holds
its own 40-video bank (video-sharded, no data-path collective; weak scaling).

value : frames/s with features + bank resident in HBM (inputs 818 MB > 126 MB L2, no flush needed).
e2e   : same pass through the public API with the per-frame features in pinned HOST memory copied
        H2D inside the timed region and preds/scores copied D2H (the bank is resident state, like
        the reference's g_LFB_* array).
roofline / cpu_baseline / clocks: see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEQ, L, C, NUM_VIDEOS = 10, 30, 7, 40
METRIC = "TMRNet head frames/sec at L=30, seq=10"
UNIT = "frames/s"
WORKLOAD = ("train_non-local_mutiConv_resnet.py head (LSTM 2048->512 x10 + TimeConv k3/5/7 + NLBlock + FC, "
            "C=7), L=30, seq=10, synthetic 40-video Cholec80-shaped bank")

# algorithmic work per clip, SURVEY.md 8(d)
FLOP_TIMECONV = 2 * 512 * 512 * 30 * 15          # 235.93 MFLOP
BYTES_GATHER = 2 * 30 * 512 * 4                  # 122 880 B
BYTES_RELATION = 30 * 512 * 4 + 2 * 512 * 4      # 65 536 B
FLOP_LSTM_STEP = 2 * 4 * 512 * 512               # 2.097 MFLOP per clip per recurrent step
FLOP_BANKCONV_ROW = 2 * 512 * 512 * 15           # 7.864 MFLOP per bank row (TimeConv deduplicated per row)
# per clip and recurrent step: projected row (fp32) 8 KB + c in/out (fp32) 2 x 2 KB + h in/out (fp16) 2 x 1 KB
BYTES_LSTM_STEP = 4 * 512 * 4 + 2 * 512 * 4 + 2 * 512 * 2
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel from the committed
# ncu --set full capture (profiles/r1_f16_ncu.md): 471.2 MB read + 106.4 MB written at 41600 clips
NCU_TRAFFIC_LSTM_STEP = (577.6e6, 41600)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops"], tensor_sustained=d.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 is None or (t0 <= t <= t1)]
        window = "timed region"
        if not rows:                       # timed region shorter than one sample: use the whole loaded run
            rows = [r for _, r in self.rows]
            window = "warm-up + timed + e2e (timed region shorter than the sampling period)"
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference head on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_head_rate(sample_clips: int, iters: int, warmup: int, seed=1234):
    """Oracle port (oracle/tmr_oracle.py: reference get_long_feature + NLBlock/TimeConv/LSTM/FC
    restated for torch-CPU) on `sample_clips` consecutive clips of the synthetic bank."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tmr_oracle as orc
    from tmrnet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lengths = synth.video_lengths(NUM_VIDEOS, seed=seed)[:2]
    starts_all = synth.clip_starts(lengths, SEQ)
    feats = synth.features(sum(lengths), seed=seed)
    bank = synth.bank(len(starts_all), seed=seed).astype(np.float64)       # reference bank dtype
    sd = synth.head_state_dict(num_class=C, seed=seed)
    d = orc.build_start_dict(starts_all.tolist())
    pick = starts_all[100:100 + sample_clips]
    x = np.stack([feats[s:s + SEQ] for s in pick])

    def step():
        lf = orc.get_long_feature(pick, d, bank, L)
        with torch.no_grad():
            logits = orc.head(x, lf, sd)[0]
            orc.eval_postproc(logits)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    dt = time.perf_counter() - t0
    return dict(value=sample_clips * iters / dt, ms_per_step=1e3 * dt / iters, cores=cores,
                sample=f"{iters} x {sample_clips} consecutive clips (gather + head + softmax/argmax), torch-CPU fp32, "
                       f"{cores} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 512
    r = cpu_head_rate(sample, iters=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "L": L, "seq": SEQ, "step": f"bounded sample of {sample} clips per step"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import tmrnet_b200 as tb
    from tmrnet_b200 import ops, synth
    from tmrnet_b200.infer import BankInference

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the head has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- rank-local synthetic bank (weak scaling: every rank owns 40 videos) ----
    seed = 1234 + 1000 * rank
    lengths = synth.video_lengths(NUM_VIDEOS, seed=seed)
    index = tb.LFBIndex.from_lengths(lengths, SEQ)
    n_frames, n_clips = sum(lengths), len(index)
    feats_host = torch.from_numpy(synth.features(n_frames, seed=seed)).pin_memory()
    bank_dev = torch.from_numpy(synth.bank(n_clips, seed=seed)).to(dev)
    feats_dev = feats_host.to(dev)
    model = tb.resnet_lstm(num_class=C, sequence_length=SEQ)
    sd = synth.head_state_dict(num_class=C, seed=1234)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.to(dev).eval()
    ops.set_math_mode(args.math)
    eng = BankInference(model, index, SEQ, L, batch_clips=args.batch or None)
    batch_clips = max(hi - lo for lo, hi, _, _ in eng.plan())
    out = None

    # ---- device-resident timing ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        for _ in range(args.warmup):
            out = eng.run(feats_dev, bank_dev, out=out)
        barrier()
        t_wall0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = eng.run(feats_dev, bank_dev, out=out)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t_wall1 = time.time()

        # ---- end to end: pinned host features -> H2D, preds/scores -> D2H, inside the timed region ----
        # BankInference.run_host double-buffers the per-batch H2D copies against the previous batch.
        host_out = (torch.empty(n_clips, dtype=torch.int64).pin_memory(),
                    torch.empty(n_clips, dtype=torch.float32).pin_memory())

        def e2e_step():
            eng.run_host(feats_host, bank_dev, out=out, host_out=host_out)

        for _ in range(max(1, args.warmup // 2)):
            e2e_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            e2e_step()
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
        clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

        # ---- per-kernel roofline probes on one batch of the same workload (rank 0) ----
        kern = {}
        if rank == 0:
            B = min(batch_clips, n_clips - 64)
            st = torch.from_numpy(eng.starts_host[:B]).to(dev)
            f2r, f2v = index.device_tables(dev)
            packs = model.packs()
            fr = feats_dev[: int(eng.starts_host[B - 1]) + SEQ]

            def timeit(fn, reps):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    fn()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / reps

            # LSTM: full recurrence minus the seq=1 pass (projection + step 0) = 9 recurrent-step launches
            t_l10 = timeit(lambda: ops.lstm_last_frames(packs[0], fr, st, SEQ), 10)
            t_l1 = timeit(lambda: ops.lstm_last_frames(packs[0], fr, st, 1), 10)
            t_step = (t_l10 - t_l1) / (SEQ - 1)
            t_bc = timeit(lambda: ops.bankconv(packs[1], bank_dev, 0, B + L), 10)
            # HBM-bound kernels on windows that do NOT dedupe in L2: random clip starts over the whole bank
            rnd = torch.from_numpy(np.random.default_rng(0).permutation(eng.starts_host)[:B].copy()).to(dev)
            t_g = timeit(lambda: ops.gather_windows(bank_dev, f2r, rnd, L), 20)
            win = ops.gather_windows(bank_dev, f2r, rnd, L)
            u = torch.from_numpy(synth.bank(B, seed=5)).to(dev)
            t_at = timeit(lambda: ops.attention(u, win), 20)
            t_tc = timeit(lambda: ops.timeconv_max(packs[1], win), 3)
            kern = {
                "lstm_step": {"ms": t_step, "tflops": FLOP_LSTM_STEP * B / t_step / 1e9, "clips": B},
                "bankconv": {"ms": t_bc, "tflops": FLOP_BANKCONV_ROW * (B + L) / t_bc / 1e9, "rows": B + L},
                "timeconv_per_clip": {"ms": t_tc, "tflops": FLOP_TIMECONV * B / t_tc / 1e9, "clips": B},
                "gather": {"ms": t_g, "gbs": BYTES_GATHER * B / t_g / 1e6, "clips": B},
                "attention": {"ms": t_at, "gbs": BYTES_RELATION * B / t_at / 1e6, "clips": B},
            }

    # ---- max over ranks ----
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        cnt = torch.tensor([n_clips], device=dev, dtype=torch.int64)
        dist.all_reduce(cnt)
        total_clips = int(cnt[0])
    else:
        total_clips = n_clips

    if rank == 0:
        pk = peaks()
        value = total_clips * args.steps / (ms / 1e3)
        e2e_v = total_clips * args.steps / (ms_e2e / 1e3)
        ls = kern["lstm_step"]
        def hbm(k, bpc):
            return {"achieved": kern[k]["gbs"], "peak": pk["hbm"], "unit": "GB/s", "frac": kern[k]["gbs"] / pk["hbm"],
                    "bytes_per_clip": bpc, "ms_per_launch": kern[k]["ms"], "clips_per_launch": kern[k]["clips"],
                    "inputs": "random clip starts over the whole bank (no L2 dedupe of window rows)"}
        def tens(k, unit_flop, units):
            return {"achieved": kern[k]["tflops"], "peak": pk["tensor"], "unit": "TFLOP/s",
                    "frac": kern[k]["tflops"] / pk["tensor"], "algorithmic_flop_per_unit": unit_flop,
                    "units_per_launch": units, "ms_per_launch": kern[k]["ms"]}
        # The recurrent step moves 14 KB per clip for 2.1 MFLOP: against the measured peaks the HBM time (2.2 ns
        # per clip) is the longer one (tensor: 1.25 ns at the measured bf16/fp16 rate), so the kernel is judged
        # on the HBM roofline; its tensor-core rate is reported beside it.
        step_gbs = BYTES_LSTM_STEP * ls["clips"] / ls["ms"] / 1e6
        roof = {"kernel": "umma_lstm_ws_kernel (weights-stationary recurrent step h.Whh^T + LSTM cell epilogue; 9 launches "
                          "per batch, largest share of the step, see profiles/r1_f16_ncu.md)",
                "bound": "hbm", "achieved": step_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": step_gbs / pk["hbm"],
                "traffic": NCU_TRAFFIC_LSTM_STEP[0] * ls["clips"] / NCU_TRAFFIC_LSTM_STEP[1],
                "traffic_source": f"ncu dram bytes of one launch at {NCU_TRAFFIC_LSTM_STEP[1]} clips, scaled to clips_per_launch",
                "peak_source": pk["source"],
                "algorithmic_bytes_per_clip": BYTES_LSTM_STEP, "clips_per_launch": ls["clips"], "ms_per_launch": ls["ms"],
                "how": "CUDA events: (10-step LSTM - 1-step LSTM) / 9 on one batch",
                "tensor": {"achieved": ls["tflops"], "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ls["tflops"] / pk["tensor"],
                           "algorithmic_flop_per_clip": FLOP_LSTM_STEP,
                           "note": "peak = measured bf16 burst (fp16 operands issue at the same rate)"},
                "tensor_kernels": {"bankconv": tens("bankconv", FLOP_BANKCONV_ROW, kern["bankconv"]["rows"]),
                                   "timeconv_per_clip": tens("timeconv_per_clip", FLOP_TIMECONV, kern["timeconv_per_clip"]["clips"])},
                "hbm_kernels": {"gather": hbm("gather", BYTES_GATHER), "attention": hbm("attention", BYTES_RELATION)}}
        cpu = None
        if world == 1 and not args.no_cpu:
            r = cpu_head_rate(256, iters=args.cpu_iters, warmup=1)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if args.math == "f16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "L": L, "seq": SEQ, "videos_per_gpu": NUM_VIDEOS,
                       "clips_per_gpu": n_clips, "frames_per_gpu": n_frames, "batch_clips": batch_clips, "host_batch_clips": eng.host_batch_clips,
                       "math": args.math, "l2": "inputs (818 MB/GPU) exceed the 126 MB L2; no flush",
                       "parallelism": f"video-sharded x{world}, no collective"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": n_frames * 2048 * 4,
                    "d2h_bytes_per_step": n_clips * 12, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": eng.launches_per_run() * args.steps,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default=os.environ.get("TMR_MATH", "f16"), choices=["fp32", "f16"])
    ap.add_argument("--batch", type=int, default=0,
                    help="clips per head launch sequence (0: the engine's default, an even split into batches of <= 65536)")
    ap.add_argument("--cpu-iters", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
