#!/usr/bin/env python
"""bench.py — TMRNet head frames/s at L=30, seq=10 (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--math fp32|f16]
    python bench.py --train [--train-batch B]        # BASELINE configs[4]: head training step (see bench_train.py)

Workload (BASELINE.json configs[1], and configs[2] for N > 1): train_non-local_mutiConv_resnet.py head —
LSTM(2048->512, 10 frames) + multi-scale TimeConv + non-local block + FCs, L=30 — over ONE synthetic
40-video Cholec80-shaped test set (~83 k frames / clips; one clip = one predicted frame).  A "step" is one
pass of the head over every clip of the set.

N > 1 (one process per GPU): the SAME 40-video set is sharded BY VIDEO (infer.shard_videos / VideoShard: each
rank holds its videos' features and bank rows plus the halo rows its first windows leak into);
no data-path collective; `value` = all clips / max-over-ranks device time -> "scaling": "strong".  After the
timed region every rank's predictions are gathered on rank 0 and compared BIT FOR BIT with the unsharded pass
rank 0 runs over the whole set (`shard_check`).  The weak-scaling number of round 1 (every rank its own
40-video set) is kept as the secondary field `weak`.

value : frames/s with features + bank resident in HBM (inputs 818 MB > 126 MB L2 at N=1; at N>1 the shard may
        fit L2, stated in config.l2).
e2e   : same pass through the public API with the per-frame features in pinned HOST memory copied H2D inside
        the timed region and preds/scores copied D2H (the bank is resident state, like the reference's
        g_LFB_* array).  `e2e_f16_features` is the optional input contract where the caller hands fp16
        features (half the PCIe bytes); the fp32 contract stays the headline.
roofline / cpu_baseline / gpu_eager_baseline / clocks: see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import hashlib
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
JOB_SEED = 1234
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
# measured DRAM traffic of the roofline kernel: profiles/lstm_traffic.json, written by scripts/ncu_traffic.py
# from the committed `ncu --set full` capture of the current kernel (never a literal in this file)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "lstm_traffic.json")


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
def cpu_head_rate(sample_clips: int, iters: int, warmup: int, seed=JOB_SEED):
    """Oracle port (oracle/tmr_oracle.py: the reference's get_long_feature walk — Python dict probes per
    (clip, slot) like TRAIN:298-326 — + NLBlock/TimeConv/LSTM/FC restated on torch-CPU modules' functional
    forms) over the SAME 40-video job as the GPU arm.  Each step is a bounded sample: `sample_clips`
    consecutive clips at a position that moves through the set from step to step."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tmr_oracle as orc
    from tmrnet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lengths = synth.video_lengths(NUM_VIDEOS, seed=seed)
    starts_all = synth.clip_starts(lengths, SEQ)
    feats = synth.features(sum(lengths), seed=seed)
    bank = synth.bank(len(starts_all), seed=seed).astype(np.float64)       # reference bank dtype
    sd = synth.head_state_dict(num_class=C, seed=seed)
    d = orc.build_start_dict(starts_all.tolist())
    n = len(starts_all)
    stride = max(1, (n - sample_clips) // max(1, iters + warmup))

    def step(i):
        lo = (i * stride) % max(1, n - sample_clips)
        pick = starts_all[lo:lo + sample_clips]
        x = np.stack([feats[s:s + SEQ] for s in pick])
        lf = orc.get_long_feature(pick, d, bank, L)
        with torch.no_grad():
            logits = orc.head(x, lf, sd)[0]
            orc.eval_postproc(logits)

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(iters):
        step(warmup + i)
    dt = time.perf_counter() - t0
    return dict(value=sample_clips * iters / dt, ms_per_step=1e3 * dt / iters, cores=cores,
                sample=f"{iters} x {sample_clips} consecutive clips of the 40-video job (clip assembly + window walk + "
                       f"head + softmax/argmax), torch-CPU fp32, {cores} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 512
    r = cpu_head_rate(sample, iters=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "L": L, "seq": SEQ, "videos": NUM_VIDEOS,
                   "step": f"bounded sample of {sample} clips per step, moving through the 40-video set"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                         "why_port": "the reference is Python scripts that cannot travel to the GPU box (no /root/reference "
                                     "there, sources may not be copied): the port restates them op for op on the same "
                                     "torch CPU kernels and is pinned to the reference's outputs by tests/golden"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# same-box GPU bar: the head from stock torch.nn modules in PyTorch eager (cuDNN / cuBLAS, TF32 allowed)
# ---------------------------------------------------------------------------------------------
def torch_eager_rate(dev, feats, bank, index, batch=16384, iters=5):
    import math
    import torch
    import torch.nn as nn
    import torch.nn.functional as Fn

    class EagerHead(nn.Module):
        def __init__(self):
            super().__init__()
            self.lstm = nn.LSTM(2048, 512, batch_first=True)
            self.c3, self.c5, self.c7 = (nn.Conv1d(512, 512, k, padding=k // 2) for k in (3, 5, 7))
            self.l1, self.l2, self.l3, self.l4 = (nn.Linear(512, 512) for _ in range(4))
            self.ln = nn.LayerNorm([1, 512])
            self.fc_h_c, self.fc_c = nn.Linear(1024, 512), nn.Linear(512, C)

        def forward(self, x, win):
            B = x.shape[0]
            y, _ = self.lstm(x)
            St = y[:, -1]
            xt = win.transpose(1, 2)
            pooled = Fn.max_pool1d(Fn.pad(xt, (1, 0)), 2, 1)
            Lt = torch.stack([xt, pooled, self.c3(xt), self.c5(xt), self.c7(xt)], 0).amax(0).transpose(1, 2)
            q = self.l1(St).view(B, 1, 512)
            att = torch.softmax(torch.matmul(q, self.l2(Lt).transpose(1, 2)) / math.sqrt(512), dim=2)
            r = torch.relu(self.ln(torch.matmul(att, self.l3(Lt))))
            y1 = St + self.l4(r).view(B, 512)
            z = torch.relu(self.fc_h_c(torch.cat([St, y1], 1)))
            return torch.softmax(self.fc_c(z), 1).max(1)

    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        f2r = torch.from_numpy(index.frame2row_host.astype(np.int64)).to(dev)
        starts_all = torch.from_numpy(np.fromiter(index.keys(), dtype=np.int64, count=len(index))).to(dev)
        m = EagerHead().to(dev).eval()
        ar, ak = torch.arange(SEQ, device=dev), torch.arange(1, L + 1, device=dev)
        B = min(batch, len(index) - 1000)
        s = starts_all[1000:1000 + B]

        def step():
            with torch.no_grad():
                x = feats[s[:, None] + ar[None]]
                win = bank[f2r[(s[:, None] - ak[None]).clamp_min(0)]]
                return m(x, win)

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        return {"value": B / ms * 1e3, "unit": UNIT, "batch_clips": B, "ms_per_batch": ms,
                "impl": "stock torch.nn modules in PyTorch eager on this GPU (cuDNN LSTM/Conv1d, cuBLAS, TF32 allowed), "
                        "windows gathered on the device; a comparator, not the product"}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
        del m
        torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import tmrnet_b200 as tb
    from tmrnet_b200 import ops, synth
    from tmrnet_b200.infer import BankInference, VideoShard, shard_videos

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

    model = tb.resnet_lstm(num_class=C, sequence_length=SEQ)
    sd = synth.head_state_dict(num_class=C, seed=JOB_SEED)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.to(dev).eval()
    ops.set_math_mode(args.math)

    class Job:
        """One rank's slice of a 40-video job: pinned host features, resident features / bank, the engine."""

        def __init__(self, seed, shard_rank=None, shard_world=1, f16_host=False):
            lengths = synth.video_lengths(NUM_VIDEOS, seed=seed)
            feats = synth.features(sum(lengths), seed=seed)
            n_rows_total = len(synth.clip_starts(lengths, SEQ))
            bank = synth.bank(n_rows_total, seed=seed)
            self.lengths = lengths
            if shard_world > 1:
                v_lo, v_hi = shard_videos(lengths, shard_world)[shard_rank]
                sh = VideoShard(lengths, SEQ, L, v_lo, v_hi)
                self.index = sh.build_index()
                starts = sh.own_local_starts()
                feats = feats[sh.frame_lo:sh.frame_hi]
                bank = bank[sh.row_lo:sh.row_hi]
                self.videos = (v_lo, v_hi)
                self.halo_frames = sh.own_frame_lo - sh.frame_lo
            else:
                self.index = tb.LFBIndex.from_lengths(lengths, SEQ)
                starts = None
                self.videos = (0, len(lengths))
                self.halo_frames = 0
            self.n_frames = feats.shape[0]
            self.feats_host = torch.from_numpy(np.ascontiguousarray(feats)).pin_memory()
            self.feats_host16 = self.feats_host.half().pin_memory() if f16_host else None
            self.bank_dev = torch.from_numpy(np.ascontiguousarray(bank)).to(dev)
            self.feats_dev = self.feats_host.to(dev)
            self.eng = BankInference(model, self.index, SEQ, L, batch_clips=args.batch or None, starts=starts)
            self.n_clips = len(self.eng.starts_host)
            self.out = None
            self.host_out = (torch.empty(self.n_clips, dtype=torch.int64).pin_memory(),
                             torch.empty(self.n_clips, dtype=torch.float32).pin_memory())

        def resident(self):
            self.out = self.eng.run(self.feats_dev, self.bank_dev, out=self.out)

        def e2e(self, f16=False):
            self.eng.run_host(self.feats_host16 if f16 else self.feats_host, self.bank_dev, out=self.out,
                              host_out=self.host_out)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1), t0, time.time()

    def reduce_max(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def reduce_sum(v):
        if world == 1:
            return int(v)
        t = torch.tensor([v], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        return int(t[0])

    # =========== headline: ONE 40-video job, sharded by video over the ranks ===========
    job = Job(JOB_SEED, rank, world, f16_host=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        ms, t_wall0, t_wall1 = timed(job.resident, args.steps, args.warmup)
        ms_e2e, _, _ = timed(job.e2e, args.steps, max(1, args.warmup // 2))
        ms_e2e16, _, _ = timed(lambda: job.e2e(True), args.steps, max(1, args.warmup // 2))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms, ms_e2e, ms_e2e16 = reduce_max([ms, ms_e2e, ms_e2e16])
    total_clips = reduce_sum(job.n_clips)
    total_frames_h2d = reduce_sum(job.n_frames)
    batch_clips = max(hi - lo for lo, hi, _, _ in job.eng.plan())
    launches = job.eng.launches_per_run()
    inputs_mb = (job.n_frames * 2048 * 4 + job.bank_dev.numel() * 4) / 1e6

    # =========== N > 1: bit-identity of the sharded job with the unsharded pass + the weak-scaling leg ===========
    shard_check = weak = None
    full = job
    if world > 1:
        mine = {"rank": rank, "videos": job.videos, "clips": job.n_clips, "halo_frames": job.halo_frames,
                "pred": job.out["pred"].cpu().numpy(), "logits": job.out["logits"].cpu().numpy()}
        parts = [None] * world if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
        # weak leg: every rank its own 40-video job (rank 0's is the headline job, unsharded)
        wjob = Job(JOB_SEED + 1000 * rank)
        with torch.no_grad():
            wms, _, _ = timed(wjob.resident, args.steps, args.warmup)
            wms_e2e, _, _ = timed(wjob.e2e, args.steps, max(1, args.warmup // 2))
        wms, wms_e2e = reduce_max([wms, wms_e2e])
        wclips = reduce_sum(wjob.n_clips)
        weak = {"scaling": "weak", "value": wclips * args.steps / (wms / 1e3), "unit": UNIT, "ms_per_step": wms / args.steps,
                "e2e_value": wclips * args.steps / (wms_e2e / 1e3), "e2e_ms_per_step": wms_e2e / args.steps,
                "videos_per_gpu": NUM_VIDEOS, "clips_total": wclips,
                "note": "round-1 configuration: every rank owns an independent 40-video set"}
        if rank == 0:
            parts.sort(key=lambda p: p["rank"])
            pred_sh = np.concatenate([p["pred"] for p in parts])
            logit_sh = np.concatenate([p["logits"] for p in parts])
            pred_full = wjob.out["pred"].cpu().numpy()
            logit_full = wjob.out["logits"].cpu().numpy()
            shard_check = {
                "equals_unsharded": bool(pred_sh.shape == pred_full.shape and np.array_equal(pred_sh, pred_full)),
                "logits_bit_identical": bool(logit_sh.shape == logit_full.shape and np.array_equal(logit_sh, logit_full)),
                "pred_sha256_sharded": hashlib.sha256(pred_sh.tobytes()).hexdigest()[:16],
                "pred_sha256_unsharded": hashlib.sha256(pred_full.tobytes()).hexdigest()[:16],
                "clips": int(len(pred_sh)),
                "shards": [{"rank": p["rank"], "videos": list(p["videos"]), "clips": p["clips"], "halo_frames": p["halo_frames"]}
                           for p in parts],
                "how": "every rank's pred / logits gathered on rank 0 after the timed region, compared with the pass "
                       "rank 0 runs over the whole 40-video set in the same process",
            }
        full = wjob if rank == 0 else None

    # =========== per-kernel roofline probes on one batch of the full job (rank 0) ===========
    kern, eager, cpu, small = {}, None, None, None
    if rank == 0:
        eng, index = full.eng, full.index
        feats_dev, bank_dev = full.feats_dev, full.bank_dev
        with torch.no_grad():
            B = min(max(hi - lo for lo, hi, _, _ in eng.plan()), full.n_clips - 64)
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

            # the recurrence alone: the same call with seq = 10 and seq = 2 runs the same projection (step 0 fused into
            # its epilogue) and the same persistent recurrence kernel over 9 / 1 steps; fixed costs cancel in the
            # difference (8 steps), scaled to the 9 steps of one launch
            t_l10 = timeit(lambda: ops.lstm_last_frames(packs[0], fr, st, SEQ), 10)
            t_l2 = timeit(lambda: ops.lstm_last_frames(packs[0], fr, st, 2), 10)
            t_step = (t_l10 - t_l2) / (SEQ - 2)
            t_rec = t_step * (SEQ - 1)
            t_bc = timeit(lambda: ops.bankconv(packs[1], bank_dev, 0, B + L), 10)
            # HBM-bound kernels on windows that do NOT dedupe in L2: random clip starts over the whole bank
            rnd = torch.from_numpy(np.random.default_rng(0).permutation(eng.starts_host)[:B].copy()).to(dev)
            t_g = timeit(lambda: ops.gather_windows(bank_dev, f2r, rnd, L), 20)
            win = ops.gather_windows(bank_dev, f2r, rnd, L)
            u = torch.from_numpy(synth.bank(B, seed=5)).to(dev)
            t_at = timeit(lambda: ops.attention(u, win), 20)
            t_tc = timeit(lambda: ops.timeconv_max(packs[1], win), 3)
            del win
            kern = {
                "lstm_step": {"ms": t_step, "tflops": FLOP_LSTM_STEP * B / t_step / 1e9, "clips": B,
                              "ms_recurrence": t_rec, "ms_lstm_total": t_l10},
                "bankconv": {"ms": t_bc, "tflops": FLOP_BANKCONV_ROW * (B + L) / t_bc / 1e9, "rows": B + L},
                "timeconv_per_clip": {"ms": t_tc, "tflops": FLOP_TIMECONV * B / t_tc / 1e9, "clips": B},
                "gather": {"ms": t_g, "gbs": BYTES_GATHER * B / t_g / 1e6, "clips": B},
                "attention": {"ms": t_at, "gbs": BYTES_RELATION * B / t_at / 1e6, "clips": B},
            }
            if not args.no_eager and world == 1:
                eager = torch_eager_rate(dev, feats_dev, bank_dev, index)
            # BASELINE configs[3] / the reference's own call pattern: ONE head call on 120 clips (1200 frames) with an explicit
            # window, as the scripts issue it (EVAL:470-495) - 7 launches: row table, feature conversion, input projection,
            # small-batch recurrence, window conversion + per-clip TimeConv (parallel branch), fused relation + classifier
            try:
                from tmrnet_b200.graphs import GraphedHead
                Bs = 120
                xs = feats_dev[: Bs * SEQ].reshape(Bs, SEQ, 2048).contiguous()
                lfs = bank_dev[: Bs * L].reshape(Bs, L, 512).contiguous()
                gh = GraphedHead(model, Bs, L)
                t_direct = timeit(lambda: model.predict(xs, lfs), 50)
                t_graph = timeit(lambda: gh.run(xs, lfs), 50)
                small = {"batch_clips": Bs, "L": L, "us_per_call_cuda_graph": t_graph * 1e3, "us_per_call_direct": t_direct * 1e3,
                         "frames_per_s": Bs / (t_graph * 1e-3), "kernels_per_call": 7,
                         "note": "one resnet_lstm.forward-sized call (120 clips x 10 frames, window given) incl. the copy of its inputs "
                                 "into the graph's static buffers; round 1: 274 us"}
                del gh
            except Exception as e:          # a probe beside the headline: never let it take the bench line down
                small = {"batch_clips": 120, "error": f"{type(e).__name__}: {e}"[:300]}
        if not args.no_cpu and world == 1:             # the CPU arm beside it: rank 0 at N = 1 only
            r = cpu_head_rate(256, iters=args.cpu_iters, warmup=1)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        pk = peaks()
        value = total_clips * args.steps / (ms / 1e3)
        e2e_v = total_clips * args.steps / (ms_e2e / 1e3)
        e2e16_v = total_clips * args.steps / (ms_e2e16 / 1e3)
        ls = kern["lstm_step"]

        def hbm(k, bpc):
            return {"achieved": kern[k]["gbs"], "peak": pk["hbm"], "unit": "GB/s", "frac": kern[k]["gbs"] / pk["hbm"],
                    "bytes_per_clip": bpc, "ms_per_launch": kern[k]["ms"], "clips_per_launch": kern[k]["clips"],
                    "inputs": "random clip starts over the whole bank (no L2 dedupe of window rows)"}

        def tens(k, unit_flop, units):
            return {"achieved": kern[k]["tflops"], "peak": pk["tensor"], "unit": "TFLOP/s",
                    "frac": kern[k]["tflops"] / pk["tensor"], "algorithmic_flop_per_unit": unit_flop,
                    "units_per_launch": units, "ms_per_launch": kern[k]["ms"]}

        # Dominant kernel: the persistent LSTM recurrence (all 9 recurrent steps of every clip in ONE launch).  SURVEY.md
        # 8(d) puts the LSTM on the tensor roofline (AI ~1 300 flop/B): algorithmic work = 9 steps x 2.097 MFLOP per clip
        # (step 0 has h = 0 and rides in the projection); its measured DRAM traffic (ncu, committed) is reported beside it.
        traffic = json.load(open(TRAFFIC_FILE)) if os.path.exists(TRAFFIC_FILE) else None
        rec_flop = FLOP_LSTM_STEP * (SEQ - 1)
        rec_tflops = rec_flop * ls["clips"] / ls["ms_recurrence"] / 1e9
        dram_per_clip = (traffic["dram_bytes_per_clip_step"] * traffic["steps_per_launch"]) if traffic else None
        roof = {"kernel": (traffic or {}).get("kernel", "umma_lstm_persist_kernel (persistent LSTM recurrence)"),
                "bound": "tensor", "achieved": rec_tflops, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": rec_tflops / pk["tensor"],
                "traffic": (dram_per_clip * ls["clips"]) if traffic else None,
                "traffic_source": (traffic or {}).get("source"),
                "peak_source": pk["source"] + " (bf16 burst; fp16 operands issue at the same rate)",
                "algorithmic_flop_per_clip": rec_flop, "clips_per_launch": ls["clips"], "ms_per_launch": ls["ms_recurrence"],
                "ms_lstm_total": ls["ms_lstm_total"], "share_of_step": ls["ms_recurrence"] * (total_clips / world) / ls["clips"] / (ms / args.steps),
                "how": "CUDA events on one batch of the full job: (10-step LSTM - 2-step LSTM) x 9/8 - both legs run the same "
                       "projection with step 0 fused and the same persistent kernel over 9 / 1 steps, so fixed costs cancel",
                "previous_round": {"kernel": "umma_lstm_ws_kernel x 9 launches (round 1)", "frac_of_the_same_tensor_peak": 0.43,
                                   "ms_for_the_same_work": 2.36,
                                   "note": "round 1 reported that kernel against HBM (0.73 of peak on 14 336 B per clip and step, "
                                           "125 KB per clip over the recurrence); the persistent kernel no longer moves those bytes, so "
                                           "it is reported on the roofline SURVEY.md 8(d) assigns the LSTM: tensor"},
                "hbm": {"dram_bytes_per_clip": dram_per_clip,
                        "note": "measured DRAM traffic of the whole recurrence per clip (c never leaves the SM, h is exchanged "
                                "through L2, projected rows of consecutive steps hit L2); the per-step kernels of round 1 moved "
                                "~125 KB per clip"},
                "tensor_kernels": {"bankconv": tens("bankconv", FLOP_BANKCONV_ROW, kern["bankconv"]["rows"]),
                                   "timeconv_per_clip": tens("timeconv_per_clip", FLOP_TIMECONV, kern["timeconv_per_clip"]["clips"])},
                "hbm_kernels": {"gather": hbm("gather", BYTES_GATHER), "attention": hbm("attention", BYTES_RELATION)}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16" if args.math == "f16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "L": L, "seq": SEQ, "videos": NUM_VIDEOS, "clips_total": total_clips,
                       "clips_rank0": job.n_clips, "frames_rank0": job.n_frames, "batch_clips": batch_clips,
                       "host_batch_clips": job.eng.host_batch_clips, "math": args.math,
                       "l2": f"rank 0 inputs {inputs_mb:.0f} MB vs 126 MB L2; no flush (every pass streams features, projected rows "
                             f"and per-step state far beyond L2)",
                       "parallelism": f"ONE 40-video set sharded by video x{world} (halo rows per shard), no collective"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": total_frames_h2d * 2048 * 4,
                    "d2h_bytes_per_step": total_clips * 12, "ms_per_step": ms_e2e / args.steps},
            "e2e_f16_features": {"value": e2e16_v, "unit": UNIT, "h2d_bytes_per_step": total_frames_h2d * 2048 * 2,
                                 "d2h_bytes_per_step": total_clips * 12, "ms_per_step": ms_e2e16 / args.steps,
                                 "note": "optional input contract: the caller hands fp16 features (the tensor-core path rounds "
                                         "them to fp16 anyway, so results are bit-identical); fp32 stays the headline"},
            "gpu_launches": launches * args.steps,
            "roofline": roof, "cpu_baseline": cpu, "gpu_eager_baseline": eager, "clocks": clocks,
            "shard_check": shard_check, "weak": weak, "reference_batch_call": small,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if shard_check is not None and not shard_check["equals_unsharded"]:
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default=os.environ.get("TMR_MATH", "f16"), choices=["fp32", "f16"])
    ap.add_argument("--batch", type=int, default=0,
                    help="clips per head launch sequence (0: the engine's default, an even split into batches of <= 131072)")
    ap.add_argument("--cpu-iters", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--train", action="store_true", help="BASELINE configs[4]: head training step (bench_train.py)")
    ap.add_argument("--train-batch", type=int, default=40)
    args = ap.parse_args()
    if args.train:
        import bench_train
        bench_train.run(batch=args.train_batch, steps=args.steps, warmup=max(3, args.warmup))
        return
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
