#!/usr/bin/env python
"""BASELINE.json configs[3]: memory-length sweep L = 10/30/60/120 x batch 32..256 clips (plus large
batches), gather / relation kernels against the measured HBM roofline with a COLD L2 (flushed before every timed
launch), and the whole head through the per-clip module API as a CUDA graph (GraphedHead, back-to-back replays).
Random clip starts over a 40-video bank (window rows do not dedupe in L2).  One JSON line per cell."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import tmrnet_b200 as tb  # noqa: E402
from tmrnet_b200 import ops, synth  # noqa: E402


_FLUSH = None


def timeit(fn, reps):
    """Median device time of ONE launch with a cold L2: a 256 MB write (> 126 MB L2) precedes every timed launch, so the
    windows a small batch touches cannot sit in L2 from the previous repetition (the round-1 sweep timed back-to-back
    launches over the same inputs and reported up to 1.66 x the HBM peak - L2 numbers, not HBM numbers)."""
    global _FLUSH
    if _FLUSH is None:
        _FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        _FLUSH.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def timeit_warm(fn, reps):
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


def main():
    dev = torch.device("cuda:0")
    peak = 6549.4
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    seq, C = 10, 7
    lengths = synth.video_lengths(40)
    idx = tb.LFBIndex.from_lengths(lengths, seq)
    starts_all = synth.clip_starts(lengths, seq)
    bank = torch.from_numpy(synth.bank(len(starts_all), seed=1234)).to(dev)
    feats = torch.from_numpy(synth.features(sum(lengths), seed=1234)).to(dev)
    sd = synth.head_state_dict(num_class=C, seed=1234)
    m = tb.resnet_lstm(num_class=C)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to(dev).eval()
    packs = m.packs()
    f2r, _ = idx.device_tables(dev)
    rng = np.random.default_rng(0)
    for L in (10, 30, 60, 120):
        for B in (32, 64, 128, 256, 4096, 16384):
            st_h = np.sort(rng.choice(starts_all, size=B, replace=False))
            st = torch.from_numpy(st_h).to(dev)
            reps = 30 if B <= 256 else 10
            t_g = timeit(lambda: ops.gather_windows(bank, f2r, st, L), reps)
            win = ops.gather_windows(bank, f2r, st, L)
            u = torch.from_numpy(synth.bank(B, seed=5)).to(dev)
            St = u
            t_a = timeit(lambda: ops.attention(u, win), reps)
            t_nl = timeit(lambda: ops.nlblock(packs[2], St, win), reps)
            x = torch.stack([feats[s:s + seq] for s in st_h[:min(B, 4096)]]) if B <= 4096 else None
            t_head = None
            if x is not None:
                from tmrnet_b200.graphs import GraphedHead
                gh = GraphedHead(m, x.shape[0], L)
                wx = win[:x.shape[0]]
                t_head = timeit_warm(lambda: gh.run(x, wx), 20)
                del gh
            gb_g = 2 * L * 512 * 4 * B / t_g / 1e6
            gb_a = (L * 512 * 4 + 2 * 512 * 4) * B / t_a / 1e6
            print(json.dumps({"L": L, "B": B, "gather_us": round(t_g * 1e3, 2), "gather_gbs": round(gb_g, 1),
                              "gather_frac_hbm": round(gb_g / peak, 3), "attention_us": round(t_a * 1e3, 2),
                              "attention_gbs": round(gb_a, 1), "attention_frac_hbm": round(gb_a / peak, 3),
                              "nlblock_us": round(t_nl * 1e3, 2),
                              "head_graph_us": None if t_head is None else round(t_head * 1e3, 1),
                              "head_clips_per_s": None if t_head is None else round(x.shape[0] / t_head * 1e3)}), flush=True)


if __name__ == "__main__":
    main()
