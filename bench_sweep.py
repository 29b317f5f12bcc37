#!/usr/bin/env python
"""BASELINE.json configs[3]: memory-length sweep L = 10/30/60/120 x batch 32..256 clips (plus a large
batch), gather / relation kernels against the measured HBM roofline, and the whole head per-clip API.
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
            reps = 50 if B <= 256 else 10
            t_g = timeit(lambda: ops.gather_windows(bank, f2r, st, L), reps)
            win = ops.gather_windows(bank, f2r, st, L)
            u = torch.from_numpy(synth.bank(B, seed=5)).to(dev)
            St = u
            t_a = timeit(lambda: ops.attention(u, win), reps)
            t_nl = timeit(lambda: ops.nlblock(packs[2], St, win), reps)
            x = torch.stack([feats[s:s + seq] for s in st_h[:min(B, 4096)]]) if B <= 4096 else None
            t_head = None
            if x is not None:
                with torch.no_grad():
                    t_head = timeit(lambda: m.predict(x, win[:x.shape[0]]), max(3, reps // 5))
            gb_g = 2 * L * 512 * 4 * B / t_g / 1e6
            gb_a = (L * 512 * 4 + 2 * 512 * 4) * B / t_a / 1e6
            print(json.dumps({"L": L, "B": B, "gather_us": round(t_g * 1e3, 2), "gather_gbs": round(gb_g, 1),
                              "gather_frac_hbm": round(gb_g / peak, 3), "attention_us": round(t_a * 1e3, 2),
                              "attention_gbs": round(gb_a, 1), "attention_frac_hbm": round(gb_a / peak, 3),
                              "nlblock_us": round(t_nl * 1e3, 2),
                              "head_per_clip_api_us": None if t_head is None else round(t_head * 1e3, 1),
                              "head_clips_per_s": None if t_head is None else round(x.shape[0] / t_head * 1e3)}), flush=True)


if __name__ == "__main__":
    main()
