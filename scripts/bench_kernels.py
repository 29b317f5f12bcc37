#!/usr/bin/env python
"""Stage-level timings (CUDA events) on one 8192-clip batch of the bench workload."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmrnet_b200 as tb
from tmrnet_b200 import ops, synth

dev = torch.device("cuda:0")
B, seq, L = 8192, 10, 30
n_frames = B + seq - 1
feats = torch.from_numpy(synth.features(n_frames, seed=1)).to(dev)
bank = torch.from_numpy(synth.bank(B + 64, seed=1)).to(dev)
sd = synth.head_state_dict(seed=1234)
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.to(dev).eval()
packs = m.packs()
starts = torch.arange(B, device=dev)
win = bank[:B * 0 + 64][None].expand(1, 64, 512)  # placeholder
win = torch.from_numpy(synth.bank(B * L, seed=2).reshape(B, L, 512)).to(dev)
St = torch.from_numpy(synth.bank(B, seed=5)).to(dev)

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

rows = {}
rows["lstm_last_frames (proj + 10 steps)"] = (timeit(lambda: ops.lstm_last_frames(packs[0], feats, starts, seq, "f16")), (n_frames * 2 * 2048 * 2048 + 9 * B * 2 * 512 * 2048) / 1e9)
rows["linear 8201x2048x2048 (projection)"] = (timeit(lambda: ops.linear(feats, m.lstm.weight_ih_l0, None, math_mode="f16")), n_frames * 2 * 2048 * 2048 / 1e9)
h = torch.from_numpy(synth.bank(B, seed=7)).to(dev)
rows["linear 8192x2048x512 (recurrent GEMM, plain epilogue)"] = (timeit(lambda: ops.linear(h, m.lstm.weight_hh_l0, None, math_mode="f16")), B * 2 * 512 * 2048 / 1e9)
rows["timeconv general"] = (timeit(lambda: ops.timeconv_max(packs[1], win, "f16"), 5), B * L * 2 * 512 * 512 * 15 / 1e9)
rows["bankconv (8222 rows)"] = (timeit(lambda: ops.bankconv(packs[1], bank, 0, B + 30)), (B + 30) * 2 * 512 * 512 * 15 / 1e9)
rows["nlblock"] = (timeit(lambda: ops.nlblock(packs[2], St, win, "f16")), B * 4 * 2 * 512 * 512 / 1e9)
rows["fc_argmax"] = (timeit(lambda: ops.fc_argmax(packs[3], St, St, 7, "f16")), B * 2 * 1024 * 512 / 1e9)
for k, (ms, gf) in rows.items():
    print(f"{k:55s} {ms*1e3:9.1f} us  {gf/ms:8.1f} TFLOP/s")
