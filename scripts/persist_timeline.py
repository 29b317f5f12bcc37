#!/usr/bin/env python
"""Per-item timeline of the persistent LSTM recurrence kernel (experiment build: python -m tmrnet_b200.build
--experiment).  Prints, for a few CTAs, where each role spends its time per item (ns, %globaltimer)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tmrnet_b200 import _lib, build  # noqa: E402
_lib.LIB_PATH = build.LIB_EXP
import tmrnet_b200 as tb  # noqa: E402
from tmrnet_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 82944
seq = 10
feats = torch.from_numpy(synth.features(B + seq - 1, seed=1)).to(dev)
sd = synth.head_state_dict(seed=1234)
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.to(dev).eval()
pk = m.packs()[0]
starts = torch.arange(B, device=dev)
for _ in range(3):
    ops.lstm_last_frames(pk, feats, starts, seq, "f16")
torch.cuda.synchronize()
ITEMS = 96
buf = np.zeros(148 * ITEMS * 8, dtype=np.uint64)
lib = _lib.load()
lib.tmr_debug_persist_timeline.argtypes = [C.c_void_p, C.c_int]
assert lib.tmr_debug_persist_timeline(buf.ctypes.data_as(C.c_void_p), buf.size) == 0
tl = buf.reshape(148, ITEMS, 8).astype(np.int64)
t0 = tl[:144, 0, 0][tl[:144, 0, 0] > 0].min()
for cta in (0, 1, 2, 16, 143):
    print(f"--- CTA {cta} (pair {cta // 2}, slice {(cta // 2) % 8}, group {cta // 16}) ---")
    print(" item | prod: wait->seen | mma: free->commit | epi: wait begin, got acc, done | epi busy, epi idle")
    for i in range(40):
        r = tl[cta, i] - t0
        print(f" {i:4d} | {r[0]:8d} {r[1] - r[0]:6d} | {r[2]:8d} {r[3] - r[2]:6d} | {r[4]:8d} {r[5] - r[4]:6d} {r[6] - r[5]:6d}")
ep = tl[:144:2, 2:90]                       # leader CTAs, steady state
busy = (ep[:, :, 6] - ep[:, :, 5]).mean()
idle = (ep[:, :, 5] - ep[:, :, 4]).mean()
flag = (tl[:144, 2:90, 1] - tl[:144, 2:90, 0]).mean()
mma = (ep[:, :, 3] - ep[:, :, 2]).mean()
period = (ep[:, 1:, 6] - ep[:, :-1, 6]).mean()
print(f"mean over CTAs/items: item period {period:.0f} ns | epilogue busy {busy:.0f}, waiting for accumulator {idle:.0f} | "
      f"producer flag wait {flag:.0f} | mma free->commit {mma:.0f}")
