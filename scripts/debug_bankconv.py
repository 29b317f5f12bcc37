#!/usr/bin/env python
"""Compare the bank-level TimeConv variants against the per-clip kernel and the fp64 oracle."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tmr_oracle as orc
import tmrnet_b200 as tb
from tmrnet_b200 import ops, synth

dev = torch.device("cuda:0")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n_rows = 700
bank = synth.bank(n_rows, seed=3)
sd = synth.head_state_dict(seed=1234)
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.to(dev).eval()
pk = m.time_conv.packed()
bd = torch.from_numpy(bank).to(dev)
pb = ops.bankconv(pk, bd, 0, n_rows).cpu()                       # (rows, 7, 512)
# clips with slot-0 row r0 = 100 .. 600: window rows r0-k
r0s = np.arange(100, 600)
rows = r0s[:, None] - np.arange(L)[None, :]
win = torch.from_numpy(bank[rows])                               # (B, L, 512)
ref64 = orc.timeconv(win, sd, dtype=torch.float64)
gen = ops.timeconv_max(pk, win.to(dev), "f16").cpu()
gen32 = ops.timeconv_max(pk, win.to(dev), "fp32").cpu()
k = np.arange(L)
v = np.where(k <= 2, k + 1, np.where(L - 1 - k <= 2, 4 + (L - 1 - k), 0))
ded = pb[torch.from_numpy(rows), torch.from_numpy(np.broadcast_to(v, rows.shape).copy())]   # (B, L, 512)
scale = float(ref64.abs().max())
print("max|ref|", scale)
for name, t in (("general f16", gen), ("general fp32", gen32), ("dedup f16", ded)):
    d = (t.double() - ref64).abs()
    print(f"{name:14s} vs fp64: max {float(d.max())/scale:.2e}  per-slot max:", " ".join(f"{float(d[:, kk].max())/scale:.1e}" for kk in range(L)))
d = (ded.double() - gen.double()).abs()
print("dedup vs general f16: max %.2e" % (float(d.max()) / scale), " per-slot:", " ".join(f"{float(d[:, kk].max())/scale:.1e}" for kk in range(L)))
print("mean signed (dedup-general)/scale per slot:", " ".join(f"{float((ded.double()-gen.double())[:, kk].mean())/scale:+.1e}" for kk in range(L)))
