#!/usr/bin/env python
"""Recurrent-step time of the LSTM on a bench-sized batch: (seq-step LSTM - 2-step LSTM) / (seq - 2), CUDA events.
Usage: lstm_step_time.py [B] [--lib path/to/variant.so] [--check]
(--lib: a variant library built for an A/B measurement; --check: h_T of 600 clips against the CPU oracle)"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
argv = [a for a in sys.argv[1:]]
lib_path = None
if "--lib" in argv:
    i = argv.index("--lib"); lib_path = argv[i + 1]; del argv[i:i + 2]
check = "--check" in argv
argv = [a for a in argv if a != "--check"]
sys.argv = [sys.argv[0]] + argv
from tmrnet_b200 import _lib
if lib_path:
    _lib.LIB_PATH = os.path.abspath(lib_path)
import tmrnet_b200 as tb
from tmrnet_b200 import ops, synth

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 41600
seq = 10
n_frames = B + seq - 1
feats = torch.from_numpy(synth.features(n_frames, seed=1)).to(dev)
sd = synth.head_state_dict(seed=1234)
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.to(dev).eval()
pk = m.packs()[0]
starts = torch.arange(B, device=dev)


def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t10 = timeit(lambda: ops.lstm_last_frames(pk, feats, starts, seq, "f16"))
t1 = timeit(lambda: ops.lstm_last_frames(pk, feats, starts, 2, "f16"))
step = (t10 - t1) / (seq - 2)
print(f"B={B} lib={os.path.basename(lib_path) if lib_path else 'product'}: "
      f"10-step {t10*1e3:.1f} us, 2-step {t1*1e3:.1f} us, recurrent step {step*1e3:.1f} us (x9 = {9*step*1e3:.0f} us) "
      f"= {B * 14336 / step / 1e6:.0f} GB/s (14 KB/clip), {B * 2.097152e6 / (step * 1e-3) / 1e12:.0f} TFLOP/s")

if check:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tmr_oracle as orc
    n = 600
    got = ops.lstm_last_frames(pk, feats[:n + seq - 1], starts[:n], seq, "f16").cpu()
    fh = feats[:n + seq - 1].cpu().numpy()
    ref = orc.lstm_last(torch.from_numpy(np.stack([fh[s:s + seq] for s in range(n)])), sd)
    err = float((got - ref).abs().max() / ref.abs().max())
    full = ops.lstm_last_frames(pk, feats, starts, seq, "f16")[:n].cpu()
    print(f"  h_T of {n} clips vs oracle: rel err {err:.2e}; same clips inside the {B}-clip launch: max |diff| {float((full - got).abs().max()):.1e}")
