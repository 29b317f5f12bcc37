#!/usr/bin/env python
"""Recurrent-step time of the LSTM on a bench-sized batch: (seq-step LSTM - 1-step LSTM) / (seq - 1), CUDA events.
Usage: lstm_step_time.py [B]   (env TMR_LSTM_WS=0: streamed GEMM engine; TMR_LSTM_WS_CFG=16: 16 epilogue warps)"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
t1 = timeit(lambda: ops.lstm_last_frames(pk, feats, starts, 1, "f16"))
step = (t10 - t1) / (seq - 1)
print(f"B={B} ws={os.environ.get('TMR_LSTM_WS', '1')} cfg={os.environ.get('TMR_LSTM_WS_CFG', '8')}: "
      f"10-step {t10*1e3:.1f} us, 1-step {t1*1e3:.1f} us, recurrent step {step*1e3:.1f} us "
      f"= {B * 14336 / step / 1e6:.0f} GB/s (14 KB/clip), {B * 2.097152e6 / (step * 1e-3) / 1e12:.0f} TFLOP/s")
