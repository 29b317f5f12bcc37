#!/usr/bin/env python
"""Latency of one head call at the reference's batch sizes: direct C-ABI call sequence vs CUDA-graph replay."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tmrnet_b200 import _lib, build
if os.environ.get("TMR_LSTM_PERSIST_MIN"): _lib.LIB_PATH = build.LIB_EXP      # experiment build reads the switch
import tmrnet_b200 as tb
from tmrnet_b200 import synth
from tmrnet_b200.graphs import GraphedHead
dev = torch.device("cuda:0")
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.head_state_dict(seed=1234).items()}); m = m.to(dev).eval()
def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for B in (4, 32, 120, 256, 1024):
    x = torch.from_numpy(synth.features(B * 10, seed=3).reshape(B, 10, 2048)).to(dev)
    lf = torch.from_numpy(synth.bank(B * 30, seed=4).reshape(B, 30, 512)).to(dev)
    gh = GraphedHead(m, B, 30)
    with torch.no_grad():
        t_direct = timeit(lambda: m.predict(x, lf))
    t_graph = timeit(lambda: gh.run(x, lf))
    print(json.dumps({"batch_clips": B, "direct_us": round(t_direct, 1), "graph_us": round(t_graph, 1),
                      "graph_frames_per_s": round(B / t_graph * 1e6)}), flush=True)
