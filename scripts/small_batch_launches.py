#!/usr/bin/env python
"""One direct head call at the reference's batch size (default 120 clips) - run under
`ncu --metrics gpu__time_duration.sum` to list the kernels of the small-batch path."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tmrnet_b200 as tb
from tmrnet_b200 import synth
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 120
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.head_state_dict(seed=1234).items()}); m = m.to(dev).eval()
x = torch.from_numpy(synth.features(B * 10, seed=3).reshape(B, 10, 2048)).to(dev)
lf = torch.from_numpy(synth.bank(B * 30, seed=4).reshape(B, 30, 512)).to(dev)
with torch.no_grad():
    for _ in range(3): m.predict(x, lf)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("call")
    out = m.predict(x, lf)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("ok", [tuple(o.shape) for o in out])
