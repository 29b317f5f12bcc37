#!/usr/bin/env python
"""Per-batch timeline of BankInference.run_host on the bench workload: when each H2D copy starts/ends
and when each batch's kernels finish, relative to the start of the pass (ms)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tmrnet_b200 as tb
from tmrnet_b200 import synth
from tmrnet_b200.infer import BankInference

batch = int(sys.argv[1]) if len(sys.argv) > 1 else None          # 0: default
host_batch = int(sys.argv[2]) if len(sys.argv) > 2 else None
tail = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
lengths = synth.video_lengths(40, seed=1234)
index = tb.LFBIndex.from_lengths(lengths, 10)
feats_host = torch.from_numpy(synth.features(sum(lengths), seed=1234)).pin_memory()
bank = torch.from_numpy(synth.bank(len(index), seed=1234)).to(dev)
model = tb.resnet_lstm(num_class=7, sequence_length=10)
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.head_state_dict(num_class=7, seed=1234).items()})
model = model.to(dev).eval()
eng = BankInference(model, index, 10, 30, batch_clips=batch, host_batch_clips=host_batch, tail_clips=tail)
with torch.no_grad():
    out, host_out = eng.run_host(feats_host, bank)
    for rep in range(3):
        eng.run_host(feats_host, bank, out=out, host_out=host_out)
        torch.cuda.synchronize()
    tl = []
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    eng.run_host(feats_host, bank, out=out, host_out=host_out, timeline=tl)
    t1.record()
    torch.cuda.synchronize()
print(f"batch_clips={eng.batch_clips} host_batch_clips={eng.host_batch_clips} tail={tail} total {t0.elapsed_time(t1):.3f} ms")
for n, c0, c1, k1 in tl:
    print(f"  clips {n:6d}  copy {t0.elapsed_time(c0):7.3f} -> {t0.elapsed_time(c1):7.3f}  ({c0.elapsed_time(c1):6.3f} ms)  kernels done {t0.elapsed_time(k1):7.3f}")
