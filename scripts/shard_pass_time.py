#!/usr/bin/env python
"""Resident pass over ONE rank's shard of the bench job on one GPU (what each rank of `bench.py --gpus W` runs):
shard_pass_time.py [world] [--exp]   (--exp: the experiment build, which reads TMR_FORK_BANK=0/1)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tmrnet_b200 import _lib, build
if "--exp" in sys.argv:
    _lib.LIB_PATH = build.build(experiment=True)
import tmrnet_b200 as tb
from tmrnet_b200 import synth
from tmrnet_b200.infer import BankInference, VideoShard, shard_videos
world = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
dev = torch.device("cuda:0")
seq, L = 10, 30
lengths = synth.video_lengths(40, seed=1234)
v_lo, v_hi = shard_videos(lengths, world)[0]
sh = VideoShard(lengths, seq, L, v_lo, v_hi)
idx = sh.build_index()
feats = torch.from_numpy(synth.features(sh.frame_hi - sh.frame_lo, seed=1)).to(dev)
bank = torch.from_numpy(synth.bank(sh.row_hi - sh.row_lo, seed=2)).to(dev)
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.head_state_dict(seed=1234).items()}); m = m.to(dev).eval()
eng = BankInference(m, idx, seq, L, starts=sh.own_local_starts())
with torch.no_grad():
    out = eng.run(feats, bank)
    for _ in range(6): eng.run(feats, bank, out=out)          # the second pass over the same buffers captures the graph
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): eng.run(feats, bank, out=out)
    b.record(); torch.cuda.synchronize()
n = len(eng.starts_host)
ms = a.elapsed_time(b) / 50
print(f"world={world} rank 0: {n} clips, TMR_FORK_BANK={os.environ.get('TMR_FORK_BANK', 'default')}: {ms * 1e3:.1f} us per pass = {n / ms / 1e3:.2f} M frames/s")
