#!/usr/bin/env python
"""Same-box bar: the head built from stock torch.nn modules (cuDNN LSTM / Conv1d, cuBLAS
linears) in PyTorch eager on the B200, on the bench workload's shapes.  Not the product and not the oracle:
a comparator.  Windows are gathered with one advanced-indexing op on the device (kinder than the reference's
Python get_long_feature, which runs on the host).  Prints one JSON line per batch size."""
import json, math, os, sys, time
import numpy as np
import torch, torch.nn as nn, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tmrnet_b200 import synth
import tmrnet_b200 as tb

SEQ, L, C = 10, 30, 7


class EagerHead(nn.Module):
    def __init__(self):
        super().__init__()
        self.lstm = nn.LSTM(2048, 512, batch_first=True)
        self.c3, self.c5, self.c7 = nn.Conv1d(512, 512, 3, padding=1), nn.Conv1d(512, 512, 5, padding=2), nn.Conv1d(512, 512, 7, padding=3)
        self.l1, self.l2, self.l3, self.l4 = (nn.Linear(512, 512) for _ in range(4))
        self.ln = nn.LayerNorm([1, 512])
        self.fc_h_c, self.fc_c = nn.Linear(1024, 512), nn.Linear(512, C)

    def forward(self, x, win):
        B = x.shape[0]
        y, _ = self.lstm(x)
        St = y[:, -1]                                              # (B,512)
        xt = win.transpose(1, 2)                                   # (B,512,L)
        pooled = F.max_pool1d(F.pad(xt, (1, 0)), 2, 1)
        Lt = torch.stack([xt, pooled, self.c3(xt), self.c5(xt), self.c7(xt)], 0).amax(0).transpose(1, 2)
        q = self.l1(St).view(B, 1, 512)
        att = torch.softmax(torch.matmul(q, self.l2(Lt).transpose(1, 2)) / math.sqrt(512), dim=2)
        r = torch.relu(self.ln(torch.matmul(att, self.l3(Lt))))
        y1 = St + self.l4(r).view(B, 512)
        z = torch.relu(self.fc_h_c(torch.cat([St, y1], 1)))
        logits = self.fc_c(z)
        p = torch.softmax(logits, 1)
        return p.max(1)


def main():
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = True          # give eager the tensor cores too
    torch.backends.cudnn.allow_tf32 = True
    lengths = synth.video_lengths(40, seed=1234)
    index = tb.LFBIndex.from_lengths(lengths, SEQ)
    feats = torch.from_numpy(synth.features(sum(lengths), seed=1234)).to(dev)
    bank = torch.from_numpy(synth.bank(len(index), seed=1234)).to(dev)
    f2r = torch.from_numpy(index.frame2row_host.astype(np.int64)).to(dev)
    starts_all = torch.from_numpy(np.fromiter(index.keys(), dtype=np.int64, count=len(index))).to(dev)
    m = EagerHead().to(dev).eval()
    ar, ak = torch.arange(SEQ, device=dev), torch.arange(1, L + 1, device=dev)
    for B in (256, 4096, 16384):
        s = starts_all[1000:1000 + B]

        def step():
            with torch.no_grad():
                x = feats[s[:, None] + ar[None]]                               # (B,10,2048)
                win = bank[f2r[(s[:, None] - ak[None]).clamp_min(0)]]          # (B,30,512)
                return m(x, win)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        a.record()
        for _ in range(iters):
            step()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        print(json.dumps({"impl": "torch eager (cuDNN/cuBLAS, TF32 allowed) on the same B200", "batch_clips": B,
                          "ms_per_batch": ms, "frames_per_s": B / ms * 1e3}), flush=True)


if __name__ == "__main__":
    main()
