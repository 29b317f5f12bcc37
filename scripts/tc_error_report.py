#!/usr/bin/env python
"""Per-stage error of the tensor-core (fp16 operands) and fp32 CUDA paths against the fp64 oracle on a
config-4-shaped batch (run on the GPU box).  max|d|/max|ref| per stage."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tmr_oracle as orc  # noqa: E402
import tmrnet_b200 as tb  # noqa: E402
from tmrnet_b200 import ops, synth  # noqa: E402


def rel(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    return float((got - ref).abs().max() / ref.abs().max())


def main():
    dev = torch.device("cuda:0")
    B, seq, L, C = 512, 10, 30, 7
    lengths = [700, 650]
    starts_all = synth.clip_starts(lengths, seq)
    rng = np.random.default_rng(3)
    pick = np.sort(rng.choice(starts_all, size=B, replace=False))
    feats = synth.features(sum(lengths), seed=21)
    bank = synth.bank(len(starts_all), seed=21)
    x = np.stack([feats[s:s + seq] for s in pick])
    sd = synth.head_state_dict(num_class=C, seed=1234)
    lf = orc.get_long_feature(pick, orc.build_start_dict(starts_all.tolist()), bank, L)
    logits64, St64, Lt64, y164 = orc.head(x, lf, sd, dtype=torch.float64)
    logits32 = orc.head(x, lf, sd)[0]
    print(f"oracle fp32 vs fp64 logits: {rel(logits32, logits64):.2e}")
    m = tb.resnet_lstm(num_class=C)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to(dev).eval()
    packs = m.packs()
    xd, lfd = torch.from_numpy(x).to(dev), torch.from_numpy(lf).to(dev)
    for mode in ("fp32", "f16"):
        St = ops.lstm_last(packs[0], xd, mode)
        Lt = ops.timeconv_max(packs[1], lfd, mode)
        y1 = ops.nlblock(packs[2], St, Lt, mode)
        y1_iso = ops.nlblock(packs[2], St64.float().to(dev), Lt64.float().to(dev), mode)
        lg_iso = ops.fc_argmax(packs[3], St64.float().to(dev), y164.float().to(dev), C, mode)[0]
        lg = ops.head_fwd(*packs, xd, lfd, C, mode)[0]
        top2 = torch.topk(logits64, 2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1])
        flips = int((lg.argmax(1).cpu() != logits64.argmax(1)).sum())
        print(f"[{mode}] St {rel(St, St64):.2e} | Lt {rel(Lt, Lt64):.2e} | y1 {rel(y1, y164):.2e} (isolated {rel(y1_iso, y164):.2e})"
              f" | classifier isolated {rel(lg_iso, logits64):.2e} | head logits {rel(lg, logits64):.2e}"
              f" | argmax flips {flips}/{B} (min margin {float(margin.min()):.2e}, max|logit| {float(logits64.abs().max()):.2f})")


if __name__ == "__main__":
    main()
