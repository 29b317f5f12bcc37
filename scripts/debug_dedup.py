import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tmr_oracle as orc
import tmrnet_b200 as tb
from tmrnet_b200 import ops, synth
from tmrnet_b200.infer import BankInference
dev = torch.device("cuda:0")
L, seq = 30, 10
lengths = [57, 12, 140, 9, 33, 210, 45, 400]
feats = synth.features(sum(lengths), seed=11)
starts = synth.clip_starts(lengths, seq)
bank = synth.bank(len(starts), seed=11)
sd = synth.head_state_dict(seed=1234)
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.to(dev).eval()
idx = tb.LFBIndex.from_lengths(lengths, seq)
f, b = torch.from_numpy(feats).to(dev), torch.from_numpy(bank).to(dev)
ref = BankInference(m, idx, seq, L, batch_clips=300, math_mode="tf32", dedup=False).run(f, b, want_st=True)
eng = BankInference(m, idx, seq, L, batch_clips=300, math_mode="tf32", dedup=True)
got = eng.run(f, b, want_st=True)
x = np.stack([feats[s:s + seq] for s in starts])
lf = orc.get_long_feature(starts, orc.build_start_dict(starts.tolist()), bank, L)
o64 = orc.head(x, lf, sd, dtype=torch.float64)[0]
scale = float(o64.abs().max())
print("St equal:", torch.equal(ref["St"], got["St"]))
d = (got["logits"] - ref["logits"]).abs().max(1).values.cpu().numpy() / scale
e_ref = (ref["logits"].cpu().double() - o64).abs().max(1).values.numpy() / scale
e_got = (got["logits"].cpu().double() - o64).abs().max(1).values.numpy() / scale
src = np.concatenate([p["src"] for p in eng.dedup_plan()])
print("max diff dedup-vs-general: regular %.2e irregular %.2e" % (d[src >= 0].max(), d[src < 0].max() if (src < 0).any() else 0))
print("vs fp64: general max %.2e mean %.2e | dedup max %.2e mean %.2e" % (e_ref.max(), e_ref.mean(), e_got.max(), e_got.mean()))
worst = np.argsort(-d)[:8]
print("worst clips:", [(int(i), int(src[i]), float("%.1e" % d[i])) for i in worst])
print("hist of d (regular):", np.histogram(d[src >= 0], bins=[0, 1e-6, 1e-5, 3e-5, 1e-4, 3e-4, 1e-3])[0])
