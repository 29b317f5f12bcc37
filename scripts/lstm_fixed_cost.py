import os, sys, torch
sys.path.insert(0, "/root/repo")
import tmrnet_b200 as tb
from tmrnet_b200 import ops, synth
dev = torch.device("cuda:0")
sd = synth.head_state_dict(seed=1234)
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.to(dev).eval()
pk = m.packs()[0]
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for B in (256, 2368, 4736, 9472, 18944):
    seq = 10
    feats = torch.from_numpy(synth.features(B + seq - 1, seed=1)).to(dev)
    st = torch.arange(B, device=dev)
    t10 = timeit(lambda: ops.lstm_last_frames(pk, feats, st, 10, "f16"))
    t2 = timeit(lambda: ops.lstm_last_frames(pk, feats, st, 2, "f16"))
    print(f"B={B:6d}: per recurrent step {(t10 - t2) / 8:7.1f} us   (10-step {t10:8.1f} us, 2-step {t2:8.1f} us)")
