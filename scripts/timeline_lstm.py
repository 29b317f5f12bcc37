import os, sys, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["TMR_TIMELINE"] = "1"
import tmrnet_b200 as tb
from tmrnet_b200 import ops, synth, _lib
dev = torch.device("cuda:0")
B, seq = 14336, 10
feats = torch.from_numpy(synth.features(B + seq - 1, seed=1)).to(dev)
sd = synth.head_state_dict(seed=1234)
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.to(dev).eval()
st = torch.arange(B, device=dev)
for _ in range(3): ops.lstm_last_frames(m.packs()[0], feats, st, seq, "tf32")
torch.cuda.synchronize()
lib = _lib.load()
n = 148 * 16 * 6
buf = (ctypes.c_longlong * n)()
lib._handle  # noqa
fn = ctypes.CDLL(_lib.LIB_PATH).tmr_debug_timeline
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn(buf, n) == 0
a = np.array(buf, dtype=np.int64).reshape(148, 16, 6)
# last launch = last recurrent step. columns: mma_wait_start, mma_start, mma_end, epi_wait_start, epi_start, epi_end
for cta in (0, 1, 73, 147):
    t0 = a[cta, 0, 0]
    print("CTA", cta)
    for it in range(7):
        r = a[cta, it] - t0
        print(f"  tile {it}: mma wait {r[1]-r[0]:6d}  mainloop {r[2]-r[1]:6d}  | epi waits acc {r[4]-r[3]:6d}  epilogue {r[5]-r[4]:6d}   [abs: mma_start {r[1]:7d} mma_end {r[2]:7d} epi_start {r[4]:7d} epi_end {r[5]:7d}]")
v = a[:, :6]
print("mean mainloop cycles", float((v[:, :, 2] - v[:, :, 1]).mean()), "mean epilogue cycles", float((v[:, :, 5] - v[:, :, 4]).mean()),
      "mean mma wait", float((v[:, :, 1] - v[:, :, 0]).mean()), "mean epi wait", float((v[:, :, 4] - v[:, :, 3]).mean()))
