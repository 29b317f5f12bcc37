import os, sys, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["TMR_TIMELINE"] = "1"
import tmrnet_b200 as tb
from tmrnet_b200 import ops, synth, _lib
dev = torch.device("cuda:0")
B, seq = 18944, 10
feats = torch.from_numpy(synth.features(B + seq - 1, seed=1)).to(dev)
sd = synth.head_state_dict(seed=1234)
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.to(dev).eval()
st = torch.arange(B, device=dev)
for _ in range(3): ops.lstm_last_frames(m.packs()[0], feats, st, seq, "f16")
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
    for it in range(8):
        r = a[cta, it] - t0
        print(f"  tile {it}: mma wait {r[1]-r[0]:6d}  mainloop {r[2]-r[1]:6d}  | epi waits acc {r[4]-r[3]:6d}  epilogue {r[5]-r[4]:6d}   [abs: mma_start {r[1]:7d} mma_end {r[2]:7d} epi_start {r[4]:7d} epi_end {r[5]:7d}]")
v = a[:, :8]
print("mean mainloop cycles", float((v[:, :, 2] - v[:, :, 1]).mean()), "mean epilogue cycles", float((v[:, :, 5] - v[:, :, 4]).mean()),
      "mean mma wait", float((v[:, :, 1] - v[:, :, 0]).mean()), "mean epi wait", float((v[:, :, 4] - v[:, :, 3]).mean()))

# CTA 0, every epilogue warp: [wait start, accumulator ready, chunk 0 done, done] relative to the MMA thread's t0
fw = ctypes.CDLL(_lib.LIB_PATH).tmr_debug_timeline_warps
fw.argtypes = [ctypes.c_void_p, ctypes.c_int]
bw = (ctypes.c_longlong * (16 * 16 * 4))()
assert fw(bw, 16 * 16 * 4) == 0
w = np.array(bw, dtype=np.int64).reshape(16, 16, 4) - a[0, 0, 0]
for it in range(8):
    print(f"tile {it}: mma_start {a[0, it, 1] - a[0, 0, 0]} mma_end {a[0, it, 2] - a[0, 0, 0]}")
    for wi in range(16):
        r = w[it, wi]
        print(f"   warp {wi + 2:2d} (q{(wi + 2) & 3}, cols {((wi) >> 2) * 64:3d}): wait_start {r[0]:7d} ready {r[1]:7d} chunk0 {r[2] - r[1]:6d} chunk1 {r[3] - r[2]:6d} done {r[3]:7d}")

# kernel span seen by the device clock of each CTA (first MMA-thread stamp -> last epilogue stamp of warp 2)
n_t = int((a[0, :, 5] > 0).sum())
span = a[::2, n_t - 1, 5] - a[::2, 0, 0]
print(f"tiles per CTA {n_t}; span per leader CTA: mean {span.mean():.0f} max {span.max():.0f} cycles; per tile {span.mean() / n_t:.0f}")
lead = a[::2, :n_t]
print("leader CTAs: mainloop", float((lead[:, :, 2] - lead[:, :, 1]).mean()), "mma wait", float((lead[:, :, 1] - lead[:, :, 0]).mean()),
      "| gap epi_end(t) -> epi_wait_start(t+1)", float((lead[:, 1:, 3] - lead[:, :-1, 5]).mean()),
      "| tile period (mma_start deltas)", float(np.diff(lead[:, :, 1], axis=1).mean()))

fc = ctypes.CDLL(_lib.LIB_PATH).tmr_debug_timeline_cta
fc.argtypes = [ctypes.c_void_p, ctypes.c_int]
bc = (ctypes.c_longlong * (148 * 4))()
assert fc(bc, 148 * 4) == 0
k = np.array(bc, dtype=np.int64).reshape(148, 4)
print("per CTA (cycles): entry->setup done", float((k[:, 1] - k[:, 0]).mean()), "| setup->roles done", float((k[:, 2] - k[:, 1]).mean()),
      "| roles done->exit", float((k[:, 3] - k[:, 2]).mean()), "| entry->exit mean", float((k[:, 3] - k[:, 0]).mean()), "max", float((k[:, 3] - k[:, 0]).max()))
print("leader: entry -> first MMA-thread stamp", float((a[::2, 0, 0] - k[::2, 0]).mean()), "| last epilogue stamp (warp 2) -> roles done", float((k[::2, 2] - a[::2, n_t - 1, 5]).mean()))
