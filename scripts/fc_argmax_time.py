#!/usr/bin/env python
"""Time the classifier tail (512 -> C FC + softmax score + argmax) alone on a bench-sized batch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tmrnet_b200 as tb
from tmrnet_b200 import _lib, synth
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 83000
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.head_state_dict(seed=1234).items()}); m = m.to(dev).eval()
z = torch.rand(B, 512, device=dev)
wc, bc = m.fc_c.weight.detach().contiguous(), m.fc_c.bias.detach().contiguous()
lib = _lib.load()
# the kernels are reached through the classifier entry point only; time the whole entry (GEMM + tail) and the GEMM-free
# part by difference is not possible from here, so time torch's equivalent for scale and the entry itself
from tmrnet_b200 import ops
St = torch.rand(B, 512, device=dev); y1 = torch.rand(B, 512, device=dev)
pk = m.packs()[3]
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
print(f"B={B}: classifier entry (conversion + GEMM + fc/argmax) {timeit(lambda: ops.fc_argmax(pk, St, y1, 7, 'f16')):.1f} us")
