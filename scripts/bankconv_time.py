#!/usr/bin/env python
"""Time umma_bankconv_kernel alone on the bench bank (env TMR_BC_ABL selects timing ablations)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tmrnet_b200 import _lib, build
if os.environ.get("TMR_BC_ABL"): _lib.LIB_PATH = build.LIB_EXP
import tmrnet_b200 as tb
from tmrnet_b200 import ops, synth
dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 41600
bank = torch.from_numpy(synth.bank(rows + 64, seed=1)).to(dev)
m = tb.resnet_lstm(); m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.head_state_dict(seed=1234).items()}); m = m.to(dev).eval()
pk = m.packs()[1]
for _ in range(3): ops.bankconv(pk, bank, 0, rows)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): ops.bankconv(pk, bank, 0, rows)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"TMR_BC_ABL={os.environ.get('TMR_BC_ABL', '0')}: bankconv {rows} rows {ms * 1e3:.1f} us  ({rows * 7.864320e6 / ms / 1e9:.0f} TFLOP/s)")
