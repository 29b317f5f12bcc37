"""CUDA-graph replay of the per-clip head for the reference's own batch sizes.

The reference calls `model.forward(inputs, long_feature)` on batches of 1200 frames / seq = 120 clips
(train_non-local_mutiConv_resnet.py:836-880, eval ...resnest.py:470-495).  At that size the ~27 kernels of
`tmr_head_fwd` are launch-bound, so `GraphedHead` captures them ONCE into a CUDA graph over static buffers
and replays it per batch: inputs are copied into the static buffers on the caller's stream, outputs are
views of static tensors (valid until the next `run`).  The graph runs the TimeConv of the window and the LSTM chain as
parallel branches (two capture streams) that join in the fused relation + classifier launch; same arithmetic, same
results bit for bit as the direct call.
"""
from __future__ import annotations

import torch

from . import ops

F, D = 2048, 512


class GraphedHead:
    """model: tmrnet_b200.resnet_lstm (eval).  One graph per (B, L); re-captured when the packed weights are
    rebuilt (parameter update / load_state_dict) or the math mode changes."""

    def __init__(self, model, batch_clips: int, L: int = 30, device=None):
        self.model = model
        self.B, self.L, self.seq = int(batch_clips), int(L), int(model.sequence_length)
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedHead needs a CUDA device (there is no CPU fallback)")
        self.device = dev
        self.x = torch.zeros((self.B, self.seq, F), dtype=torch.float32, device=dev)
        self.long_feature = torch.zeros((self.B, self.L, D), dtype=torch.float32, device=dev)
        self._graph = None
        self._key = None
        self._out = None

    def _capture_key(self):
        packs = self.model.packs()
        mode = ops._mode(self.model.math_mode)
        return tuple(p.data_ptr() if p is not None else 0 for p in packs) + (mode,)

    def _capture(self):
        packs = self.model.packs()
        mode = self.model.math_mode

        fork = torch.cuda.Stream(device=self.device)

        def call():
            # The LSTM chain (conversion, projection, seq-1 recurrent steps) and the TimeConv of the window do not depend on
            # each other: captured on two streams they become parallel branches of the graph, and the ~35 us TimeConv
            # of a 120-clip batch hides behind the launch-bound LSTM chain.  Same kernels, same operands as
            # tmr_head_fwd (whose deferred residual add is the same two fp32 roundings as the GEMM epilogue's).
            cur = torch.cuda.current_stream(self.device)
            if packs[1] is None:
                return ops.head_fwd(*packs, self.x, self.long_feature, self.model.num_class, mode)
            fork.wait_stream(cur)
            with torch.cuda.stream(fork):
                Lt = ops.timeconv_max(packs[1], self.long_feature, mode)
            St = ops.lstm_last(packs[0], self.x, mode)
            cur.wait_stream(fork)
            return ops.relation_head(packs[2], packs[3], St, Lt, self.model.num_class, mode)

        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                     # warm-up outside capture (attribute setup, allocator)
            for _ in range(2):
                call()
        torch.cuda.current_stream(self.device).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = call()
        self._graph, self._out, self._packs = g, out, packs      # keep the packs (and workspace) alive

    def run(self, x, long_feature):
        """x: (B, seq, 2048) or (B*seq, 2048); long_feature: (B, L, 512) - any CUDA tensors (copied into the static buffers
        on the caller's stream) or the static buffers `self.x` / `self.long_feature` themselves, filled in place by the
        producer (no copy).  Returns (logits, pred, score), views of static tensors valid until the next run()."""
        with torch.no_grad(), torch.cuda.device(self.device):
            key = self._capture_key()
            if self._graph is None or key != self._key:
                self._capture()
                self._key = key
            # a caller that fills the static buffers itself (gh.x / gh.long_feature) pays no copy
            if x.data_ptr() != self.x.data_ptr():
                self.x.copy_(x.reshape(self.B, self.seq, F), non_blocking=True)
            if long_feature.data_ptr() != self.long_feature.data_ptr():
                self.long_feature.copy_(long_feature, non_blocking=True)
            self._graph.replay()
        return self._out
