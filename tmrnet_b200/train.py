"""Head training step, the reference's inner training loop body
(code/Training TMRNet/train_non-local_mutiConv_resnet.py:856-887) for the head with frozen backbone
features: forward in training mode -> CrossEntropyLoss(reduction='sum'[, weight]) -> backward ->
[one NCCL all-reduce (SUM) of the flat fp32 head gradient over NVLink when world_size > 1] ->
torch.optim.SGD-equivalent update (momentum 0.9, weight decay 5e-4, dampening 0; LSTM at lr/10 like
the reference's parameter groups, :797-805).  Forward/backward/update are hand-written CUDA
(csrc/train.cu) behind the C ABI; torch.distributed is only the collective's plumbing.

The loss is sum-reduced, so SUMMING gradients over ranks reproduces the single-GPU gradient of the
union batch; every rank then applies the identical update.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .ops import D, F, _dev, _ptr, _stream, _ws, check

PARAM_ORDER = [
    "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
    "time_conv.timeconv1.weight", "time_conv.timeconv1.bias", "time_conv.timeconv2.weight", "time_conv.timeconv2.bias",
    "time_conv.timeconv3.weight", "time_conv.timeconv3.bias",
    "nl_block.linear1.weight", "nl_block.linear1.bias", "nl_block.linear2.weight", "nl_block.linear2.bias",
    "nl_block.linear3.weight", "nl_block.linear3.bias", "nl_block.linear4.weight", "nl_block.linear4.bias",
    "nl_block.layer_norm.weight", "nl_block.layer_norm.bias",
    "fc_h_c.weight", "fc_h_c.bias", "fc_c.weight", "fc_c.bias",
]


class FlatBuffer:
    """One contiguous fp32 buffer with per-tensor views (the all-reduce bucket)."""

    def __init__(self, shapes, device):
        self.sizes = [int(torch.Size(s).numel()) if s is not None else 0 for s in shapes]
        self.offsets = [0]
        for n in self.sizes:
            self.offsets.append(self.offsets[-1] + (n + 3) // 4 * 4)      # keep every view 16-byte aligned
        self.flat = torch.zeros(self.offsets[-1], dtype=torch.float32, device=device)
        self.views = [self.flat[o:o + n].view(s) if s is not None else None
                      for o, n, s in zip(self.offsets, self.sizes, shapes)]


class HeadTrainer:
    """model: tmrnet_b200.resnet_lstm on a CUDA device.  One `step()` = TRAIN:856-887 for the head."""

    def __init__(self, model, lr=5e-4, momentum=0.9, weight_decay=5e-4, lstm_lr_scale=0.1, class_weight=None,
                 p_nl=0.2, p_fc=0.5, seed=0, process_group=None):
        self.model = model
        named = dict(model.named_parameters())
        self.params = [named.get(k) for k in PARAM_ORDER]
        if any(p is None for i, p in enumerate(self.params) if not PARAM_ORDER[i].startswith("time_conv.")):
            raise ValueError("model is missing head parameters")
        self.has_tc = self.params[4] is not None
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("HeadTrainer is CUDA-only (no CPU fallback)")
        for p in self.params:
            if p is not None and (p.dtype != torch.float32 or not p.is_contiguous()):
                raise TypeError("head parameters must be contiguous fp32")
        shapes = [tuple(p.shape) if p is not None else None for p in self.params]
        self.grads = FlatBuffer(shapes, dev)
        self.momentum_buf = FlatBuffer(shapes, dev)
        self.lr, self.momentum, self.weight_decay, self.lstm_lr_scale = lr, momentum, weight_decay, lstm_lr_scale
        self.class_weight = None if class_weight is None else torch.as_tensor(class_weight, dtype=torch.float32, device=dev)
        self.p_nl, self.p_fc, self.seed = float(p_nl), float(p_fc), int(seed)
        self.steps = 0
        self.pg = process_group
        self._pp = (C.c_void_p * 24)(*[p.data_ptr() if p is not None else 0 for p in self.params])
        self._gp = (C.c_void_p * 24)(*[g.data_ptr() if g is not None else 0 for g in self.grads.views])
        self._ws = None
        self.device = dev

    @property
    def num_grad_elements(self):
        return sum(self.grads.sizes)

    def forward_backward(self, x, long_feature, labels, dropout=True):
        """Fills self.grads (local, not reduced).  Returns (loss tensor (1,), logits (B,C), pred (B,))."""
        x = _dev(x, "x").reshape(-1, self.model.sequence_length, F)
        long_feature = _dev(long_feature, "long_feature")
        labels = _dev(labels, "labels", torch.int64)
        B, seq, _ = x.shape
        L = long_feature.shape[1]
        Cn = self.model.num_class
        if long_feature.shape[0] != B or labels.numel() != B:
            raise ValueError("batch size mismatch between x, long_feature and labels")
        lib = _lib.load()
        need = lib.tmr_head_train_workspace_bytes(B, seq, L, D, F, Cn)
        if self._ws is None or self._ws.numel() < need:
            self._ws = _ws(need, self.device)
        logits = torch.empty((B, Cn), dtype=torch.float32, device=self.device)
        loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        pred = torch.empty(B, dtype=torch.int64, device=self.device)
        p_nl, p_fc = (self.p_nl, self.p_fc) if dropout else (0.0, 0.0)
        with torch.cuda.device(self.device):
            check(lib.tmr_head_train_fwd_bwd(self._pp, self._gp, _ptr(x), _ptr(long_feature), _ptr(labels),
                                             _ptr(self.class_weight), B, seq, L, F, D, Cn, p_nl, p_fc,
                                             self.seed * 1000003 + self.steps, _ptr(logits), _ptr(loss), _ptr(pred),
                                             _ptr(self._ws), self._ws.numel(), _stream()))
        return loss, logits, pred

    def allreduce_grads(self):
        """The only collective of the system: SUM of the flat fp32 head gradient (about 43 MB) over NCCL."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1:
            dist.all_reduce(self.grads.flat, op=dist.ReduceOp.SUM, group=self.pg)

    def sgd_update(self):
        lib = _lib.load()
        first = 1 if self.steps == 0 else 0
        with torch.cuda.device(self.device):
            for i, p in enumerate(self.params):
                if p is None:
                    continue
                lr = self.lr * (self.lstm_lr_scale if PARAM_ORDER[i].startswith("lstm.") else 1.0)
                check(lib.tmr_sgd_step(_ptr(p), _ptr(self.grads.views[i]), _ptr(self.momentum_buf.views[i]), p.numel(),
                                       lr, self.momentum, self.weight_decay, first, _stream()))
        self.steps += 1
        self.model.invalidate_packs()          # weights changed under the inference caches

    def step(self, x, long_feature, labels):
        loss, logits, pred = self.forward_backward(x, long_feature, labels)
        self.allreduce_grads()
        self.sgd_update()
        return loss, logits, pred
