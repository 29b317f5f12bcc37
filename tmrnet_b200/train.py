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
from .ops import D, F, TMR_MATH_F16, TMR_MATH_FP32, _dev, _ptr, _stream, _ws, check

def _train_mode(mode) -> int:
    """Training math mode: 'fp32' (default: exact-parity CUDA-core GEMMs, what the reference's fp32 training does) or
    'f16' (fp16-rounded GEMM operands on the tensor cores, fp32 accumulation - see include/tmr_b200.h)."""
    if mode is None or mode in ("fp32", TMR_MATH_FP32):
        return TMR_MATH_FP32
    if mode in ("f16", "fp16", TMR_MATH_F16):
        return TMR_MATH_F16
    raise ValueError(f"unknown training math mode {mode!r} (use 'fp32' or 'f16')")


PARAM_ORDER = [
    "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
    "time_conv.timeconv1.weight", "time_conv.timeconv1.bias", "time_conv.timeconv2.weight", "time_conv.timeconv2.bias",
    "time_conv.timeconv3.weight", "time_conv.timeconv3.bias",
    "nl_block.linear1.weight", "nl_block.linear1.bias", "nl_block.linear2.weight", "nl_block.linear2.bias",
    "nl_block.linear3.weight", "nl_block.linear3.bias", "nl_block.linear4.weight", "nl_block.linear4.bias",
    "nl_block.layer_norm.weight", "nl_block.layer_norm.bias",
    "fc_h_c.weight", "fc_h_c.bias", "fc_c.weight", "fc_c.bias",
]


def _param_list(model):
    named = dict(model.named_parameters())
    params = [named.get(k) for k in PARAM_ORDER]
    if any(p is None for i, p in enumerate(params) if not PARAM_ORDER[i].startswith("time_conv.")):
        raise ValueError("model is missing head parameters")
    return params


class HeadTrainFunction(torch.autograd.Function):
    """Training-mode forward of the whole head as ONE autograd node: forward = tmr_head_train_fwd (activations stay
    in a workspace the node owns), backward = tmr_head_train_bwd on the incoming dlogits.  This is what makes the
    reference's own loop body work on the module (TRAIN:876-887) with stock torch losses / optimisers / schedulers:

        model.train(); outputs = model.forward(x, long_feature); loss = criterion(outputs, labels)
        optimizer.zero_grad(); loss.backward(); optimizer.step()

    x (backbone features) and long_feature (bank windows) receive no gradient: the head trains on frozen
    features."""

    @staticmethod
    def forward(ctx, x, long_feature, meta, *params):
        seq, num_class, p_nl, p_fc, seed, mode = meta
        dev = x.device
        B, L = x.shape[0], long_feature.shape[1]
        lib = _lib.load()
        ws = _ws(lib.tmr_head_train_workspace_bytes(B, seq, L, D, F, num_class), dev)
        logits = torch.empty((B, num_class), dtype=torch.float32, device=dev)
        pp = (C.c_void_p * 24)(*[p.data_ptr() if p is not None else 0 for p in params])
        with torch.cuda.device(dev):
            check(lib.tmr_head_train_fwd(pp, _ptr(x), _ptr(long_feature), B, seq, L, F, D, num_class, float(p_nl), float(p_fc),
                                         int(seed), _ptr(logits), _ptr(ws), ws.numel(), mode, _stream()))
        ctx.mode = mode
        ctx.save_for_backward(x, long_feature, *[p for p in params if p is not None])
        ctx.present = [p is not None for p in params]
        ctx.shapes = [tuple(p.shape) if p is not None else None for p in params]
        ctx.ws, ctx.dims = ws, (B, seq, L, num_class)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        saved = ctx.saved_tensors            # raises if a parameter was modified in place since the forward
        x, long_feature = saved[0], saved[1]
        it = iter(saved[2:])
        params = [next(it) if present else None for present in ctx.present]
        B, seq, L, num_class = ctx.dims
        dev = x.device
        grads = FlatBuffer(ctx.shapes, dev)
        dlogits = dlogits.to(torch.float32).contiguous()
        lib = _lib.load()
        pp = (C.c_void_p * 24)(*[p.data_ptr() if p is not None else 0 for p in params])
        gp = (C.c_void_p * 24)(*[g.data_ptr() if g is not None else 0 for g in grads.views])
        scratch = torch.empty((B, num_class), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.tmr_head_train_bwd(pp, gp, _ptr(x), _ptr(long_feature), _ptr(dlogits), B, seq, L, F, D, num_class,
                                         _ptr(scratch), _ptr(ctx.ws), ctx.ws.numel(), ctx.mode, _stream()))
        ctx.ws = None
        return (None, None, None) + tuple(grads.views)


def head_train_forward(model, x, long_feature, dropout=True):
    """resnet_lstm.forward in training mode (or in eval mode with autograd on: dropout off)."""
    if x.requires_grad or long_feature.requires_grad:
        raise RuntimeError("tmrnet_b200: the head trains on frozen backbone features and bank rows; x / long_feature "
                           "cannot require grad (fine-tuning the backbone is outside this path)")
    x = _dev(x, "x").reshape(-1, model.sequence_length, F)
    long_feature = _dev(long_feature, "long_feature")
    if long_feature.dim() != 3 or long_feature.shape[0] != x.shape[0] or long_feature.shape[2] != D:
        raise ValueError(f"head: x {tuple(x.shape)} / long_feature {tuple(long_feature.shape)} mismatch")
    params = _param_list(model)
    for prm in params:
        if prm is not None and (prm.dtype != torch.float32 or not prm.is_contiguous() or prm.device != x.device):
            raise TypeError("head parameters must be contiguous fp32 tensors on the input's device")
    # dropout masks: counter-based generator keyed by torch's seed and a per-model call counter
    model._train_calls = getattr(model, "_train_calls", 0) + 1
    seed = (torch.initial_seed() * 1000003 + model._train_calls) & 0x7FFFFFFFFFFFFFFF
    p_nl = model.nl_block.dropout.p if dropout else 0.0
    p_fc = model.dropout.p if dropout else 0.0
    meta = (model.sequence_length, model.num_class, p_nl, p_fc, seed, _train_mode(getattr(model, "train_math_mode", None)))
    return HeadTrainFunction.apply(x, long_feature, meta, *params)


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
                 p_nl=0.2, p_fc=0.5, seed=0, process_group=None, math_mode=None):
        self.model = model
        self.math_mode = _train_mode(math_mode)
        self.params = _param_list(model)
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
        self.lstm_lr = lr * lstm_lr_scale       # the reference's second parameter group (TRAIN:797-805); schedulers move both
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
        if self.class_weight is not None and self.class_weight.numel() != Cn:
            raise ValueError(f"class_weight has {self.class_weight.numel()} entries for {Cn} classes")
        # labels outside [0, C) would index the logits / class weights out of bounds in the fused loss kernel; torch's
        # CrossEntropyLoss raises for them - so does this, as a device-side assert (no host synchronisation)
        torch._assert_async(((labels >= 0) & (labels < Cn)).all())
        # parameters may have been moved / reassigned since construction (model.to(), load_state_dict keeps storage)
        if any(p is not None and p.data_ptr() != int(self._pp[i] or 0) for i, p in enumerate(self.params)):
            named = dict(self.model.named_parameters())
            self.params = [named.get(k) for k in PARAM_ORDER]
            if any(p is not None and (p.device != self.device or p.dtype != torch.float32 or not p.is_contiguous())
                   for p in self.params):
                raise RuntimeError("HeadTrainer: parameters moved to another device / dtype after construction")
            self._pp = (C.c_void_p * 24)(*[p.data_ptr() if p is not None else 0 for p in self.params])
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
                                             _ptr(self._ws), self._ws.numel(), self.math_mode, _stream()))
        return loss, logits, pred

    def allreduce_grads(self):
        """The only collective of the system: SUM of the flat fp32 head gradient (about 43 MB) over NCCL."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1:
            dist.all_reduce(self.grads.flat, op=dist.ReduceOp.SUM, group=self.pg)

    def sgd_update(self):
        lib = _lib.load()
        first = 1 if self.steps == 0 else 0
        n = len(self.params)
        ptr = lambda t: t.data_ptr() if t is not None else 0
        pp = (C.c_void_p * n)(*[ptr(p) for p in self.params])
        gp = (C.c_void_p * n)(*[ptr(g) for g in self.grads.views])
        bp = (C.c_void_p * n)(*[ptr(b) for b in self.momentum_buf.views])
        sizes = (C.c_int64 * n)(*[p.numel() if p is not None else 0 for p in self.params])
        lrs = (C.c_float * n)(*[self.lstm_lr if PARAM_ORDER[i].startswith("lstm.") else self.lr for i in range(n)])
        with torch.cuda.device(self.device):       # all 24 tensors in one launch, each with its group's learning rate
            check(lib.tmr_sgd_step_multi(pp, gp, bp, sizes, lrs, n, self.momentum, self.weight_decay, first, _stream()))
        self.steps += 1
        self.model.invalidate_packs()          # weights changed under the inference caches

    def step(self, x, long_feature, labels):
        loss, logits, pred = self.forward_backward(x, long_feature, labels)
        self.allreduce_grads()
        self.sgd_update()
        return loss, logits, pred


# ---------------------------------------------------------------------------------------------
# Host control flow around the step (TRAIN:780-812, 925-1056): learning-rate schedules on the
# validation loss, per-epoch bookkeeping, best-model selection.  Plain Python on purpose - this is
# the reference's epoch loop with the head's forward/backward/update swapped for HeadTrainer.step().
# ---------------------------------------------------------------------------------------------
class PlateauLR:
    """torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, 'min') as the reference constructs it (TRAIN:808-809:
    every default - factor 0.1, patience 10, relative threshold 1e-4, no cooldown, min_lr 0, eps 1e-8), acting
    on HeadTrainer.lr.  The reference's two parameter groups (lr/10 for the LSTM, lr for the rest) are both
    multiplied by `factor`, which is what scaling the trainer's base lr does.  step(metric) after every epoch
    (TRAIN:1023-1027)."""

    def __init__(self, trainer, factor=0.1, patience=10, threshold=1e-4, cooldown=0, min_lr=0.0, eps=1e-8):
        if factor >= 1.0:
            raise ValueError("Factor should be < 1.0.")
        self.trainer, self.factor, self.patience, self.threshold = trainer, factor, patience, threshold
        self.cooldown, self.min_lr, self.eps = cooldown, min_lr, eps
        self.best = float("inf")
        self.num_bad_epochs = 0
        self.cooldown_counter = 0
        self.last_epoch = 0

    def _is_better(self, a):
        return a < self.best * (1.0 - self.threshold)

    def step(self, metric):
        current = float(metric)
        self.last_epoch += 1
        if self._is_better(current):
            self.best = current
            self.num_bad_epochs = 0
        else:
            self.num_bad_epochs += 1
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.num_bad_epochs = 0
        if self.num_bad_epochs > self.patience:
            # torch applies the eps rule to every parameter group on its own: the LSTM group (lr/10) stops
            # decaying one decade before the other group does
            for attr in ("lr", "lstm_lr"):
                old = getattr(self.trainer, attr)
                new_lr = max(old * self.factor, self.min_lr)
                if old - new_lr > self.eps:
                    setattr(self.trainer, attr, new_lr)
            self.cooldown_counter = self.cooldown
            self.num_bad_epochs = 0
        return self.trainer.lr


class StepLR:
    """lr_scheduler.StepLR(optimizer, step_size, gamma) (TRAIN:806-807), acting on HeadTrainer.lr."""

    def __init__(self, trainer, step_size, gamma=0.1):
        self.trainer, self.step_size, self.gamma = trainer, int(step_size), float(gamma)
        self.base_lr, self.base_lstm_lr = trainer.lr, trainer.lstm_lr
        self.last_epoch = 0

    def step(self, metric=None):
        self.last_epoch += 1
        k = self.gamma ** (self.last_epoch // self.step_size) if self.step_size > 0 else 1.0
        self.trainer.lr = self.base_lr * k
        self.trainer.lstm_lr = self.base_lstm_lr * k
        return self.trainer.lr


def per_phase_precision_recall(labels, preds, num_class=None):
    """sklearn.metrics.precision_score / recall_score(average=None) as TRAIN:983-984 uses them: one value per class
    that occurs in labels or preds (sorted), 0 where the denominator is empty."""
    import numpy as np
    labels = np.asarray(labels, dtype=np.int64).ravel()
    preds = np.asarray(preds, dtype=np.int64).ravel()
    classes = np.unique(np.concatenate([labels, preds])) if num_class is None else np.arange(num_class)
    prec, rec = [], []
    for c in classes:
        tp = float(np.sum((preds == c) & (labels == c)))
        pp, ap = float(np.sum(preds == c)), float(np.sum(labels == c))
        prec.append(tp / pp if pp > 0 else 0.0)
        rec.append(tp / ap if ap > 0 else 0.0)
    return np.asarray(prec), np.asarray(rec)


class BestModelTracker:
    """TRAIN:1029-1039: keep the weights of the epoch with the highest validation accuracy; on an exact tie, the one
    with the higher training accuracy.  checkpoint_name() reproduces the reference's file stem (TRAIN:1041-1052)."""

    def __init__(self):
        self.best_val_acc = 0.0
        self.correspond_train_acc = 0.0
        self.best_epoch = 0
        self.best_state = None

    def update(self, epoch, train_acc, val_acc, state_dict_fn):
        took = False
        if val_acc > self.best_val_acc:
            self.best_val_acc, self.correspond_train_acc, self.best_epoch = val_acc, train_acc, epoch
            took = True
        if val_acc == self.best_val_acc and train_acc > self.correspond_train_acc:
            self.correspond_train_acc, self.best_epoch = train_acc, epoch
            took = True
        if took:
            self.best_state = {k: v.detach().clone() for k, v in state_dict_fn().items()}
        return took

    def checkpoint_name(self, seq, train_bs, optimizer_choice=0, multi_optim=1, use_flip=1, crop_type=1):
        save_val = int("{:4.0f}".format(self.best_val_acc * 10000))
        save_train = int("{:4.0f}".format(self.correspond_train_acc * 10000))
        return ("lstm_epoch_" + str(self.best_epoch) + "_length_" + str(seq) + "_opt_" + str(optimizer_choice)
                + "_mulopt_" + str(multi_optim) + "_flip_" + str(use_flip) + "_crop_" + str(crop_type)
                + "_batch_" + str(train_bs) + "_train_" + str(save_train) + "_val_" + str(save_val))


def fit(trainer, index, feats, bank, labels, train_starts, val_starts, epochs, batch_clips=120, L=30, seed=0,
        scheduler=None, val_index=None, val_feats=None, val_bank=None, val_labels=None, log=None):
    """The reference's epoch loop for the head (TRAIN:815-1056) over precomputed features and a cached bank:
    shuffled clip starts -> get_long_feature -> HeadTrainer.step; validation in eval mode with the same
    sum-reduced loss; scheduler.step(val loss); best-model tracking.  `labels` are per FRAME; a clip's label is
    that of its last frame (TRAIN:838, 942).  Returns (history list of per-epoch dicts, BestModelTracker).
    The validation split defaults to the training tensors (index / feats / bank / labels) when not given."""
    import numpy as np
    from . import lfb
    model = trainer.model
    seq = model.sequence_length
    dev = trainer.device
    rng = np.random.default_rng(seed)
    val_index = index if val_index is None else val_index
    val_feats = feats if val_feats is None else val_feats
    val_bank = bank if val_bank is None else val_bank
    val_labels = labels if val_labels is None else val_labels
    labels_t = torch.as_tensor(labels, dtype=torch.int64, device=dev)
    val_labels_t = torch.as_tensor(val_labels, dtype=torch.int64, device=dev)
    train_starts = np.asarray(train_starts, dtype=np.int64)
    val_starts = np.asarray(val_starts, dtype=np.int64)
    frame_off = torch.arange(seq, device=dev)
    tracker = BestModelTracker()
    history = []

    def clip_features(f, starts_dev):
        return f[(starts_dev[:, None] + frame_off[None, :]).reshape(-1)].reshape(len(starts_dev), seq, F)

    for epoch in range(epochs):
        model.train()
        order = rng.permutation(train_starts)
        tr_loss, tr_correct = 0.0, 0
        for lo in range(0, len(order), batch_clips):
            s = order[lo:lo + batch_clips]
            s_dev = torch.from_numpy(s).to(dev)
            lf = lfb.get_long_feature(s, index, bank, L)
            y = labels_t[s_dev + (seq - 1)]
            loss, _, pred = trainer.step(clip_features(feats, s_dev), lf, y)
            tr_loss += float(loss)
            tr_correct += int((pred == y).sum())
        model.eval()
        va_loss, va_correct, va_preds, va_labels = 0.0, 0, [], []
        for lo in range(0, len(val_starts), batch_clips):
            s = val_starts[lo:lo + batch_clips]
            s_dev = torch.from_numpy(s).to(dev)
            lf = lfb.get_long_feature(s, val_index, val_bank, L)
            y = val_labels_t[s_dev + (seq - 1)]
            with torch.no_grad():
                logits, pred, _ = model.predict(clip_features(val_feats, s_dev), lf)
            logp = torch.log_softmax(logits.double(), dim=1)
            w = trainer.class_weight.double()[y] if trainer.class_weight is not None else 1.0
            va_loss += float(-(w * logp[torch.arange(len(y), device=dev), y]).sum())
            va_correct += int((pred == y).sum())
            va_preds.append(pred.cpu()); va_labels.append(y.cpu())
        n_tr, n_va = max(1, len(train_starts)), max(1, len(val_starts))
        prec, rec = per_phase_precision_recall(torch.cat(va_labels).numpy(), torch.cat(va_preds).numpy()) if va_labels else ([], [])
        rec_e = dict(epoch=epoch, lr=trainer.lr, train_loss=tr_loss / n_tr, train_acc=tr_correct / n_tr,
                     val_loss=va_loss / n_va, val_acc=va_correct / n_va,
                     val_precision_each_phase=prec, val_recall_each_phase=rec)
        if scheduler is not None:
            scheduler.step(rec_e["val_loss"])
        rec_e["best"] = tracker.update(epoch, rec_e["train_acc"], rec_e["val_acc"], model.state_dict)
        history.append(rec_e)
        if log is not None:
            log(rec_e)
    return history, tracker
