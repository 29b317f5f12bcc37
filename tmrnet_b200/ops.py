"""Tensor-level wrappers over the C ABI (include/tmr_b200.h).  PyTorch here is plumbing: it owns
device memory and the stream; every op hands raw pointers to libtmr_b200.so.  CUDA-only by design:
CPU tensors raise, there is no fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import TMR_MATH_FP32, TMR_MATH_F16, TMR_PAD_REPEAT, TMR_PAD_ZERO, check

D = 512
F = 2048

_default_math = [TMR_MATH_F16]      # tensor cores by default; 'fp32' = CUDA-core exact-order mode
_MATH_NAMES = {"fp32": TMR_MATH_FP32, "f16": TMR_MATH_F16}


def set_math_mode(mode):
    """'fp32' (CUDA-core FFMA, reference-order parity) or 'f16' (tcgen05 tensor cores: fp16 operands,
    fp32 accumulate)."""
    _default_math[0] = _MATH_NAMES[mode] if isinstance(mode, str) else int(mode)


def get_math_mode() -> int:
    return _default_math[0]


def _mode(mode):
    if mode is None:
        return _default_math[0]
    return _MATH_NAMES[mode] if isinstance(mode, str) else int(mode)


def _dev(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: tmrnet_b200 is CUDA-only (got a {t.device} tensor); there is no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------------------------------------
# index tables
# ---------------------------------------------------------------------------------------------
def build_frame2row(list_each_length, seq: int):
    """Host tables for the window gather: (frame2row int32[n_frames], frame2vstart int32[n_frames],
    n_rows).  Closed form of get_useful_start_idx + the start-index dict + the repeat-fill walk."""
    lens = np.ascontiguousarray(np.asarray(list_each_length, dtype=np.int64))
    total = int(lens.sum())
    f2r = np.empty(max(total, 1), dtype=np.int32)
    f2v = np.empty(max(total, 1), dtype=np.int32)
    n_rows = C.c_int64(0)
    check(_lib.load().tmr_build_frame2row(lens.ctypes.data_as(C.c_void_p), len(lens), int(seq),
                                          f2r.ctypes.data_as(C.c_void_p), f2v.ctypes.data_as(C.c_void_p),
                                          C.byref(n_rows)))
    return f2r[:total], f2v[:total], int(n_rows.value)


def gather_windows(bank, frame2row, starts, L: int, frame2vstart=None, pad_mode=TMR_PAD_REPEAT,
                   return_rows=False, validate=False):
    """validate=True: the kernel's status word is read back (one sync) and a start that is not a valid clip
    start of the table raises KeyError like the reference's dict probe (TRAIN:310).  Without it such clips
    come back as all-zero windows; the kernel never reads out of bounds either way."""
    bank = _dev(bank, "bank")
    frame2row = _dev(frame2row, "frame2row", torch.int32)
    starts = _dev(starts, "starts", torch.int64)
    if bank.dim() != 2 or bank.shape[1] != D:
        raise ValueError(f"bank must be (N,{D}), got {tuple(bank.shape)}")
    B = starts.numel()
    out = torch.empty((B, L, D), dtype=torch.float32, device=bank.device)
    rows = torch.empty((B, L), dtype=torch.int32, device=bank.device) if return_rows else None
    if frame2vstart is not None:
        frame2vstart = _dev(frame2vstart, "frame2vstart", torch.int32)
    status = torch.zeros(1, dtype=torch.int32, device=bank.device) if validate else None
    with torch.cuda.device(bank.device):
        check(_lib.load().tmr_gather_windows(_ptr(bank), bank.shape[0], _ptr(frame2row), _ptr(frame2vstart),
                                             frame2row.numel(), _ptr(starts), B, int(L), D, int(pad_mode),
                                             _ptr(out), _ptr(rows), _ptr(status), _stream()))
    if validate:
        code = int(status.item())
        if code & 1:
            n = frame2row.numel()
            s = starts.clamp(0, max(n - 1, 0))
            own = frame2row[s]
            nxt = frame2row[(s + 1).clamp(max=max(n - 1, 0))]
            bad = (starts < 0) | (starts >= n) | (own < 0) | ((s + 1 < n) & (nxt == own))
            raise KeyError(int(starts[bad][0]))
        if code & 2:
            raise IndexError("frame2row names a bank row outside the bank")
    return (out, rows) if return_rows else out


# ---------------------------------------------------------------------------------------------
# weight packs
# ---------------------------------------------------------------------------------------------
def pack_timeconv(w3, b3, w5, b5, w7, b7):
    ts = [_dev(t, "timeconv weight") for t in (w3, b3, w5, b5, w7, b7)]
    lib = _lib.load()
    packed = _ws(lib.tmr_timeconv_packed_bytes(D), ts[0].device)
    with torch.cuda.device(packed.device):
        check(lib.tmr_timeconv_pack(*[_ptr(t) for t in ts], D, _ptr(packed), _stream()))
    return packed


def pack_nlblock(w1, b1, w2, b2, w3, b3, w4, b4, ln_w, ln_b):
    ts = [_dev(t, "nlblock weight") for t in (w1, b1, w2, b2, w3, b3, w4, b4, ln_w, ln_b)]
    lib = _lib.load()
    packed = _ws(lib.tmr_nlblock_packed_bytes(D), ts[0].device)
    with torch.cuda.device(packed.device):
        check(lib.tmr_nlblock_pack(*[_ptr(t) for t in ts], D, _ptr(packed), _stream()))
    return packed


def pack_lstm(w_ih, w_hh, b_ih, b_hh):
    ts = [_dev(t, "lstm weight") for t in (w_ih, w_hh, b_ih, b_hh)]
    lib = _lib.load()
    packed = _ws(lib.tmr_lstm_packed_bytes(F, D), ts[0].device)
    with torch.cuda.device(packed.device):
        check(lib.tmr_lstm_pack(*[_ptr(t) for t in ts], F, D, _ptr(packed), _stream()))
    return packed


def pack_classifier(w_h, b_h, w_c, b_c):
    ts = [_dev(t, "classifier weight") for t in (w_h, b_h, w_c, b_c)]
    num_class = ts[2].shape[0]
    lib = _lib.load()
    nbytes = lib.tmr_classifier_packed_bytes(D, num_class)
    if nbytes == 0:
        raise ValueError(f"unsupported number of classes {num_class}")
    packed = _ws(nbytes, ts[0].device)
    with torch.cuda.device(packed.device):
        check(lib.tmr_classifier_pack(*[_ptr(t) for t in ts], D, num_class, _ptr(packed), _stream()))
    return packed


# ---------------------------------------------------------------------------------------------
# stages
# ---------------------------------------------------------------------------------------------
def timeconv_max(packed, x, math_mode=None):
    x = _dev(x, "x")
    if x.dim() != 3 or x.shape[2] != D:
        raise ValueError(f"TimeConv input must be (B,L,{D}), got {tuple(x.shape)}")
    B, L, _ = x.shape
    out = torch.empty_like(x)
    lib = _lib.load()
    mode = _mode(math_mode)
    ws = _ws(lib.tmr_timeconv_workspace_bytes(B, L, D), x.device) if mode == TMR_MATH_F16 else None
    with torch.cuda.device(x.device):
        check(lib.tmr_timeconv_max_fwd(_ptr(packed), _ptr(x), B, L, D, _ptr(out), _ptr(ws),
                                       ws.numel() if ws is not None else 0, mode, _stream()))
    return out


def nlblock(packed, St, Lt, math_mode=None):
    St = _dev(St, "St")
    Lt = _dev(Lt, "Lt")
    St2 = St.reshape(-1, D)
    if Lt.dim() != 3 or Lt.shape[2] != D or Lt.shape[0] != St2.shape[0]:
        raise ValueError(f"NLBlock: St {tuple(St.shape)} / Lt {tuple(Lt.shape)} mismatch")
    B, L, _ = Lt.shape
    lib = _lib.load()
    out = torch.empty((B, D), dtype=torch.float32, device=St.device)
    ws = _ws(lib.tmr_nlblock_workspace_bytes(B, D), St.device)
    with torch.cuda.device(St.device):
        check(lib.tmr_nlblock_fwd(_ptr(packed), _ptr(St2), _ptr(Lt), B, L, D, _ptr(out), _ptr(ws), ws.numel(),
                                  _mode(math_mode), _stream()))
    return out


def lstm_last(packed, x, math_mode=None):
    x = _dev(x, "x")
    if x.dim() != 3 or x.shape[2] != F:
        raise ValueError(f"LSTM input must be (B,seq,{F}), got {tuple(x.shape)}")
    B, seq, _ = x.shape
    lib = _lib.load()
    out = torch.empty((B, D), dtype=torch.float32, device=x.device)
    ws = _ws(lib.tmr_lstm_workspace_bytes(B * seq, B, D), x.device)
    with torch.cuda.device(x.device):
        check(lib.tmr_lstm_last_fwd(_ptr(packed), _ptr(x), B, seq, F, D, _ptr(out), _ptr(ws), ws.numel(),
                                    _mode(math_mode), _stream()))
    return out


def lstm_seq(packed, x):
    """h of every step, (B, seq, 512) (stage-1 surface, code/models.py:43-45); fp32 CUDA-core path."""
    x = _dev(x, "x")
    if x.dim() != 3 or x.shape[2] != F:
        raise ValueError(f"LSTM input must be (B,seq,{F}), got {tuple(x.shape)}")
    B, seq, _ = x.shape
    lib = _lib.load()
    out_tm = torch.empty((seq, B, D), dtype=torch.float32, device=x.device)
    ws = _ws(lib.tmr_lstm_workspace_bytes(B * seq, B, D), x.device)
    with torch.cuda.device(x.device):
        check(lib.tmr_lstm_seq_fwd(_ptr(packed), _ptr(x), B, seq, F, D, _ptr(out_tm), _ptr(ws), ws.numel(), _stream()))
    return out_tm.permute(1, 0, 2).contiguous()


def lstm_last_frames(packed, feats, starts, seq: int, math_mode=None):
    feats = _dev(feats, "feats")
    starts = _dev(starts, "starts", torch.int64)
    if feats.dim() != 2 or feats.shape[1] != F:
        raise ValueError(f"feats must be (n_frames,{F}), got {tuple(feats.shape)}")
    B = starts.numel()
    lib = _lib.load()
    out = torch.empty((B, D), dtype=torch.float32, device=feats.device)
    ws = _ws(lib.tmr_lstm_workspace_bytes(feats.shape[0], B, D), feats.device)
    with torch.cuda.device(feats.device):
        check(lib.tmr_lstm_last_frames_fwd(_ptr(packed), _ptr(feats), feats.shape[0], _ptr(starts), B, int(seq),
                                           F, D, _ptr(out), _ptr(ws), ws.numel(), _mode(math_mode), _stream()))
    return out


def fc_argmax(packed, St, y1, num_class: int, math_mode=None, want_pred=True):
    St = _dev(St, "St")
    y1 = _dev(y1, "y1")
    B = St.shape[0]
    lib = _lib.load()
    logits = torch.empty((B, num_class), dtype=torch.float32, device=St.device)
    pred = torch.empty((B,), dtype=torch.int64, device=St.device) if want_pred else None
    score = torch.empty((B,), dtype=torch.float32, device=St.device) if want_pred else None
    ws = _ws(lib.tmr_classifier_workspace_bytes(B, D), St.device)
    with torch.cuda.device(St.device):
        check(lib.tmr_fc_argmax_fwd(_ptr(packed), _ptr(St), _ptr(y1), B, D, int(num_class), _ptr(logits), _ptr(pred),
                                    _ptr(score), _ptr(ws), ws.numel(), _mode(math_mode), _stream()))
    return logits, pred, score


def relation_head(nlblock_packed, classifier_packed, St, Lt, num_class: int, math_mode=None):
    """Everything of resnet_lstm.forward after the LSTM and the TimeConv (TRAIN:245-252, eval): NLBlock(St, Lt), the
    classifier, softmax score and argmax - ONE launch for batches of up to 512 clips in the tensor-core mode."""
    St = _dev(St, "St").reshape(-1, D)
    Lt = _dev(Lt, "Lt")
    if Lt.dim() != 3 or Lt.shape[2] != D or Lt.shape[0] != St.shape[0]:
        raise ValueError(f"relation_head: St {tuple(St.shape)} / Lt {tuple(Lt.shape)} mismatch")
    B, L, _ = Lt.shape
    lib = _lib.load()
    logits = torch.empty((B, num_class), dtype=torch.float32, device=St.device)
    pred = torch.empty((B,), dtype=torch.int64, device=St.device)
    score = torch.empty((B,), dtype=torch.float32, device=St.device)
    ws = _ws(lib.tmr_relation_head_workspace_bytes(B, D), St.device)
    with torch.cuda.device(St.device):
        check(lib.tmr_relation_head_fwd(_ptr(nlblock_packed), _ptr(classifier_packed), _ptr(St), _ptr(Lt), B, L, D,
                                        int(num_class), _ptr(logits), _ptr(pred), _ptr(score), _ptr(ws), ws.numel(),
                                        _mode(math_mode), _stream()))
    return logits, pred, score


def head_fwd(lstm_packed, timeconv_packed, nlblock_packed, classifier_packed, x, long_feature, num_class: int,
             math_mode=None):
    x = _dev(x, "x")
    long_feature = _dev(long_feature, "long_feature")
    B, seq, _ = x.shape
    L = long_feature.shape[1]
    if long_feature.shape[0] != B or long_feature.shape[2] != D or x.shape[2] != F:
        raise ValueError(f"head: x {tuple(x.shape)} / long_feature {tuple(long_feature.shape)} mismatch")
    lib = _lib.load()
    logits = torch.empty((B, num_class), dtype=torch.float32, device=x.device)
    pred = torch.empty((B,), dtype=torch.int64, device=x.device)
    score = torch.empty((B,), dtype=torch.float32, device=x.device)
    ws = _ws(lib.tmr_head_workspace_bytes(B, seq, L, D), x.device)
    with torch.cuda.device(x.device):
        check(lib.tmr_head_fwd(_ptr(lstm_packed), _ptr(timeconv_packed), _ptr(nlblock_packed), _ptr(classifier_packed),
                               _ptr(x), _ptr(long_feature), B, seq, L, F, D, int(num_class), _ptr(logits), _ptr(pred),
                               _ptr(score), _ptr(ws), ws.numel(), _mode(math_mode), _stream()))
    return logits, pred, score


def linear(a, w, bias=None, relu=False, math_mode=None):
    a = _dev(a, "a")
    w = _dev(w, "w")
    if bias is not None:
        bias = _dev(bias, "bias")
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if _mode(math_mode) == TMR_MATH_F16:      # the entry point takes fp16 operands in tensor-core mode
        a, w = a.half().contiguous(), w.half().contiguous()
    with torch.cuda.device(a.device):
        check(_lib.load().tmr_linear_fwd(_ptr(a), _ptr(w), _ptr(bias), M, N, K, _ptr(out), int(relu), _mode(math_mode),
                                         _stream()))
    return out


def bankconv(timeconv_packed, bank, row_base: int = 0, pb_rows: int = None):
    """Bank-level TimeConv (tensor-core mode only): (pb_rows, 7, 512) variants per bank row, as float16."""
    bank = _dev(bank, "bank")
    if pb_rows is None:
        pb_rows = bank.shape[0] - row_base
    lib = _lib.load()
    pb = torch.empty((pb_rows, 7, D), dtype=torch.float16, device=bank.device)
    ws = _ws(lib.tmr_bankconv_workspace_bytes(pb_rows, D), bank.device)
    with torch.cuda.device(bank.device):
        check(lib.tmr_bankconv_fwd(_ptr(timeconv_packed), _ptr(bank), bank.shape[0], int(row_base), int(pb_rows), D,
                                   _ptr(pb), _ptr(ws), ws.numel(), _stream()))
    return pb


def attention(u, Lt):
    """softmax-over-L weighted sum for a folded query u (B,512) over Lt (B,L,512) -> (B,512)."""
    u = _dev(u, "u")
    Lt = _dev(Lt, "Lt")
    B, L, _ = Lt.shape
    out = torch.empty((B, D), dtype=torch.float32, device=u.device)
    with torch.cuda.device(u.device):
        check(_lib.load().tmr_attention_fwd(_ptr(u), _ptr(Lt), B, L, D, _ptr(out), _stream()))
    return out
