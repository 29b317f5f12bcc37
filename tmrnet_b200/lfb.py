"""Long-range feature bank (LFB) store + window gather: the host-side mirror of the reference's
`get_useful_start_idx`, `dict_*_start_idx_LFB` and `get_long_feature`
(code/Training TMRNet/train_non-local_mutiConv_resnet.py:288-326, 643-644, 752-763).

Same positional signatures; `get_long_feature` returns the (B, L, 512) fp32 window tensor on the
bank's CUDA device instead of a nested Python list (the reference converts with
torch.Tensor(np.array(..)).to(device) right after, :873-876).
"""
from __future__ import annotations

import pickle
import threading

import numpy as np
import torch

from . import ops

LFB_LENGTH_DEFAULT = 30   # reference --LFB_l default in the mutiConv scripts


def get_useful_start_idx(sequence_length, list_each_length):
    """Global frame ids that can start a `sequence_length`-frame clip inside one video, video by
    video — same list as the reference (TRAIN:288-295)."""
    idx = []
    count = 0
    for n in list_each_length:
        n = int(n)
        idx.extend(range(count, count + max(0, n + 1 - sequence_length)))
        count += n
    return idx


class LFBIndex(dict):
    """The reference's `dict_start_idx_LFB` ({global clip-start frame id -> bank row}) plus the
    device tables the gather kernel reads.  It IS a dict, so reference code that probes it keeps
    working; built either from video lengths or from a reference-style dict."""

    def __init__(self, mapping, frame2row, frame2vstart=None):
        super().__init__(mapping)
        self.frame2row_host = np.ascontiguousarray(frame2row, dtype=np.int32)
        self.frame2vstart_host = None if frame2vstart is None else np.ascontiguousarray(frame2vstart, dtype=np.int32)
        self._dev = {}
        n = len(self.frame2row_host)
        self.valid_host = np.zeros(n, dtype=bool)
        if len(self):
            k = np.fromiter(self.keys(), dtype=np.int64, count=len(self))
            self.valid_host[k[(k >= 0) & (k < n)]] = True

    def check_starts(self, starts):
        """Reference behaviour: dict_start_idx_LFB[start] raises KeyError for a frame that cannot
        start a clip (TRAIN:310)."""
        s = np.asarray(starts, dtype=np.int64)
        bad = (s < 0) | (s >= len(self.valid_host))
        bad |= ~self.valid_host[np.clip(s, 0, max(len(self.valid_host) - 1, 0))]
        if bad.any():
            raise KeyError(int(s[bad][0]))

    @classmethod
    def from_lengths(cls, list_each_length, sequence_length):
        f2r, f2v, n_rows = ops.build_frame2row(list_each_length, sequence_length)
        starts = get_useful_start_idx(sequence_length, list_each_length)
        assert len(starts) == n_rows
        return cls(zip(starts, range(n_rows)), f2r, f2v)

    @classmethod
    def from_dict(cls, d):
        """From a plain reference dict: an invalid frame maps to the row of the smallest valid
        start above it (what the reference walk is still remembering when it gets there)."""
        if not d:
            raise ValueError("empty start-index dict")
        keys = np.fromiter(d.keys(), dtype=np.int64, count=len(d))
        vals = np.fromiter(d.values(), dtype=np.int64, count=len(d))
        order = np.argsort(keys)
        keys, vals = keys[order], vals[order]
        frames = np.arange(int(keys[-1]) + 1, dtype=np.int64)
        pos = np.searchsorted(keys, frames, side="left")      # first key >= frame
        return cls(d, vals[pos].astype(np.int32), None)

    def device_tables(self, device):
        device = torch.device(device)
        if device not in self._dev:
            f2r = torch.from_numpy(self.frame2row_host).to(device)
            f2v = None if self.frame2vstart_host is None else torch.from_numpy(self.frame2vstart_host).to(device)
            self._dev[device] = (f2r, f2v)
        return self._dev[device]


_dict_cache = []                  # [(reference dict, its LFBIndex)]: one entry, swapped atomically
_dict_cache_lock = threading.Lock()


def _as_index(dict_start_idx_LFB):
    """The reference passes the same plain dict on every call; its device tables are built once.  The
    reference's DataParallel calls forward from one Python thread per GPU, so the cache is guarded."""
    if isinstance(dict_start_idx_LFB, LFBIndex):
        return dict_start_idx_LFB
    with _dict_cache_lock:
        for d, idx in _dict_cache:
            if d is dict_start_idx_LFB and len(d) == len(idx):
                return idx
        idx = LFBIndex.from_dict(dict_start_idx_LFB)
        _dict_cache[:] = [(dict_start_idx_LFB, idx)]
        return idx


def to_device_bank(lfb, device="cuda"):
    """Reference banks are pickled numpy float64 (N,512) whose values are exact fp32
    (TRAIN:541-542,729-732); cast to fp32 is lossless.  Accepts numpy or torch."""
    if isinstance(lfb, torch.Tensor):
        t = lfb
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(lfb, dtype=np.float32)))
    return t.to(device=device, dtype=torch.float32).contiguous()


def load_bank(path, device="cuda"):
    """Read a reference bank pickle (TRAIN:758-765)."""
    with open(path, "rb") as f:
        return to_device_bank(pickle.load(f), device)


def save_bank(bank, path):
    """Write a bank the reference scripts can read back (TRAIN:752-756): numpy float64 (N,512)."""
    arr = bank.detach().cpu().numpy().astype(np.float64) if isinstance(bank, torch.Tensor) else np.asarray(bank, np.float64)
    with open(path, "wb") as f:
        pickle.dump(arr, f)


def get_long_feature(start_index_list, dict_start_idx_LFB, lfb, LFB_length=LFB_LENGTH_DEFAULT,
                     pad_mode="repeat", return_rows=False, trusted=False):
    """Past-`LFB_length` window of every clip in `start_index_list` (TRAIN:298-326).

    start_index_list: global clip-start frame ids (list / numpy / tensor; e.g. data[2][0::seq]).
    dict_start_idx_LFB: the reference dict or an LFBIndex.
    lfb: the bank, a CUDA fp32 tensor (N,512) (numpy float64 banks are uploaded on every call —
         convert once with to_device_bank()).
    pad_mode 'repeat' = reference semantics; 'zero' = zero rows before the clip's own video
    (needs an LFBIndex built from_lengths).  Returns a CUDA tensor (B, LFB_length, 512).

    A frame that cannot start a clip raises KeyError as the reference's dict probe does (TRAIN:310): host
    starts are checked on the host, CUDA starts by the kernel (its status word is read back: one sync).
    trusted=True skips the read-back for CUDA starts; invalid starts then yield all-zero windows."""
    index = _as_index(dict_start_idx_LFB)
    bank = lfb if isinstance(lfb, torch.Tensor) and lfb.is_cuda else to_device_bank(lfb)
    validate = False
    if isinstance(start_index_list, torch.Tensor) and start_index_list.is_cuda:
        starts = start_index_list.to(device=bank.device, dtype=torch.int64).reshape(-1)
        validate = not trusted
    else:
        host = np.asarray(start_index_list.cpu() if isinstance(start_index_list, torch.Tensor) else start_index_list,
                          dtype=np.int64).reshape(-1)
        index.check_starts(host)
        starts = torch.from_numpy(host).to(bank.device)
    f2r, f2v = index.device_tables(bank.device)
    mode = {"repeat": ops.TMR_PAD_REPEAT, "zero": ops.TMR_PAD_ZERO}[pad_mode]
    if mode == ops.TMR_PAD_ZERO and f2v is None:
        raise ValueError("pad_mode='zero' needs an LFBIndex.from_lengths(...) index (video boundaries)")
    return ops.gather_windows(bank, f2r, starts, int(LFB_length), f2v, mode, return_rows, validate=validate)
