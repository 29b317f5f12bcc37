"""ctypes binding of libtmr_b200.so (include/tmr_b200.h).  No torch types cross this boundary: only
raw device pointers, sizes and the CUDA stream handle.  There is no CPU fallback — if the library
cannot be loaded the import of any op fails loudly."""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtmr_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "tmr_b200.h")

TMR_MATH_FP32 = 0
TMR_MATH_F16 = 1
TMR_PAD_REPEAT = 0
TMR_PAD_ZERO = 1

_p, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t

# name -> (restype, argtypes); must list every function include/tmr_b200.h declares
SIGNATURES = {
    "tmr_last_error": (C.c_char_p, []),
    "tmr_version": (_i, []),
    "tmr_device_arch": (_i, []),
    "tmr_build_frame2row": (_i, [_p, _i, _i, _p, _p, _p]),
    "tmr_gather_windows": (_i, [_p, _i64, _p, _p, _i64, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "tmr_timeconv_packed_bytes": (_sz, [_i]),
    "tmr_timeconv_pack": (_i, [_p] * 6 + [_i, _p, _p]),
    "tmr_nlblock_packed_bytes": (_sz, [_i]),
    "tmr_nlblock_pack": (_i, [_p] * 10 + [_i, _p, _p]),
    "tmr_lstm_packed_bytes": (_sz, [_i, _i]),
    "tmr_lstm_pack": (_i, [_p] * 4 + [_i, _i, _p, _p]),
    "tmr_classifier_packed_bytes": (_sz, [_i, _i]),
    "tmr_classifier_pack": (_i, [_p] * 4 + [_i, _i, _p, _p]),
    "tmr_timeconv_workspace_bytes": (_sz, [_i, _i, _i]),
    "tmr_timeconv_max_fwd": (_i, [_p, _p, _i, _i, _i, _p, _p, _sz, _i, _p]),
    "tmr_attention_fwd": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "tmr_nlblock_workspace_bytes": (_sz, [_i, _i]),
    "tmr_nlblock_fwd": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _sz, _i, _p]),
    "tmr_lstm_workspace_bytes": (_sz, [_i64, _i, _i]),
    "tmr_lstm_last_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _sz, _i, _p]),
    "tmr_lstm_last_frames_fwd": (_i, [_p, _p, _i64, _p, _i, _i, _i, _i, _p, _p, _sz, _i, _p]),
    "tmr_lstm_seq_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "tmr_classifier_workspace_bytes": (_sz, [_i, _i]),
    "tmr_fc_argmax_fwd": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _sz, _i, _p]),
    "tmr_relation_head_workspace_bytes": (_sz, [_i, _i]),
    "tmr_relation_head_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _i, _p]),
    "tmr_head_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "tmr_head_fwd": (_i, [_p] * 6 + [_i] * 6 + [_p, _p, _p, _p, _sz, _i, _p]),
    "tmr_head_frames_workspace_bytes": (_sz, [_i64, _i, _i, _i]),
    "tmr_head_frames_fwd": (_i, [_p] * 5 + [_i64, _i64, _p, _i64, _p, _p, _i64, _p] + [_i] * 7
                            + [_p, _p, _p, _p, _p, _sz, _i, _p]),
    "tmr_bankconv_workspace_bytes": (_sz, [_i64, _i]),
    "tmr_bankconv_fwd": (_i, [_p, _p, _i64, _i64, _i64, _i, _p, _p, _sz, _p]),
    "tmr_head_frames_dedup_workspace_bytes": (_sz, [_i64, _i, _i, _i, _i64, _i, _i]),
    "tmr_head_frames_dedup_fwd": (_i, [_p] * 5 + [_i, _i64, _i64, _p, _i64, _p, _p, _i64, _p, _i, _p, _p, _i, _p, _i, _i64, _i64]
                                  + [_i] * 6 + [_p, _p, _p, _p, _p, _sz, _p]),
    "tmr_head_train_workspace_bytes": (_sz, [_i] * 6),
    "tmr_head_train_fwd_bwd": (_i, [_p, _p, _p, _p, _p, _p] + [_i] * 6 + [C.c_float, C.c_float, C.c_uint64, _p, _p, _p, _p, _sz, _i, _p]),
    "tmr_head_train_fwd": (_i, [_p, _p, _p] + [_i] * 6 + [C.c_float, C.c_float, C.c_uint64, _p, _p, _sz, _i, _p]),
    "tmr_head_train_bwd": (_i, [_p, _p, _p, _p, _p] + [_i] * 6 + [_p, _p, _sz, _i, _p]),
    "tmr_sgd_step": (_i, [_p, _p, _p, _i64, C.c_float, C.c_float, C.c_float, _i, _p]),
    "tmr_sgd_step_multi": (_i, [_p, _p, _p, _p, _p, _i, C.c_float, C.c_float, _i, _p]),
    "tmr_linear_fwd": (_i, [_p, _p, _p, _i64, _i, _i, _p, _i, _i, _p]),
}

_lib = None


def header_symbols():
    """Function names declared in include/tmr_b200.h (used by the CPU-side ABI test)."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tmr_[a-z0-9_]+)\s*\(", txt)))


def load():
    """Load (building first if the .so is missing or stale and nvcc is present) and type the ABI."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    try:
        stale = _build.needs_build()          # missing, or older than any csrc/ file or the header
        have_nvcc = bool(_build.nvcc_path())
    except RuntimeError:
        have_nvcc = False
    if not os.path.exists(LIB_PATH) or (stale and have_nvcc):
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here = ABI/header drift: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class TmrError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        msg = load().tmr_last_error()
        raise TmrError(f"libtmr_b200 error {rc}: {msg.decode() if msg else '?'}")
