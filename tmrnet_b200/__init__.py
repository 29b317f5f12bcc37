"""tmrnet_b200 — B200-native (sm_100a) implementation of TMRNet's temporal memory relation head.

Public surface mirrors the reference (lucieDLE/TMRNet): NLBlock, TimeConv, resnet_lstm,
resnet_lstm_LFB, get_useful_start_idx, get_long_feature.  Everything computes through the C ABI of
libtmr_b200.so (include/tmr_b200.h); there is no CPU path.
"""
from .lfb import (LFBIndex, get_long_feature, get_useful_start_idx, load_bank, save_bank,  # noqa: F401
                  to_device_bank)
from .modules import NLBlock, TimeConv, resnet_lstm, resnet_lstm_LFB  # noqa: F401
from .ops import set_math_mode, get_math_mode  # noqa: F401
from . import models  # noqa: F401  (stage-1 surface of code/models.py: tmrnet_b200.models.resnet_lstm(args, num_class))

__all__ = ["NLBlock", "TimeConv", "resnet_lstm", "resnet_lstm_LFB", "get_useful_start_idx",
           "get_long_feature", "LFBIndex", "load_bank", "save_bank", "to_device_bank",
           "set_math_mode", "get_math_mode"]
