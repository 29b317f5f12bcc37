"""Host-side mirror of the reference's module surface for the temporal-memory-relation head.

Same class names, constructor defaults, state-dict keys and initialisers as the reference
(`NLBlock`, `TimeConv`: code/Training TMRNet/NLBlock_MutiConv6_3.py:10-79; `resnet_lstm`,
`resnet_lstm_LFB`: code/Training TMRNet/train_non-local_mutiConv_resnet.py:208-285), so a reference
`.pth` loads with `load_state_dict(..., strict=False)` (backbone `share.*` keys are ignored: the
head takes precomputed 2048-d features instead of images).  torch.nn modules are used only as
PARAMETER CONTAINERS that give the reference's key names/shapes; every forward runs the sm_100a
kernels through the C ABI (tmrnet_b200.ops).  CUDA-only: CPU tensors raise.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.init as init

from . import ops


def _key(params):
    return tuple((p.data_ptr(), p._version, str(p.device)) for p in params)


class _PackedCache:
    """Repacked device weights, rebuilt when any source parameter changes (in-place update, load)."""

    def __init__(self):
        self._k = None
        self._v = None

    def invalidate(self):
        self._k = None

    def get(self, params, fn):
        k = _key(params)
        if k != self._k:
            self._v = fn(*[p.detach() for p in params])
            self._k = k
        return self._v


def _no_autograd(*tensors):
    if torch.is_grad_enabled() and any(t.requires_grad for t in tensors if isinstance(t, torch.Tensor)):
        raise RuntimeError("tmrnet_b200 forward kernels are inference-only: call under torch.no_grad(); "
                           "use tmrnet_b200.train for the head training step")


class NLBlock(nn.Module):
    """Non-local relation block (NLBlock_MutiConv6_3.py:10-40), eval-mode forward."""

    def __init__(self, feature_num=512):
        super().__init__()
        if feature_num != 512:
            raise ValueError("NLBlock: the reference forward hard-codes 512 (view(-1,1,512), LayerNorm([1,512]))")
        self.linear1 = nn.Linear(feature_num, feature_num)
        self.linear2 = nn.Linear(feature_num, feature_num)
        self.linear3 = nn.Linear(feature_num, feature_num)
        self.linear4 = nn.Linear(feature_num, feature_num)
        self.layer_norm = nn.LayerNorm([1, 512])
        self.dropout = nn.Dropout(0.2)      # kept for surface parity; identity in eval
        init.xavier_uniform_(self.linear1.weight)
        init.xavier_uniform_(self.linear2.weight)
        init.xavier_uniform_(self.linear3.weight)
        init.xavier_uniform_(self.linear4.weight)
        self._cache = _PackedCache()
        self.math_mode = None

    def _params(self):
        return [self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                self.linear3.weight, self.linear3.bias, self.linear4.weight, self.linear4.bias,
                self.layer_norm.weight, self.layer_norm.bias]

    def packed(self):
        return self._cache.get(self._params(), ops.pack_nlblock)

    def forward(self, St, Lt):
        _no_autograd(St, Lt, *self._params())
        if self.training:
            raise RuntimeError("NLBlock: dropout(0.2) in training mode is not part of the inference kernels; call .eval()")
        return ops.nlblock(self.packed(), St, Lt, self.math_mode)


class TimeConv(nn.Module):
    """Multi-scale temporal conv + max branch (NLBlock_MutiConv6_3.py:43-79).  Any L (the reference
    hard-codes 30 in its views)."""

    def __init__(self):
        super().__init__()
        self.timeconv1 = nn.Conv1d(512, 512, kernel_size=3, padding=1)
        self.timeconv2 = nn.Conv1d(512, 512, kernel_size=5, padding=2)
        self.timeconv3 = nn.Conv1d(512, 512, kernel_size=7, padding=3)
        self._cache = _PackedCache()
        self.math_mode = None

    def _params(self):
        return [self.timeconv1.weight, self.timeconv1.bias, self.timeconv2.weight, self.timeconv2.bias,
                self.timeconv3.weight, self.timeconv3.bias]

    def packed(self):
        return self._cache.get(self._params(), ops.pack_timeconv)

    def forward(self, x):
        _no_autograd(x, *self._params())
        return ops.timeconv_max(self.packed(), x, self.math_mode)


class resnet_lstm_LFB(nn.Module):
    """Bank-builder head half (train_non-local_mutiConv_resnet.py:256-285 minus `share`):
    features (B,seq,2048) -> LSTM h at the last step (B,512)."""

    def __init__(self, sequence_length=10):
        super().__init__()
        self.sequence_length = sequence_length
        self.lstm = nn.LSTM(2048, 512, batch_first=True)
        init.xavier_normal_(self.lstm.all_weights[0][0])
        init.xavier_normal_(self.lstm.all_weights[0][1])
        self._cache = _PackedCache()
        self.math_mode = None

    def _params(self):
        return [self.lstm.weight_ih_l0, self.lstm.weight_hh_l0, self.lstm.bias_ih_l0, self.lstm.bias_hh_l0]

    def packed(self):
        return self._cache.get(self._params(), ops.pack_lstm)

    def forward(self, x):
        _no_autograd(x, *self._params())
        x = x.reshape(-1, self.sequence_length, 2048)
        return ops.lstm_last(self.packed(), x, self.math_mode)


class resnet_lstm(nn.Module):
    """TMRNet head (train_non-local_mutiConv_resnet.py:208-253 / eval twin
    test_singlenet_phase_non-local_pretrained_2fc_copy_mutiConv6_resnest.py:85-126) minus the
    backbone: forward(x, long_feature) with x = backbone features (B,seq,2048) or (B*seq,2048).

    num_class: 7 (Cholec80; reference resnest/eval scripts), 6 (lucieDLE resnet variant), 8 (M2CAI).
    use_timeconv=False gives the NL-only wiring of train_only_non-local_pretrained.py:226-240."""

    def __init__(self, num_class=7, sequence_length=10, use_timeconv=True):
        super().__init__()
        self.sequence_length = sequence_length
        self.num_class = num_class
        self.lstm = nn.LSTM(2048, 512, batch_first=True)
        self.fc_c = nn.Linear(512, num_class)
        self.fc_h_c = nn.Linear(1024, 512)
        self.nl_block = NLBlock()
        self.dropout = nn.Dropout(p=0.5)    # surface parity; identity in eval
        self.time_conv = TimeConv() if use_timeconv else None
        init.xavier_normal_(self.lstm.all_weights[0][0])
        init.xavier_normal_(self.lstm.all_weights[0][1])
        init.xavier_uniform_(self.fc_c.weight)
        init.xavier_uniform_(self.fc_h_c.weight)
        self._lstm_cache = _PackedCache()
        self._cls_cache = _PackedCache()
        self.math_mode = None

    # -- packed weights -------------------------------------------------------------------
    def _lstm_params(self):
        return [self.lstm.weight_ih_l0, self.lstm.weight_hh_l0, self.lstm.bias_ih_l0, self.lstm.bias_hh_l0]

    def _cls_params(self):
        return [self.fc_h_c.weight, self.fc_h_c.bias, self.fc_c.weight, self.fc_c.bias]

    def packs(self):
        return (self._lstm_cache.get(self._lstm_params(), ops.pack_lstm),
                self.time_conv.packed() if self.time_conv is not None else None,
                self.nl_block.packed(),
                self._cls_cache.get(self._cls_params(), ops.pack_classifier))

    def invalidate_packs(self):
        """Call after parameters were updated in place by non-torch code (tmrnet_b200.train)."""
        self._lstm_cache.invalidate()
        self._cls_cache.invalidate()
        self.nl_block._cache.invalidate()
        if self.time_conv is not None:
            self.time_conv._cache.invalidate()

    def load_reference_state_dict(self, sd):
        """Load a reference checkpoint (torch.save(model.module.state_dict()),
        train_non-local_mutiConv_resnet.py:1053): backbone `share.*` keys are dropped."""
        own = {k: v for k, v in sd.items() if not k.startswith(("share.", "res."))}
        return self.load_state_dict(own, strict=True)

    # -- forward --------------------------------------------------------------------------
    def forward(self, x, long_feature=None):
        if long_feature is None:
            raise TypeError("resnet_lstm.forward: long_feature is required")
        _no_autograd(x, long_feature, *self.parameters())
        if self.training:
            raise RuntimeError("resnet_lstm: dropout in training mode is not part of the inference kernels; call .eval()")
        x = x.reshape(-1, self.sequence_length, 2048)
        logits, _, _ = ops.head_fwd(*self.packs(), x, long_feature, self.num_class, self.math_mode)
        return logits

    def predict(self, x, long_feature):
        """forward + the eval scripts' Softmax/torch.max post-processing (eval ...resnest.py:491-493):
        returns (logits, pred int64, score fp32) as device tensors."""
        _no_autograd(x, long_feature, *self.parameters())
        x = x.reshape(-1, self.sequence_length, 2048)
        return ops.head_fwd(*self.packs(), x, long_feature, self.num_class, self.math_mode)
