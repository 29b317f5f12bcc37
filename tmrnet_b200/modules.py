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
        self._backbone_state = {}

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

    def load_reference_state_dict(self, sd, strict=True):
        """Load a reference stage-2 checkpoint (torch.save(model.module.state_dict()), TRAIN:1053,1056): the head's
        keys must match exactly (strict, like EVAL:443-447); the backbone's `share.*` tensors are not used by the head
        and are kept aside untouched so save_reference_checkpoint() can write them back."""
        self._backbone_state = {k: v for k, v in sd.items() if k.startswith(("share.", "res."))}
        own = {k: v for k, v in sd.items() if not k.startswith(("share.", "res."))}
        return self.load_state_dict(own, strict=strict)

    def load_stage1_state_dict(self, sd):
        """What the reference does before stage-2 training (TRAIN:772-774): load a STAGE-1 checkpoint with
        strict=False - only `share.*` (kept aside here) and `lstm.*` match; the stage-1 classifier `fc.*` has no
        counterpart in the head (`fc_c` is 512 -> C on the 512-d fc_h_c output, a different layer) and is dropped.
        Returns (missing head keys, dropped keys) like load_state_dict(strict=False)."""
        self._backbone_state = {k: v for k, v in sd.items() if k.startswith(("share.", "res."))}
        own = dict(self.state_dict())
        take = {k: v for k, v in sd.items() if k in own and tuple(v.shape) == tuple(own[k].shape)}
        dropped = [k for k in sd if k not in take and not k.startswith(("share.", "res."))]
        res = self.load_state_dict(take, strict=False)
        return list(res.missing_keys), dropped

    def reference_state_dict(self, backbone_state=None):
        """State dict in the reference's layout and key order (share.* first, then lstm, fc_c, fc_h_c, nl_block,
        time_conv - the attribute order of TRAIN:209-230): `share.*` from `backbone_state` or from the checkpoint
        this model was loaded from.  With them the reference's strict load (EVAL:443-447) accepts the file; without
        a backbone only strict=False loads (TRAIN:774) do."""
        from collections import OrderedDict
        out = OrderedDict()
        bb = getattr(self, "_backbone_state", {}) if backbone_state is None else backbone_state
        for k, v in bb.items():
            out[k if k.startswith("share.") else "share." + k.split(".", 1)[1]] = v
        for k, v in self.state_dict().items():
            out[k] = v.detach().cpu()
        return out

    def save_reference_checkpoint(self, path, backbone_state=None):
        """torch.save(model.module.state_dict(), path) as TRAIN:1053,1056 writes it."""
        torch.save(self.reference_state_dict(backbone_state), path)

    # -- forward --------------------------------------------------------------------------
    def forward(self, x, long_feature=None):
        """eval() under torch.no_grad(): the inference kernels (tensor cores by default).  train(), or eval() with
        autograd recording: the fp32 training forward as one autograd node (tmrnet_b200.train.HeadTrainFunction;
        dropout 0.2 / 0.5 as NLB:18,38 and TRAIN:228,250 in train(), off in eval()), so the reference's loop body
        `outputs = model.forward(inputs, long_feature); loss = criterion(outputs, labels); loss.backward();
        optimizer.step()` (TRAIN:876-887) runs on this module with stock torch losses and optimisers."""
        if long_feature is None:
            raise TypeError("resnet_lstm.forward: long_feature is required")
        needs_graph = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if self.training or needs_graph:
            from .train import head_train_forward
            if not needs_graph:            # train() under no_grad: dropout active, nothing recorded
                with torch.no_grad():
                    return head_train_forward(self, x, long_feature, dropout=self.training)
            return head_train_forward(self, x, long_feature, dropout=self.training)
        x = x.reshape(-1, self.sequence_length, 2048)
        logits, _, _ = ops.head_fwd(*self.packs(), x, long_feature, self.num_class, self.math_mode)
        return logits

    def predict(self, x, long_feature):
        """forward + the eval scripts' Softmax/torch.max post-processing (eval ...resnest.py:491-493):
        returns (logits, pred int64, score fp32) as device tensors."""
        _no_autograd(x, long_feature, *self.parameters())
        x = x.reshape(-1, self.sequence_length, 2048)
        return ops.head_fwd(*self.packs(), x, long_feature, self.num_class, self.math_mode)
