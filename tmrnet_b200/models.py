"""Stage-1 ("memory bank") model surface of the reference, code/models.py:7-69, on precomputed backbone features.

`resnet_lstm(args, num_class)`: same constructor, attribute names (`lstm`, `fc`, `dropout`), initialisers,
state-dict keys and `get_optimizers()` as the reference class; the ResNet-50 trunk (`res.*` there, `share.*` in
the training scripts) is outside this path, so `forward` takes the trunk's OUTPUT - features (B, seq, 2048) or
(B*seq, 2048) - instead of images and returns what the reference returns: per-frame logits (B*seq, C)
(code/models.py:38-48).  Eval-mode forward runs the sm_100a kernels through the C ABI; training this model is
stage 1 of the method, not the temporal-memory-relation head, and raises (OUT).

Its `lstm.*` weights are what fills the memory bank (resnet_lstm_LFB) and what the stage-2 head starts from
(TRAIN:772-774, load_state_dict(strict=False)).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.optim as optim

from . import ops
from .modules import _PackedCache

BACKBONE_PREFIXES = ("res.", "share.")


class resnet_lstm(nn.Module):
    def __init__(self, args=None, num_class=7, sequence_length=None):
        super().__init__()
        self.args = args
        self.num_class = num_class
        self.sequence_length = sequence_length if sequence_length is not None else getattr(args, "seq", 10)
        self.lstm = nn.LSTM(2048, 512, batch_first=True)
        self.fc = nn.Linear(512, num_class)
        self.dropout = nn.Dropout(p=0.2)
        nn.init.xavier_normal_(self.lstm.all_weights[0][0])
        nn.init.xavier_normal_(self.lstm.all_weights[0][1])
        nn.init.xavier_uniform_(self.fc.weight)
        self._cache = _PackedCache()
        self._backbone_state = {}

    def _params(self):
        return [self.lstm.weight_ih_l0, self.lstm.weight_hh_l0, self.lstm.bias_ih_l0, self.lstm.bias_hh_l0]

    def packed(self):
        return self._cache.get(self._params(), ops.pack_lstm)

    def forward(self, x):
        if self.training:
            raise RuntimeError("tmrnet_b200.models.resnet_lstm: training the stage-1 model is outside the temporal-memory-"
                               "relation head (SURVEY.md section 2); call .eval()")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("stage-1 forward is inference-only: call under torch.no_grad()")
        x = x.reshape(-1, self.sequence_length, 2048)
        y = ops.lstm_seq(self.packed(), x)                       # (B, seq, 512): every step's h (code/models.py:43-45)
        y = y.reshape(-1, 512)                                   # dropout is the identity in eval
        return ops.linear(y, self.fc.weight.detach(), self.fc.bias.detach(), math_mode="fp32")

    def get_optimizers(self):
        """code/models.py:50-69 without the trunk's parameter group (there are no trunk parameters here)."""
        a = self.args
        if a.opt == 0:
            return optim.SGD([
                {"params": self.lstm.parameters(), "lr": a.lr},
                {"params": self.fc.parameters(), "lr": a.lr},
            ], lr=a.lr / 10, momentum=a.momentum, dampening=a.dampening, weight_decay=a.weightdecay, nesterov=a.nesterov)
        if a.opt == 1:
            return optim.Adam([
                {"params": self.lstm.parameters(), "lr": a.lr},
                {"params": self.fc.parameters(), "lr": a.lr},
            ], lr=a.lr / 10)
        return None

    def load_reference_state_dict(self, sd):
        """A stage-1 checkpoint of the reference (`res.*` per code/models.py or `share.*` per the training scripts,
        `lstm.*`, `fc.*`): the trunk's tensors are kept aside untouched so reference_state_dict() can write them back."""
        self._backbone_state = {k: v for k, v in sd.items() if k.startswith(BACKBONE_PREFIXES)}
        return self.load_state_dict({k: v for k, v in sd.items() if not k.startswith(BACKBONE_PREFIXES)}, strict=True)

    def reference_state_dict(self, backbone_state=None):
        out = dict(self._backbone_state if backbone_state is None else backbone_state)
        out.update({k: v.detach().cpu() for k, v in self.state_dict().items()})
        return out
