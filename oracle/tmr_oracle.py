"""CPU oracle for the TMRNet temporal-memory-relation head.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  Nothing under tmrnet_b200/ imports it; the product path is CUDA-only.

It is a restatement (plain numpy for the index work, torch-CPU fp32 for the arithmetic) of the
reference algorithm, each function citing the reference file:line it follows (paths relative to
/root/reference).  Shorthand:
  NLB   = code/Training TMRNet/NLBlock_MutiConv6_3.py
  TRAIN = code/Training TMRNet/train_non-local_mutiConv_resnet.py
  NLONLY= code/Training TMRNet/train_only_non-local_pretrained.py
  EVAL  = code/eval/python/test_singlenet_phase_non-local_pretrained_2fc_copy_mutiConv6_resnest.py
  EXPORT= code/eval/python/export_phase_copy.py

Parity pin: the reference has no tests or golden vectors of its own, so the
oracle is pinned against outputs of the reference itself, executed in the build container by
oracle/gen_golden.py (imports NLBlock/TimeConv by file path, exec's get_long_feature /
get_useful_start_idx out of the training script) and committed under tests/golden/.
tests/test_oracle_golden.py checks every function here against those fixtures.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as Fn

NL_SCALE = (1 / 512) ** 0.5      # NLB:31 python float, multiplied after the dot product


# --------------------------------------------------------------------------------------
# index plumbing
# --------------------------------------------------------------------------------------
def get_useful_start_idx(sequence_length, list_each_length):
    """TRAIN:288-295 — global frame ids that can start a `sequence_length`-frame clip."""
    count = 0
    idx = []
    for n in list_each_length:
        for j in range(count, count + (n + 1 - sequence_length)):
            idx.append(j)
        count += n
    return idx


def build_start_dict(start_idx):
    """TRAIN:643-644 — {global start frame id -> bank row}."""
    return {int(s): r for r, s in enumerate(start_idx)}


def window_rows(start_index_list, dict_start_idx_LFB, L):
    """TRAIN:298-326 restated on ROW INDICES (the payload is lfb[row]): for k = 0..L-1 the key is
    start-k-1; a key that is a valid clip start yields its row and is remembered; any other key
    repeats the last remembered row (initially the clip's own row).  Time-reversed, repeat-filled,
    leaks into the previous video's tail.  Returns int64 (B, L)."""
    out = np.empty((len(start_index_list), L), dtype=np.int64)
    for b, start in enumerate(start_index_list):
        start = int(start)
        last = dict_start_idx_LFB[start]
        for k in range(L):
            key = start - k - 1
            if key in dict_start_idx_LFB:
                last = dict_start_idx_LFB[key]
            out[b, k] = last
    return out


def get_long_feature(start_index_list, dict_start_idx_LFB, lfb, L):
    """TRAIN:298-326 + :873-876 — (B, L, 512) fp32 window tensor (np.array -> torch.Tensor cast)."""
    rows = window_rows(start_index_list, dict_start_idx_LFB, L)
    return np.asarray(lfb)[rows].astype(np.float32)


def frame2row_closed_form(list_each_length, sequence_length):
    """Closed form of the walk above (extended to videos shorter than seq):
    frame2row[g] = row(g) if g is a valid start else row(smallest valid start > g); frames after
    the last valid start are never queried and hold -1.  window[b,k] = frame2row[s-k-1] if
    s-k-1 >= 0 else 0."""
    total = int(sum(list_each_length))
    valid = np.full(total, -1, dtype=np.int64)
    starts = get_useful_start_idx(sequence_length, list_each_length)
    valid[np.asarray(starts, dtype=np.int64)] = np.arange(len(starts), dtype=np.int64)
    f2r = valid.copy()
    nxt = -1
    for g in range(total - 1, -1, -1):
        if valid[g] >= 0:
            nxt = valid[g]
        f2r[g] = nxt
    return f2r


def window_rows_closed_form(starts, frame2row, L):
    starts = np.asarray(starts, dtype=np.int64)
    keys = starts[:, None] - np.arange(1, L + 1, dtype=np.int64)[None, :]
    rows = np.where(keys >= 0, frame2row[np.maximum(keys, 0)], 0)
    return rows.astype(np.int64)


# --------------------------------------------------------------------------------------
# arithmetic (torch CPU, fp32 unless dtype=torch.float64 is requested for error references)
# --------------------------------------------------------------------------------------
def _t(x, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        return x.detach().to("cpu", dtype)
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)


def timeconv(x, sd, dtype=torch.float32, prefix="time_conv."):
    """NLB:43-79 with the literal 30 replaced by L = x.shape[1] (equal to the reference at L=30,
    checked against the golden fixture).  x (B, L, 512) -> (B, L, 512):
        out[b,k,c] = max(x[b,k,c], max(x[b,k,c], k>0 ? x[b,k-1,c] : 0), conv3, conv5, conv7)."""
    x = _t(x, dtype)
    xt = x.transpose(1, 2)                                     # (B, 512, L)            NLB:53
    ys = [xt]
    for i, k in ((1, 3), (2, 5), (3, 7)):                       # NLB:55-65
        w = _t(sd[f"{prefix}timeconv{i}.weight"], dtype)
        b = _t(sd[f"{prefix}timeconv{i}.bias"], dtype)
        ys.append(Fn.conv1d(xt, w, b, padding=k // 2))
    x4 = Fn.pad(xt, (1, 0), mode="constant", value=0)           # NLB:67
    ys.append(Fn.max_pool1d(x4, 2, stride=1))                   # NLB:68
    y = torch.stack(ys, dim=0).max(dim=0).values               # NLB:76-77 (max over the 5 branches)
    return y.transpose(1, 2).contiguous()


def nlblock(St, Lt, sd, dtype=torch.float32, prefix="nl_block."):
    """NLB:25-40 in eval mode (dropout is the identity)."""
    St = _t(St, dtype)
    Lt = _t(Lt, dtype)
    W = lambda i: _t(sd[f"{prefix}linear{i}.weight"], dtype)
    b = lambda i: _t(sd[f"{prefix}linear{i}.bias"], dtype)
    St_1 = Fn.linear(St.view(-1, 1, 512), W(1), b(1))           # NLB:26-27
    Lt_1 = Fn.linear(Lt, W(2), b(2)).transpose(1, 2)            # NLB:28-29
    SL = torch.matmul(St_1, Lt_1) * NL_SCALE                    # NLB:30-31
    SL = Fn.softmax(SL, dim=2)                                  # NLB:32
    Lt_2 = Fn.linear(Lt, W(3), b(3))                            # NLB:33
    SLL = torch.matmul(SL, Lt_2)                                # NLB:34
    SLL = Fn.layer_norm(SLL, [1, 512], _t(sd[f"{prefix}layer_norm.weight"], dtype),
                        _t(sd[f"{prefix}layer_norm.bias"], dtype), 1e-5)   # NLB:35
    SLL = Fn.relu(SLL)                                          # NLB:36
    SLL = Fn.linear(SLL, W(4), b(4))                            # NLB:37
    return St + SLL.view(-1, 512)                               # NLB:39-40


def lstm_last(x, sd, dtype=torch.float32, prefix="lstm."):
    """TRAIN:224,241-244 — torch.nn.LSTM(2048,512,batch_first) semantics from zero state, gate
    order i,f,g,o, both biases added; returns h at the last step, (B, 512)."""
    x = _t(x, dtype)
    Wih = _t(sd[f"{prefix}weight_ih_l0"], dtype)
    Whh = _t(sd[f"{prefix}weight_hh_l0"], dtype)
    bias = _t(sd[f"{prefix}bias_ih_l0"], dtype) + _t(sd[f"{prefix}bias_hh_l0"], dtype)
    B, seq, _ = x.shape
    H = Whh.shape[1]
    h = torch.zeros(B, H, dtype=dtype)
    c = torch.zeros(B, H, dtype=dtype)
    for t in range(seq):
        g = Fn.linear(x[:, t], Wih) + Fn.linear(h, Whh) + bias
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
    return h


def classifier(St, y1, sd, dtype=torch.float32):
    """TRAIN:249-252 in eval mode / EVAL:122-125 — fc_h_c(cat) -> relu -> fc_c."""
    z = torch.cat([_t(St, dtype), _t(y1, dtype)], dim=1)
    z = Fn.relu(Fn.linear(z, _t(sd["fc_h_c.weight"], dtype), _t(sd["fc_h_c.bias"], dtype)))
    return Fn.linear(z, _t(sd["fc_c.weight"], dtype), _t(sd["fc_c.bias"], dtype))


def head(x_feat, long_feature, sd, use_timeconv=True, dtype=torch.float32):
    """TRAIN:237-253 minus `share` (features precomputed); NLONLY:226-240 when use_timeconv=False.
    Returns (logits, St, Lt, y1)."""
    St = lstm_last(x_feat, sd, dtype)
    Lt = timeconv(long_feature, sd, dtype) if use_timeconv else _t(long_feature, dtype)
    y1 = nlblock(St, Lt, sd, dtype)
    return classifier(St, y1, sd, dtype), St, Lt, y1


def eval_postproc(logits):
    """EVAL:491-493 — softmax over classes, (score, pred) = max (first index on ties)."""
    p = Fn.softmax(_t(logits), dim=1)
    score, pred = torch.max(p, 1)
    return score, pred


def export_phase_lines(preds, list_each_length, sequence_length, fps=25):
    """EXPORT:43-75 — per-video lists of '<fps*k>\\t<phase>' lines; the first seq-1 frames of each
    video get phase 0 (EXPORT:55-60)."""
    out = []
    p = 0
    for n in list_each_length:
        phases = [0] * (sequence_length - 1) + [int(v) for v in preds[p:p + n - sequence_length + 1]]
        p += n - sequence_length + 1
        out.append([f"{fps * k}\t{ph}" for k, ph in enumerate(phases)])
    return out
