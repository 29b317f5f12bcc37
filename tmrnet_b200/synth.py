"""Seeded synthetic Cholec80-/M2CAI-shaped inputs for the TMRNet head (numpy only).

Everything here is a pure function of its seed so that the committed golden fixtures
(tests/golden/, produced by oracle/gen_golden.py from the reference modules), the oracle,
the CUDA path and bench.py all see bit-identical inputs on any machine.  numpy's PCG64
`Generator.random` / `standard_normal` streams are used directly (no torch RNG).

Shapes follow SURVEY.md section 8(d):
  videos    V lengths in [1500, 2500)            (Cholec80-shaped, V=40)
  features  (N_frames, 2048) fp32 = relu(N(0,1)) * 0.5   (post-avgpool ResNet features are >= 0)
  bank      (N_clips, 512)  fp32 = tanh(N(0,1)) * 0.5    (LSTM outputs live in (-1, 1))
  weights   reference initialisers restated: xavier_normal_ on LSTM ih/hh, xavier_uniform_ on
            linears/FCs (train_non-local_mutiConv_resnet.py:232-235, NLBlock_MutiConv6_3.py:20-23),
            torch defaults (uniform +-1/sqrt(fan_in)) for conv weights/biases and linear/LSTM biases.
            LayerNorm affine is perturbed away from (1, 0) so parity tests exercise it.
"""
from __future__ import annotations

import numpy as np

D = 512          # bank row width / LSTM hidden
F = 2048         # backbone feature width


def video_lengths(num_videos: int = 40, lo: int = 1500, hi: int = 2500, seed: int = 1234):
    rng = np.random.default_rng(seed)
    return [int(v) for v in rng.integers(lo, hi, size=num_videos)]


def features(num_frames: int, seed: int = 1234, width: int = F) -> np.ndarray:
    rng = np.random.default_rng(seed + 1)
    x = rng.standard_normal((num_frames, width), dtype=np.float32)
    np.maximum(x, 0.0, out=x)
    x *= np.float32(0.5)
    return x


def bank(num_rows: int, seed: int = 1234, width: int = D) -> np.ndarray:
    rng = np.random.default_rng(seed + 2)
    x = rng.standard_normal((num_rows, width), dtype=np.float32)
    np.tanh(x, out=x)
    x *= np.float32(0.5)
    return x


def _uniform(rng, shape, bound):
    return ((rng.random(shape, dtype=np.float32) * 2.0 - 1.0) * np.float32(bound)).astype(np.float32)


def _normal(rng, shape, std):
    return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)


def head_state_dict(num_class: int = 7, seed: int = 1234, with_timeconv: bool = True) -> dict:
    """State dict (numpy fp32) with the reference's key names and shapes."""
    rng = np.random.default_rng(seed + 3)
    sd = {}
    # nn.LSTM(2048, 512): xavier_normal_ on weights, default uniform(-1/sqrt(H), 1/sqrt(H)) on biases
    sd["lstm.weight_ih_l0"] = _normal(rng, (4 * D, F), (2.0 / (4 * D + F)) ** 0.5)
    sd["lstm.weight_hh_l0"] = _normal(rng, (4 * D, D), (2.0 / (4 * D + D)) ** 0.5)
    sd["lstm.bias_ih_l0"] = _uniform(rng, (4 * D,), D ** -0.5)
    sd["lstm.bias_hh_l0"] = _uniform(rng, (4 * D,), D ** -0.5)
    if with_timeconv:
        for i, k in ((1, 3), (2, 5), (3, 7)):
            bound = (D * k) ** -0.5
            sd[f"time_conv.timeconv{i}.weight"] = _uniform(rng, (D, D, k), bound)
            sd[f"time_conv.timeconv{i}.bias"] = _uniform(rng, (D,), bound)
    for i in (1, 2, 3, 4):
        sd[f"nl_block.linear{i}.weight"] = _uniform(rng, (D, D), (6.0 / (2 * D)) ** 0.5)
        sd[f"nl_block.linear{i}.bias"] = _uniform(rng, (D,), D ** -0.5)
    sd["nl_block.layer_norm.weight"] = (1.0 + 0.1 * rng.standard_normal((1, D), dtype=np.float32)).astype(np.float32)
    sd["nl_block.layer_norm.bias"] = (0.1 * rng.standard_normal((1, D), dtype=np.float32)).astype(np.float32)
    sd["fc_h_c.weight"] = _uniform(rng, (D, 2 * D), (6.0 / (3 * D)) ** 0.5)
    sd["fc_h_c.bias"] = _uniform(rng, (D,), (2 * D) ** -0.5)
    sd["fc_c.weight"] = _uniform(rng, (num_class, D), (6.0 / (D + num_class)) ** 0.5)
    sd["fc_c.bias"] = _uniform(rng, (num_class,), D ** -0.5)
    return sd


def clip_starts(lengths, seq: int):
    """All valid global clip-start frame ids, video by video (same values as the reference's
    get_useful_start_idx, train_non-local_mutiConv_resnet.py:288-295), as int64 numpy."""
    out = []
    base = 0
    for n in lengths:
        if n >= seq:
            out.append(np.arange(base, base + n - seq + 1, dtype=np.int64))
        base += n
    return np.concatenate(out) if out else np.zeros((0,), np.int64)
