"""Downstream of the path: prediction export in the reference's wire format
(code/eval/python/export_phase_copy.py:43-75) so outputs drop into the MATLAB evaluation unchanged.

One `<fps*k>\\t<phase>` line per frame of a video; the first `seq-1` frames, which no clip ends on,
get phase 0 (export_phase_copy.py:55-60).  The reference's sanity check
`num_labels == num_preds + (seq-1)*num_video` (:32) is enforced."""
from __future__ import annotations

import os

import numpy as np
import torch


def phase_lines(preds, list_each_length, sequence_length: int = 10, fps: int = 25):
    """preds: per-clip argmax in global clip order (tensor / array / list).  Returns a list (one
    entry per video) of lists of '<fps*k>\\t<phase>' strings."""
    if isinstance(preds, torch.Tensor):
        preds = preds.detach().cpu().numpy()          # one vectorised D2H instead of per-element .cpu()
    preds = np.asarray(preds).astype(np.int64).reshape(-1)
    lengths = [int(v) for v in list_each_length]
    usable = [n for n in lengths if n >= sequence_length]
    num_labels = sum(usable)
    if len(usable) != len(lengths):
        raise ValueError("videos shorter than the clip length produce no predictions")
    if num_labels != len(preds) + (sequence_length - 1) * len(lengths):
        raise ValueError("number error, please check: num_labels %d != num_preds %d + (seq-1)*num_video %d"
                         % (num_labels, len(preds), (sequence_length - 1) * len(lengths)))
    out, p = [], 0
    for n in lengths:
        k = n - (sequence_length - 1)
        phases = np.concatenate([np.zeros(sequence_length - 1, np.int64), preds[p:p + k]])
        p += k
        out.append([f"{fps * i}\t{int(ph)}" for i, ph in enumerate(phases)])
    return out


def export_phase_files(preds, list_each_length, out_dir, sequence_length: int = 10, first_video: int = 41,
                       fps: int = 25):
    """Write video{N}-phase.txt files like export_phase_copy.py:43-75 (videos 41..80 for Cholec80's
    test split).  Returns the paths."""
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for i, lines in enumerate(phase_lines(preds, list_each_length, sequence_length, fps)):
        path = os.path.join(out_dir, f"video{first_video + i}-phase.txt")
        with open(path, "w") as f:
            f.write("\n".join(lines) + "\n")
        paths.append(path)
    return paths
