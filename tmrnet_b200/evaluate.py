"""Relaxed-boundary phase-recognition metrics: a faithful Python port of the reference's MATLAB
evaluation (code/eval/result/matlab-eval/Evaluate.m:1-89 and the aggregation of Main.m:22-110),
so accuracy / Jaccard / precision / recall can be checked without MATLAB.

Faithful means bug-for-bug: in Evaluate.m the "early transition" relaxation
`curDiff(curDiff(end-t+1:end)==1) = 0` builds its logical mask from the LAST t samples of a
ground-truth segment but, being a length-t mask, MATLAB applies it to the FIRST t positions; and
the late-transition statement runs first, so the second mask sees its edits.  Both are reproduced.
Labels are 0-based phase ids on input (as in the phase txt files) and 1-based inside, like Main.m:42-45.
"""
from __future__ import annotations

import numpy as np


def _runs(mask):
    """bwconncomp on a 1-D logical vector: [start, end] (inclusive) of every run of True."""
    m = np.concatenate([[0], mask.astype(np.int8), [0]])
    d = np.diff(m)
    return list(zip(np.flatnonzero(d == 1), np.flatnonzero(d == -1) - 1))


def evaluate(gt, pred, fps: int = 1, num_phases: int = 7):
    """Evaluate.m: returns (jaccard[num_phases], precision[num_phases], recall[num_phases], accuracy), all in
    percent, NaN for phases absent from the ground truth."""
    gt = np.asarray(gt, dtype=np.int64) + 1
    pred = np.asarray(pred, dtype=np.int64) + 1
    if gt.shape != pred.shape:
        raise ValueError("Ground truth and prediction have different sizes")
    oriT = 10 * fps
    diff = pred - gt
    upd_len = 0
    updated = np.zeros(len(gt), dtype=np.int64)
    for ph in range(1, num_phases + 1):
        for s, e in _runs(gt == ph):
            cur = diff[s:e + 1].copy()
            t = min(oriT, len(cur))
            head = cur[:t]                                   # a VIEW of the first t samples
            if ph in (4, 5):
                head[cur[:t] == -1] = 0
                tail = cur[len(cur) - t:]
                head[(tail == 1) | (tail == 2)] = 0          # mask from the tail, applied to the head (MATLAB quirk)
            elif ph in (6, 7):
                head[(cur[:t] == -1) | (cur[:t] == -2)] = 0
                tail = cur[len(cur) - t:]
                head[(tail == 1) | (tail == 2)] = 0
            else:
                head[cur[:t] == -1] = 0
                tail = cur[len(cur) - t:]
                head[tail == 1] = 0
            updated[s:e + 1] = cur
            upd_len = max(upd_len, e + 1)
    jac, prec, rec = [], [], []
    with np.errstate(divide="ignore", invalid="ignore"):
        for ph in range(1, num_phases + 1):
            g = gt == ph
            if not g.any():
                jac.append(np.nan); prec.append(np.nan); rec.append(np.nan)
                continue
            union = g | (pred == ph)
            tp = int(np.sum(updated[union] == 0))
            jac.append(tp / int(union.sum()) * 100.0)
            prec.append(np.float64(tp) * 100.0 / np.float64((pred == ph).sum()))
            rec.append(np.float64(tp) * 100.0 / np.float64(g.sum()))
    acc = float(np.sum(updated[:upd_len] == 0)) / len(gt) * 100.0
    return np.array(jac), np.array(prec), np.array(rec), acc


def evaluate_videos(gts, preds, fps: int = 1, num_phases: int = 7):
    """Main.m:22-110 aggregation over videos.  gts/preds: lists of per-video 0-based label arrays."""
    J, P, R, A = [], [], [], []
    for g, p in zip(gts, preds):
        j, pr, r, a = evaluate(g, p, fps, num_phases)
        J.append(j); P.append(pr); R.append(r); A.append(a)
    J, P, R = (np.minimum(np.array(x).T, 100.0) for x in (J, P, R))          # (phases, videos), capped at 100
    with np.errstate(invalid="ignore"):
        jp, pp, rp = np.nanmean(J, axis=1), np.nanmean(P, axis=1), np.nanmean(R, axis=1)
    A = np.array(A)
    return {
        "jaccard_per_phase": jp, "precision_per_phase": pp, "recall_per_phase": rp,
        "mean_jaccard": float(np.mean(jp)), "std_jaccard": float(np.std(jp, ddof=1)),
        "mean_precision": float(np.nanmean(pp)), "std_precision": float(np.nanstd(pp, ddof=1)),
        "mean_recall": float(np.mean(rp)), "std_recall": float(np.std(rp, ddof=1)),
        "mean_accuracy": float(A.mean()), "std_accuracy": float(A.std(ddof=1)) if len(A) > 1 else 0.0,
        "accuracy_per_video": A,
    }
