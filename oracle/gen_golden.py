"""Generate tests/golden/* by EXECUTING THE REFERENCE (run in the build container only).

    python oracle/gen_golden.py            # needs /root/reference; writes tests/golden/

The reference has no golden vectors of its own, so the fixtures are
outputs of the reference code itself:
  * `NLBlock`, `TimeConv` imported by file path from
    code/Training TMRNet/NLBlock_MutiConv6_3.py:10-79,
  * `get_long_feature`, `get_useful_start_idx` pulled out of
    code/Training TMRNet/train_non-local_mutiConv_resnet.py:288-326 with ast + exec (the script is
    not importable: top-level argparse, comet_ml),
  * `torch.nn.LSTM` / `nn.Linear` wired exactly as train_non-local_mutiConv_resnet.py:237-253
    minus the backbone (`share`), in eval mode.
Inputs and weights come from tmrnet_b200.synth (pure functions of a seed) so only seeds and the
reference OUTPUTS are stored.  Test infrastructure; never imported by the product.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tmrnet_b200 import synth  # noqa: E402

REF = "/root/reference/code"
NLB_PATH = os.path.join(REF, "Training TMRNet", "NLBlock_MutiConv6_3.py")
TRAIN_PATH = os.path.join(REF, "Training TMRNet", "train_non-local_mutiConv_resnet.py")
OUT = os.path.join(ROOT, "tests", "golden")


def load_reference_modules():
    spec = importlib.util.spec_from_file_location("ref_nlblock", NLB_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.NLBlock, mod.TimeConv


def load_reference_functions(L):
    src = open(TRAIN_PATH).read()
    tree = ast.parse(src)
    wanted = {"get_long_feature", "get_useful_start_idx"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    ns = {"LFB_lENGTH": L, "np": np}
    exec(compile(ast.Module(body=body, type_ignores=[]), TRAIN_PATH, "exec"), ns)
    return ns


def gather_kats():
    """Known-answer tests for the window walk: lfb[row] = [row] so the payload IS the row id."""
    rng = np.random.default_rng(20260101)
    cases = []
    arrays = {}
    fixed = [  # SURVEY.md 8c survey-time KAT first
        dict(seq=4, L=12, lengths=[7, 6, 5]),
        dict(seq=10, L=30, lengths=[45, 12, 10, 70]),
        dict(seq=10, L=30, lengths=[9, 40, 3, 25]),          # videos shorter than seq
        dict(seq=10, L=120, lengths=[60, 50, 200]),
        dict(seq=1, L=5, lengths=[3, 4]),
    ]
    for _ in range(60):
        seq = int(rng.integers(1, 12))
        L = int(rng.choice([1, 3, 10, 30, 40, 60]))
        nv = int(rng.integers(1, 7))
        lengths = [int(v) for v in rng.integers(max(1, seq - 2), 90, size=nv)]
        if sum(max(0, n - seq + 1) for n in lengths) == 0:
            lengths.append(seq + 5)
        fixed.append(dict(seq=seq, L=L, lengths=lengths))
    for c in fixed:
        ns = load_reference_functions(c["L"])
        starts = ns["get_useful_start_idx"](c["seq"], c["lengths"])
        d = {s: r for r, s in enumerate(starts)}          # TRAIN:643-644
        lfb = np.arange(len(starts), dtype=np.float64)[:, None]
        lf = ns["get_long_feature"](starts, d, lfb)
        rows = np.array(lf)[:, :, 0].astype(np.int64)
        i = len(cases)
        cases.append((c["seq"], c["L"]))
        arrays[f"c{i}_lengths"] = np.array(c["lengths"], np.int64)
        arrays[f"c{i}_starts"] = np.array(starts, np.int64)
        arrays[f"c{i}_rows"] = rows.astype(np.int32)
    arrays["meta"] = np.array(cases, np.int64)                 # (n_cases, 2) = (seq, L)
    np.savez_compressed(os.path.join(OUT, "gather_kat.npz"), **arrays)
    print("gather_kat.npz:", len(cases), "cases")


def head_golden():
    NLBlock, TimeConv = load_reference_modules()
    torch.set_num_threads(1)
    seed, seq, L, B = 1234, 10, 30, 4
    out = {}
    for C in (7, 8):
        sd = synth.head_state_dict(num_class=C, seed=seed)
        tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
        lstm = torch.nn.LSTM(2048, 512, batch_first=True)
        lstm.load_state_dict({k[5:]: v for k, v in tsd.items() if k.startswith("lstm.")})
        tc = TimeConv()
        tc.load_state_dict({k[10:]: v for k, v in tsd.items() if k.startswith("time_conv.")})
        nl = NLBlock()
        nl.load_state_dict({k[9:]: v for k, v in tsd.items() if k.startswith("nl_block.")})
        fc_h_c = torch.nn.Linear(1024, 512)
        fc_h_c.load_state_dict({"weight": tsd["fc_h_c.weight"], "bias": tsd["fc_h_c.bias"]})
        fc_c = torch.nn.Linear(512, C)
        fc_c.load_state_dict({"weight": tsd["fc_c.weight"], "bias": tsd["fc_c.bias"]})
        for m in (lstm, tc, nl, fc_h_c, fc_c):
            m.eval()

        lengths = [47, 23, 64]
        ns = load_reference_functions(L)
        starts = ns["get_useful_start_idx"](seq, lengths)
        d = {s: r for r, s in enumerate(starts)}
        feats = synth.features(sum(lengths), seed=seed)
        bank = synth.bank(len(starts), seed=seed).astype(np.float64)   # reference bank is float64
        pick = [starts[0], starts[17], starts[40], starts[-1]]        # video starts + interior
        assert len(pick) == B
        x = np.stack([feats[s:s + seq] for s in pick])                 # (B, seq, 2048)
        lf = ns["get_long_feature"](pick, d, bank)
        long_feature = torch.Tensor(np.array(lf))                      # TRAIN:873-876
        with torch.no_grad():
            # TRAIN:237-253 minus share
            xt = torch.from_numpy(x).view(-1, seq, 2048)
            y, _ = lstm(xt)
            y = y.contiguous().view(-1, 512)
            y = y[seq - 1::seq]
            Lt = tc(long_feature)
            y_1 = nl(y, Lt)
            z = torch.cat([y, y_1], dim=1)
            z = torch.nn.functional.relu(fc_h_c(z))
            logits = fc_c(z)
            # NL-only wiring (train_only_non-local_pretrained.py:226-240): Lt = long_feature
            y_1n = nl(y, long_feature)
            zn = torch.cat([y, y_1n], dim=1)
            logits_nlonly = fc_c(torch.nn.functional.relu(fc_h_c(zn)))
            prob = torch.nn.Softmax(dim=1)(logits)                     # EVAL:491-493
            score, pred = torch.max(prob, 1)
        if C == 7:
            out.update(pick=np.array(pick, np.int64), lengths=np.array(lengths, np.int64),
                       long_feature=long_feature.numpy(), St=y.numpy(), Lt=Lt.numpy(),
                       y1=y_1.numpy(), y1_nlonly=y_1n.numpy())
            out["weight_checksum"] = np.array(
                [float(np.float64(v).sum()) for k, v in sorted(sd.items())], np.float64)
        out[f"logits_c{C}"] = logits.numpy()
        out[f"logits_nlonly_c{C}"] = logits_nlonly.numpy()
        out[f"score_c{C}"] = score.numpy()
        out[f"pred_c{C}"] = pred.numpy().astype(np.int64)
    out["meta"] = np.array([seed, seq, L, B], np.int64)
    np.savez_compressed(os.path.join(OUT, "head_b4_l30.npz"), **out)
    print("head_b4_l30.npz written; logits c7:\n", out["logits_c7"])


def timeconv_l_sweep_check():
    """The L-parametrised TimeConv restatement in the oracle must equal the reference at L=30
    (the only L the reference supports, NLBlock_MutiConv6_3.py:57-77)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tmr_oracle as orc
    _, TimeConv = load_reference_modules()
    sd = synth.head_state_dict(seed=77)
    tc = TimeConv()
    tc.load_state_dict({k[10:]: torch.from_numpy(v) for k, v in sd.items() if k.startswith("time_conv.")})
    x = torch.from_numpy(synth.bank(5 * 30, seed=5).reshape(5, 30, 512))
    with torch.no_grad():
        ref = tc(x)
    got = orc.timeconv(x, sd)
    print("TimeConv oracle vs reference max|d| =", float((ref - got).abs().max()))
    assert torch.equal(ref, got) or float((ref - got).abs().max()) < 1e-6


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gather_kats()
    head_golden()
    timeconv_l_sweep_check()
