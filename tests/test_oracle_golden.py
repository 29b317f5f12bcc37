"""The oracle (oracle/tmr_oracle.py) pinned against fixtures produced by executing the reference
(oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import torch

import tmr_oracle as orc
from tmrnet_b200 import synth


def _kat(golden_dir):
    z = np.load(os.path.join(golden_dir, "gather_kat.npz"))
    for i, (seq, L) in enumerate(z["meta"]):
        yield int(seq), int(L), z[f"c{i}_lengths"].tolist(), z[f"c{i}_starts"], z[f"c{i}_rows"].astype(np.int64)


def test_survey_kat(golden_dir):
    # seq=4, lengths [7,6,5], L=12
    seq, L, lengths, starts, rows = next(_kat(golden_dir))
    assert (seq, L, lengths) == (4, 12, [7, 6, 5])
    assert starts.tolist() == [0, 1, 2, 3, 7, 8, 9, 13, 14]
    by = {int(s): rows[i].tolist() for i, s in enumerate(starts)}
    assert by[0] == [0] * 12
    assert by[3] == [2, 1, 0] + [0] * 9
    assert by[7] == [4, 4, 4, 3, 2, 1, 0, 0, 0, 0, 0, 0]
    assert by[13] == [7, 7, 7, 6, 5, 4, 4, 4, 4, 3, 2, 1]
    assert by[14] == [7, 7, 7, 7, 6, 5, 4, 4, 4, 4, 3, 2]


def test_start_idx_and_window_walk_match_reference(golden_dir):
    n = 0
    for seq, L, lengths, starts, rows in _kat(golden_dir):
        mine = orc.get_useful_start_idx(seq, lengths)
        assert mine == starts.tolist()
        assert synth.clip_starts(lengths, seq).tolist() == starts.tolist()
        d = orc.build_start_dict(mine)
        assert np.array_equal(orc.window_rows(mine, d, L), rows)
        n += 1
    assert n >= 60


def test_closed_form_matches_reference(golden_dir):
    for seq, L, lengths, starts, rows in _kat(golden_dir):
        f2r = orc.frame2row_closed_form(lengths, seq)
        assert np.array_equal(orc.window_rows_closed_form(starts, f2r, L), rows)


def test_get_long_feature_values(golden_dir):
    z = np.load(os.path.join(golden_dir, "head_b4_l30.npz"))
    seed, seq, L, B = (int(v) for v in z["meta"])
    lengths = z["lengths"].tolist()
    starts = orc.get_useful_start_idx(seq, lengths)
    bank = synth.bank(len(starts), seed=seed).astype(np.float64)
    lf = orc.get_long_feature(z["pick"], orc.build_start_dict(starts), bank, L)
    assert lf.dtype == np.float32 and np.array_equal(lf, z["long_feature"])


def test_weights_regenerate_identically(golden_dir):
    z = np.load(os.path.join(golden_dir, "head_b4_l30.npz"))
    sd = synth.head_state_dict(num_class=7, seed=int(z["meta"][0]))
    chk = np.array([float(np.float64(v).sum()) for k, v in sorted(sd.items())])
    assert np.array_equal(chk, z["weight_checksum"])


def test_head_stages_match_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "head_b4_l30.npz"))
    seed, seq, L, B = (int(v) for v in z["meta"])
    feats = synth.features(int(z["lengths"].sum()), seed=seed)
    x = np.stack([feats[s:s + seq] for s in z["pick"]])
    torch.set_num_threads(1)
    for C in (7, 8):
        sd = synth.head_state_dict(num_class=C, seed=seed)
        logits, St, Lt, y1 = orc.head(x, z["long_feature"], sd)
        if C == 7:
            assert np.allclose(St.numpy(), z["St"], rtol=0, atol=2e-6)
            assert np.array_equal(Lt.numpy(), z["Lt"])
            assert np.allclose(y1.numpy(), z["y1"], rtol=0, atol=5e-6)
        assert np.allclose(logits.numpy(), z[f"logits_c{C}"], rtol=0, atol=5e-6)
        score, pred = orc.eval_postproc(logits)
        assert np.array_equal(pred.numpy(), z[f"pred_c{C}"])
        assert np.allclose(score.numpy(), z[f"score_c{C}"], atol=1e-6)
        ln, _, _, y1n = orc.head(x, z["long_feature"], sd, use_timeconv=False)
        assert np.allclose(ln.numpy(), z[f"logits_nlonly_c{C}"], rtol=0, atol=5e-6)
        if C == 7:
            assert np.allclose(y1n.numpy(), z["y1_nlonly"], rtol=0, atol=5e-6)


def test_fp64_oracle_close_to_fp32(golden_dir):
    z = np.load(os.path.join(golden_dir, "head_b4_l30.npz"))
    seed, seq, L, B = (int(v) for v in z["meta"])
    feats = synth.features(int(z["lengths"].sum()), seed=seed)
    x = np.stack([feats[s:s + seq] for s in z["pick"]])
    sd = synth.head_state_dict(num_class=7, seed=seed)
    l64 = orc.head(x, z["long_feature"], sd, dtype=torch.float64)[0]
    assert np.abs(l64.numpy() - z["logits_c7"]).max() < 1e-5


def test_export_phase_lines():
    lines = orc.export_phase_lines(list(range(1, 8)), [5, 6], 3)
    assert lines[0] == ["0\t0", "25\t0", "50\t1", "75\t2", "100\t3"]
    assert lines[1][:3] == ["0\t0", "25\t0", "50\t4"] and len(lines[1]) == 6
