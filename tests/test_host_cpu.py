"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the host
index logic matches the reference-generated KATs, the module surface mirrors the reference's
state-dict keys, and the product refuses CPU tensors (no fallback)."""
import os

import numpy as np
import pytest
import torch

import tmr_oracle as orc
import tmrnet_b200 as tb
from tmrnet_b200 import _lib, ops, synth


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.header_symbols()
    assert len(declared) >= 20
    assert declared == sorted(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tmr_version() >= 100


def test_no_oracle_import_in_product():
    root = os.path.dirname(os.path.abspath(tb.__file__))
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "tmr_oracle" not in txt and "import oracle" not in txt, f


def _kat(golden_dir):
    z = np.load(os.path.join(golden_dir, "gather_kat.npz"))
    for i, (seq, L) in enumerate(z["meta"]):
        yield int(seq), int(L), z[f"c{i}_lengths"].tolist(), z[f"c{i}_starts"], z[f"c{i}_rows"].astype(np.int64)


def test_build_frame2row_matches_reference_walk(golden_dir):
    for seq, L, lengths, starts, rows in _kat(golden_dir):
        f2r, f2v, n_rows = ops.build_frame2row(lengths, seq)
        assert n_rows == len(starts)
        assert np.array_equal(f2r.astype(np.int64), orc.frame2row_closed_form(lengths, seq))
        got = orc.window_rows_closed_form(starts, f2r.astype(np.int64), L)
        assert np.array_equal(got, rows)
        base = np.concatenate([[0], np.cumsum(lengths)[:-1]])
        assert np.array_equal(f2v, np.repeat(base, lengths))


def test_index_from_reference_dict(golden_dir):
    for seq, L, lengths, starts, rows in _kat(golden_dir):
        a = tb.LFBIndex.from_lengths(lengths, seq)
        assert tb.get_useful_start_idx(seq, lengths) == starts.tolist()
        assert dict(a) == orc.build_start_dict(starts.tolist())
        b = tb.LFBIndex.from_dict(dict(a))
        n = len(b.frame2row_host)
        assert np.array_equal(a.frame2row_host[:n], b.frame2row_host)
        with pytest.raises(KeyError):
            a.check_starts([int(sum(lengths))])


def test_invalid_start_raises_keyerror():
    idx = tb.LFBIndex.from_lengths([7, 6, 5], 4)
    idx.check_starts([0, 3, 7, 14])
    for bad in (4, 6, 15, -1, 99):
        with pytest.raises(KeyError):
            idx.check_starts([0, bad])


def test_state_dict_keys_match_reference_shapes():
    m = tb.resnet_lstm(num_class=7)
    sd = m.state_dict()
    want = synth.head_state_dict(num_class=7)
    assert sorted(sd) == sorted(want)
    for k, v in want.items():
        assert tuple(sd[k].shape) == v.shape, k
    m.load_reference_state_dict({**{k: torch.from_numpy(v) for k, v in want.items()},
                                 "share.conv1.weight": torch.zeros(1)})
    assert torch.equal(m.nl_block.layer_norm.weight, torch.from_numpy(want["nl_block.layer_norm.weight"]))
    nlonly = tb.resnet_lstm(num_class=8, use_timeconv=False)
    assert not any(k.startswith("time_conv") for k in nlonly.state_dict())
    assert nlonly.fc_c.weight.shape == (8, 512)


def test_cpu_tensors_are_refused():
    m = tb.resnet_lstm().eval()
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="CUDA-only"):
            m(torch.zeros(1, 10, 2048), torch.zeros(1, 30, 512))
        with pytest.raises(RuntimeError, match="CUDA-only"):
            tb.TimeConv()(torch.zeros(1, 30, 512))


def test_bank_pickle_roundtrip(tmp_path):
    bank = synth.bank(17, seed=3)
    p = tmp_path / "g_LFB_test.pkl"
    tb.save_bank(torch.from_numpy(bank), str(p))
    import pickle
    raw = pickle.load(open(p, "rb"))
    assert raw.dtype == np.float64 and raw.shape == (17, 512)
    assert np.array_equal(raw.astype(np.float32), bank)


def test_export_matches_reference_format(tmp_path):
    from tmrnet_b200.export import export_phase_files, phase_lines
    lengths, seq = [12, 15], 10
    preds = torch.arange(1, 10) % 7                     # 3 + 6 clips
    lines = phase_lines(preds, lengths, seq)
    assert lines == orc.export_phase_lines(preds.tolist(), lengths, seq)
    assert lines[0][:10] == [f"{25 * k}\t0" for k in range(9)] + ["225\t1"]
    paths = export_phase_files(preds, lengths, tmp_path, seq)
    assert [os.path.basename(p) for p in paths] == ["video41-phase.txt", "video42-phase.txt"]
    assert open(paths[1]).read().splitlines() == lines[1]
    with pytest.raises(ValueError, match="number error"):
        phase_lines(preds[:-1], lengths, seq)


def test_relaxed_boundary_evaluator_hand_cases():
    """Port of matlab-eval/Evaluate.m: hand-computed cases incl. the relaxed boundary window."""
    from tmrnet_b200.evaluate import evaluate, evaluate_videos
    gt = np.array([0] * 30 + [1] * 40 + [2] * 30)
    # perfect prediction
    j, p, r, a = evaluate(gt, gt.copy())
    assert np.allclose(j[:3], 100) and np.isnan(j[3:]).all() and a == 100.0
    # late transition 0->1 by 5 frames (pred still 0 inside the first 10 s of phase 1): forgiven
    pred = gt.copy(); pred[30:35] = 0
    j, p, r, a = evaluate(gt, pred)
    assert a == 100.0 and np.allclose(j[:3], 100)
    assert abs(r[0] - 35 / 30 * 100) < 1e-9      # tp counts over the union, like the MATLAB code (Main.m caps at 100)
    # an EARLY transition is not forgiven: the reference's length-t mask lands on the head of the segment
    pred = gt.copy(); pred[25:30] = 1
    assert abs(evaluate(gt, pred)[3] - 95.0) < 1e-9
    # the same error 15 frames deep is outside the 10 s window: 5 frames counted wrong
    pred = gt.copy(); pred[30:45] = 0
    j, p, r, a = evaluate(gt, pred)
    assert abs(a - 95.0) < 1e-9
    assert abs(r[1] - (40 - 5) / 40 * 100) < 1e-9
    # a wrong phase that is not an adjacent transition is never forgiven
    pred = gt.copy(); pred[50:60] = 0
    j, p, r, a = evaluate(gt, pred)
    assert abs(a - 90.0) < 1e-9
    out = evaluate_videos([gt, gt], [gt.copy(), pred])
    assert abs(out["mean_accuracy"] - 95.0) < 1e-9 and out["jaccard_per_phase"].shape == (7,)
    with pytest.raises(ValueError):
        evaluate(gt, gt[:-1])


# ------------------------------------------------------------------------------------------
# a10 host control flow: LR schedules, best-model selection, per-phase metrics (TRAIN:806-809, 983-984, 1023-1052)
# ------------------------------------------------------------------------------------------
class _FakeTrainer:
    def __init__(self, lr):
        self.lr = lr
        self.lstm_lr = lr * 0.1


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_plateau_lr_matches_torch_reduce_on_plateau(seed):
    """Both reference parameter groups (lr and lr/10, TRAIN:797-805) against torch, through enough reductions
    that the eps = 1e-8 rule stops the LSTM group one decade before the other group."""
    from tmrnet_b200.train import PlateauLR
    rng = np.random.default_rng(seed)
    # a loss curve that improves, stalls for long stretches, and improves again
    losses = np.concatenate([np.linspace(3, 1, 8), 1.0 + 1e-5 * rng.random(25), np.linspace(0.99, 0.7, 5),
                             0.7 + 0.05 * rng.random(30), [0.1], 0.1 + rng.random(70)])
    p, q = torch.nn.Parameter(torch.zeros(1)), torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([{"params": [p]}, {"params": [q], "lr": 5e-5}], lr=5e-4)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, "min")
    tr = _FakeTrainer(5e-4)
    mine = PlateauLR(tr)
    for v in losses:
        ref.step(float(v))
        mine.step(float(v))
        assert tr.lr == opt.param_groups[0]["lr"]
        assert tr.lstm_lr == opt.param_groups[1]["lr"]
    assert tr.lr < 5e-4                      # the curve did trigger reductions
    assert tr.lstm_lr > tr.lr * 0.1 * 0.999 and tr.lr <= 5e-8   # ... down to where the eps rule holds a group back


def test_step_lr_matches_torch():
    from tmrnet_b200.train import StepLR
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1e-3)
    ref = torch.optim.lr_scheduler.StepLR(opt, step_size=3, gamma=0.1)
    tr = _FakeTrainer(1e-3)
    mine = StepLR(tr, 3, 0.1)
    for _ in range(10):
        opt.step(); ref.step(); mine.step()
        assert abs(tr.lr - opt.param_groups[0]["lr"]) <= 1e-12 * 1e-3 + 1e-18


def test_best_model_tracker_tie_rule_and_checkpoint_name():
    from tmrnet_b200.train import BestModelTracker
    t = BestModelTracker()
    sd = {"w": torch.zeros(2)}
    assert t.update(0, 0.50, 0.60, lambda: sd) and t.best_epoch == 0
    assert not t.update(1, 0.90, 0.55, lambda: sd)                 # lower val acc: ignored whatever the train acc
    assert t.update(2, 0.70, 0.60, lambda: sd) and t.best_epoch == 2          # exact tie, higher train acc: taken
    assert not t.update(3, 0.65, 0.60, lambda: sd) and t.best_epoch == 2      # exact tie, lower train acc: kept
    assert t.update(4, 0.10, 0.61, lambda: sd) and t.best_epoch == 4 and t.correspond_train_acc == 0.10
    t.best_val_acc, t.correspond_train_acc, t.best_epoch = 0.87654, 0.91239, 7
    assert t.checkpoint_name(seq=10, train_bs=400) == "lstm_epoch_7_length_10_opt_0_mulopt_1_flip_1_crop_1_batch_400_train_9124_val_8765"


def test_per_phase_precision_recall_matches_sklearn():
    from sklearn import metrics
    from tmrnet_b200.train import per_phase_precision_recall
    rng = np.random.default_rng(0)
    labels = rng.integers(0, 7, size=500)
    preds = np.where(rng.random(500) < 0.7, labels, rng.integers(0, 6, size=500))     # class 6 is never predicted wrongly-only
    p, r = per_phase_precision_recall(labels, preds)
    assert np.allclose(p, metrics.precision_score(labels, preds, average=None, zero_division=0))
    assert np.allclose(r, metrics.recall_score(labels, preds, average=None, zero_division=0))


# ------------------------------------------------------------------------------------------
# 8f-3: checkpoint I/O either side of the path (TRAIN:772-774, 1053; EVAL:443-447; code/models.py)
# ------------------------------------------------------------------------------------------
class _RefShapedHead(torch.nn.Module):
    """A torch module with the attribute layout of the reference's stage-2 `resnet_lstm` (TRAIN:209-230: share, lstm,
    fc_c, fc_h_c, nl_block, dropout, time_conv) and a stand-in trunk: what the reference's STRICT load
    (EVAL:443-447) would check our checkpoint against."""

    def __init__(self, C):
        super().__init__()
        nn = torch.nn
        self.share = nn.Sequential(nn.Conv2d(3, 4, 1), nn.BatchNorm2d(4))
        self.lstm = nn.LSTM(2048, 512, batch_first=True)
        self.fc_c = nn.Linear(512, C)
        self.fc_h_c = nn.Linear(1024, 512)
        self.nl_block = nn.Module()
        for i in (1, 2, 3, 4):
            setattr(self.nl_block, f"linear{i}", nn.Linear(512, 512))
        self.nl_block.layer_norm = nn.LayerNorm([1, 512])
        self.dropout = nn.Dropout(0.5)
        self.time_conv = nn.Module()
        for i, k in ((1, 3), (2, 5), (3, 7)):
            setattr(self.time_conv, f"timeconv{i}", nn.Conv1d(512, 512, k, padding=k // 2))


@pytest.mark.parametrize("C", [6, 7, 8])
def test_reference_checkpoint_round_trip_is_strict_loadable(tmp_path, C):
    import tmrnet_b200 as tb
    ref = _RefShapedHead(C)
    ref_sd = ref.state_dict()                                   # what torch.save(model.module.state_dict()) holds
    m = tb.resnet_lstm(num_class=C)
    m.load_reference_state_dict(ref_sd)                         # strict on the head's keys; share.* kept aside
    for k, v in m.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    path = str(tmp_path / "latest_model.pth")
    m.save_reference_checkpoint(path)
    back = torch.load(path)
    assert list(back.keys()) == list(ref_sd.keys())             # same keys in the same order
    fresh = _RefShapedHead(C)
    fresh.load_state_dict(back, strict=True)                    # EVAL:443-447
    for k, v in fresh.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    # a head-only checkpoint (no trunk available) still loads the reference way it is used at TRAIN:774
    bare = tb.resnet_lstm(num_class=C)
    bare.load_state_dict(m.state_dict())
    res = _RefShapedHead(C).load_state_dict(bare.reference_state_dict(), strict=False)
    assert all(k.startswith("share.") for k in res.missing_keys) and not res.unexpected_keys
    # wrong class count is an error, as in the reference
    with pytest.raises(RuntimeError):
        tb.resnet_lstm(num_class=C + 1).load_reference_state_dict(ref_sd)


def test_stage1_checkpoint_feeds_the_head_and_the_stage1_surface():
    """TRAIN:772-774: the stage-2 model starts from a STAGE-1 checkpoint with strict=False - `share.*` and `lstm.*`
    match, the stage-1 classifier `fc.*` does not exist in the head.  code/models.py names the trunk `res.*`."""
    import types
    import tmrnet_b200 as tb
    g = torch.Generator().manual_seed(0)
    stage1 = {"res.0.weight": torch.randn(4, 3, 1, 1, generator=g),
              "lstm.weight_ih_l0": torch.randn(2048, 2048, generator=g), "lstm.weight_hh_l0": torch.randn(2048, 512, generator=g),
              "lstm.bias_ih_l0": torch.randn(2048, generator=g), "lstm.bias_hh_l0": torch.randn(2048, generator=g),
              "fc.weight": torch.randn(7, 512, generator=g), "fc.bias": torch.randn(7, generator=g)}
    head = tb.resnet_lstm(num_class=7)
    before = head.fc_c.weight.detach().clone()
    missing, dropped = head.load_stage1_state_dict(stage1)
    assert torch.equal(head.lstm.weight_ih_l0, stage1["lstm.weight_ih_l0"]) and torch.equal(head.fc_c.weight, before)
    assert sorted(dropped) == ["fc.bias", "fc.weight"] and "fc_c.weight" in missing and not any(k.startswith("lstm.") for k in missing)
    assert list(head.reference_state_dict().keys())[0] == "share.0.weight"      # res.* written back under the scripts' prefix
    # the stage-1 surface itself: constructor, keys, optimisers of code/models.py
    args = types.SimpleNamespace(opt=0, lr=5e-4, momentum=0.9, dampening=0, weightdecay=5e-4, nesterov=True, seq=10)
    s1 = tb.models.resnet_lstm(args, 7)
    s1.load_reference_state_dict(stage1)
    assert sorted(s1.state_dict().keys()) == sorted(k for k in stage1 if not k.startswith("res."))
    assert sorted(s1.reference_state_dict().keys()) == sorted(stage1.keys())
    opt = s1.get_optimizers()
    assert isinstance(opt, torch.optim.SGD) and [g["lr"] for g in opt.param_groups] == [5e-4, 5e-4] and opt.defaults["nesterov"]
    args.opt = 1
    assert isinstance(s1.get_optimizers(), torch.optim.Adam)
    with pytest.raises(RuntimeError):
        s1.train()(torch.zeros(1, 10, 2048))


def test_launch_accounting_of_the_bench_pass():
    """bench.py's gpu_launches claim comes from BankInference.launches_per_run(): for the 83 022-clip bench job in the
    tensor-core mode one resident pass is the 20 kernels of profiles/r2au_launches_one_pass.md (feature conversion,
    row table, projection, step-0 fix-up, persistent recurrence + its small-batch remainder, 2 + 4 bank-side kernels,
    3 relation GEMMs + attention + LayerNorm, [St || y] conversion, classifier GEMM, FC / argmax); fp16 features drop
    the conversion."""
    import tmrnet_b200 as tb
    from tmrnet_b200 import synth
    from tmrnet_b200.infer import BankInference
    lengths = synth.video_lengths(40, seed=1234)
    idx = tb.LFBIndex.from_lengths(lengths, 10)
    m = tb.resnet_lstm(num_class=7)
    m.math_mode = "f16"
    eng = BankInference(m, idx, 10, 30)
    assert len(eng.plan()) == 1 and len(eng.starts_host) == 83022
    assert eng.launches_per_run() == 20
    assert eng.launches_per_run(feats_f16=True) == 19
