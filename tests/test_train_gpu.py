"""Head training step (a10): gradients of the hand-written backward against torch autograd over the
oracle graph (fp64, dropout off), union-batch equivalence of summed per-rank gradients, and the SGD
update against torch.optim.SGD."""
import numpy as np
import pytest
import torch
import torch.nn.functional as Fn

import tmr_oracle as orc
import tmrnet_b200 as tb
from tmrnet_b200 import synth
from tmrnet_b200.train import PARAM_ORDER, HeadTrainer

pytestmark = pytest.mark.gpu


def _setup(C=8, B=12, L=30, seq=10, seed=3, use_timeconv=True):
    dev = torch.device("cuda:0")
    sd = synth.head_state_dict(num_class=C, seed=seed)
    if not use_timeconv:
        sd = {k: v for k, v in sd.items() if not k.startswith("time_conv.")}
    m = tb.resnet_lstm(num_class=C, sequence_length=seq, use_timeconv=use_timeconv)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to(dev)
    rng = np.random.default_rng(seed)
    x = synth.features(B * seq, seed=seed).reshape(B, seq, 2048)
    lf = synth.bank(B * L, seed=seed + 1).reshape(B, L, 512)
    labels = rng.integers(0, C, size=B)
    cw = (0.5 + rng.random(C)).astype(np.float32)
    return dev, sd, m, x, lf, labels, cw


def _autograd_reference(sd, x, lf, labels, cw, use_timeconv=True):
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in sd.items()}
    logits = orc.head(torch.from_numpy(x).double(), torch.from_numpy(lf).double(), _NoDetach(P), use_timeconv, torch.float64)[0]
    loss = Fn.cross_entropy(logits, torch.from_numpy(labels), weight=torch.from_numpy(cw).double(), reduction="sum")
    loss.backward()
    return loss.detach(), logits.detach(), {k: p.grad for k, p in P.items()}


class _NoDetach(dict):
    """orc._t() detaches plain tensors; hand it parameters through a mapping whose values are kept."""


def _patch_oracle(monkeypatch):
    def keep(x, dtype=torch.float32):
        if isinstance(x, torch.Tensor):
            return x.to(dtype)
        return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)
    monkeypatch.setattr(orc, "_t", keep)


@pytest.mark.parametrize("use_timeconv", [True, False])
def test_gradients_match_autograd(monkeypatch, use_timeconv):
    _patch_oracle(monkeypatch)
    dev, sd, m, x, lf, labels, cw = _setup(use_timeconv=use_timeconv)
    ref_loss, ref_logits, ref_g = _autograd_reference(sd, x, lf, labels, cw, use_timeconv)
    tr = HeadTrainer(m, class_weight=cw)
    loss, logits, pred = tr.forward_backward(torch.from_numpy(x).to(dev), torch.from_numpy(lf).to(dev),
                                             torch.from_numpy(labels).to(dev), dropout=False)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref_loss)) < 1e-4 * abs(float(ref_loss))
    assert float((logits.cpu().double() - ref_logits).abs().max()) < 1e-5
    assert torch.equal(pred.cpu(), ref_logits.argmax(1))
    for i, k in enumerate(PARAM_ORDER):
        if k not in sd:
            continue
        g = tr.grads.views[i].cpu().double()
        r = ref_g[k].reshape(g.shape)
        scale = float(r.abs().max())
        if k == "nl_block.linear2.bias":            # exactly zero analytically (softmax shift invariance)
            assert float(g.abs().max()) == 0.0 and scale < 1e-9
            continue
        err = float((g - r).abs().max())
        assert err <= 2e-4 * scale + 1e-7, (k, err, scale)


@pytest.mark.parametrize("use_timeconv", [True, False])
def test_f16_backward_gemms_on_fp32_activations(monkeypatch, use_timeconv):
    """TMR_MATH_F16 backward alone: forward in fp32 (same ReLU masks / max-branch winners as the reference graph), then
    the backward with fp16-rounded GEMM operands on the tensor cores - gradients within 2e-3 of each tensor's largest
    entry of torch autograd over the fp64 graph (measured <= 9.1e-4)."""
    import ctypes as C
    from tmrnet_b200 import _lib, ops
    from tmrnet_b200.ops import D, F, _ptr, _stream, _ws, check
    from tmrnet_b200.train import FlatBuffer, _param_list
    _patch_oracle(monkeypatch)
    dev, sd, m, x, lf, labels, cw = _setup(use_timeconv=use_timeconv)
    _, _, ref_g = _autograd_reference(sd, x, lf, labels, cw, use_timeconv)
    X, LF, Y = torch.from_numpy(x).to(dev), torch.from_numpy(lf).to(dev), torch.from_numpy(labels).to(dev)
    lib = _lib.load()
    params = _param_list(m)
    B, seq, L, Cn = X.shape[0], 10, LF.shape[1], m.num_class
    ws = _ws(lib.tmr_head_train_workspace_bytes(B, seq, L, D, F, Cn), dev)
    logits = torch.empty((B, Cn), dtype=torch.float32, device=dev)
    pp = (C.c_void_p * 24)(*[p.data_ptr() if p is not None else 0 for p in params])
    check(lib.tmr_head_train_fwd(pp, _ptr(X), _ptr(LF), B, seq, L, F, D, Cn, 0.0, 0.0, 0, _ptr(logits), _ptr(ws), ws.numel(),
                                 ops.TMR_MATH_FP32, _stream()))
    lg = logits.detach().clone().requires_grad_(True)
    Fn.cross_entropy(lg, Y, weight=torch.from_numpy(cw).to(dev), reduction="sum").backward()
    grads = FlatBuffer([tuple(p.shape) if p is not None else None for p in params], dev)
    gp = (C.c_void_p * 24)(*[g.data_ptr() if g is not None else 0 for g in grads.views])
    scratch = torch.empty((B, Cn), dtype=torch.float32, device=dev)
    check(lib.tmr_head_train_bwd(pp, gp, _ptr(X), _ptr(LF), _ptr(lg.grad.contiguous()), B, seq, L, F, D, Cn, _ptr(scratch),
                                 _ptr(ws), ws.numel(), ops.TMR_MATH_F16, _stream()))
    torch.cuda.synchronize()
    for i, k in enumerate(PARAM_ORDER):
        if k not in sd or k == "nl_block.linear2.bias":
            continue
        g = grads.views[i].cpu().double()
        r = ref_g[k].reshape(g.shape)
        assert float((g - r).abs().max()) <= 2e-3 * float(r.abs().max()) + 1e-7, k


@pytest.mark.parametrize("use_timeconv", [True, False])
def test_f16_training_step_against_the_fp64_graph(monkeypatch, use_timeconv):
    """The whole TMR_MATH_F16 step: logits within 3e-3, loss within 1e-3; the gradients point the same way as the fp64
    graph's (cosine >= 0.998, relative L2 error <= 8e-2 per tensor).  They are not compared entry by entry: a forward
    that is 1e-3 off flips the ReLU / max-branch decision of the few activations that close to zero, and each flip
    moves individual gradient entries by O(1) of their size (test_f16_backward_gemms_on_fp32_activations holds the
    backward itself to 2e-3 on identical activations)."""
    _patch_oracle(monkeypatch)
    dev, sd, m, x, lf, labels, cw = _setup(use_timeconv=use_timeconv)
    ref_loss, ref_logits, ref_g = _autograd_reference(sd, x, lf, labels, cw, use_timeconv)
    tr = HeadTrainer(m, class_weight=cw, math_mode="f16")
    loss, logits, pred = tr.forward_backward(torch.from_numpy(x).to(dev), torch.from_numpy(lf).to(dev),
                                             torch.from_numpy(labels).to(dev), dropout=False)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref_loss)) < 1e-3 * abs(float(ref_loss))
    assert float((logits.cpu().double() - ref_logits).abs().max()) < 3e-3
    for i, k in enumerate(PARAM_ORDER):
        if k not in sd or k == "nl_block.linear2.bias":
            continue
        g = tr.grads.views[i].cpu().double().ravel()
        r = ref_g[k].ravel()
        assert float(torch.dot(g, r) / (g.norm() * r.norm())) >= 0.998, k
        assert float((g - r).norm() / r.norm()) <= 8e-2, k


def test_summed_rank_gradients_equal_union_batch():
    """Loss is sum-reduced: grads(A) + grads(B) == grads(A u B) -> all-reduce SUM reproduces 1-GPU training."""
    dev, sd, m, x, lf, labels, cw = _setup(B=16)
    tr = HeadTrainer(m, class_weight=cw)
    X, LF, Y = (torch.from_numpy(a).to(dev) for a in (x, lf, labels))
    tr.forward_backward(X, LF, Y, dropout=False)
    full = tr.grads.flat.clone()
    tr.forward_backward(X[:7], LF[:7], Y[:7], dropout=False)
    part = tr.grads.flat.clone()
    tr.forward_backward(X[7:], LF[7:], Y[7:], dropout=False)
    part += tr.grads.flat
    assert float((part - full).abs().max()) <= 2e-5 * float(full.abs().max())


def test_sgd_update_matches_torch_and_invalidates_packs():
    dev, sd, m, x, lf, labels, cw = _setup(B=8, C=7)
    X, LF, Y = (torch.from_numpy(a).to(dev) for a in (x, lf, labels))
    tr = HeadTrainer(m, lr=1e-3, momentum=0.9, weight_decay=5e-4, lstm_lr_scale=0.1, class_weight=cw)
    named = dict(m.named_parameters())
    ref = {k: named[k].detach().clone() for k in PARAM_ORDER}
    opt = torch.optim.SGD([{"params": [ref[k] for k in PARAM_ORDER if k.startswith("lstm.")], "lr": 1e-4},
                           {"params": [ref[k] for k in PARAM_ORDER if not k.startswith("lstm.")]}],
                          lr=1e-3, momentum=0.9, weight_decay=5e-4, dampening=0)
    with torch.no_grad():
        before = m.eval()(X, LF).clone()
    for step in range(3):
        tr.forward_backward(X, LF, Y, dropout=False)
        for i, k in enumerate(PARAM_ORDER):
            ref[k].grad = tr.grads.views[i].detach().clone().reshape(ref[k].shape)
        opt.step()
        tr.sgd_update()
        for k in PARAM_ORDER:
            assert torch.allclose(named[k].detach(), ref[k], rtol=1e-6, atol=1e-8), (step, k)
    with torch.no_grad():
        after = m.eval()(X, LF)
    assert not torch.equal(before, after)          # inference caches were refreshed after the update


def test_dropout_masks_are_seeded_and_scaled():
    dev, sd, m, x, lf, labels, cw = _setup(B=8, C=7)
    X, LF, Y = (torch.from_numpy(a).to(dev) for a in (x, lf, labels))
    a = HeadTrainer(m, class_weight=cw, seed=5)
    l1, _, _ = a.forward_backward(X, LF, Y)
    g1 = a.grads.flat.clone()
    l2, _, _ = a.forward_backward(X, LF, Y)
    assert torch.equal(g1, a.grads.flat) and torch.equal(l1, l2)       # same (seed, step) -> same masks
    b = HeadTrainer(m, class_weight=cw, seed=6)
    l3, _, _ = b.forward_backward(X, LF, Y)
    assert not torch.equal(g1, b.grads.flat)
    l0, _, _ = a.forward_backward(X, LF, Y, dropout=False)
    assert float(l0) != float(l1)


def test_fit_epoch_loop_learns_and_tracks_best():
    """The reference's epoch loop (TRAIN:815-1056) around HeadTrainer.step on a tiny synthetic job whose labels are a
    function of the features: training loss falls, validation loss equals the oracle's sum-reduced CE on the same
    weights, the plateau scheduler and the best-model tracker are driven per epoch."""
    from tmrnet_b200.train import PlateauLR, fit
    dev = torch.device("cuda:0")
    C, seq, L = 7, 10, 30
    lengths = [60, 45, 70]
    index = tb.LFBIndex.from_lengths(lengths, seq)
    starts = np.array(tb.get_useful_start_idx(seq, lengths))
    feats = synth.features(sum(lengths), seed=9)
    bank = synth.bank(len(starts), seed=9)
    labels = (np.argmax(feats[:, :C], axis=1)).astype(np.int64)          # learnable from the last frame's features
    sd = synth.head_state_dict(num_class=C, seed=4)
    m = tb.resnet_lstm(num_class=C, sequence_length=seq)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to(dev)
    tr = HeadTrainer(m, lr=1e-4, p_nl=0.0, p_fc=0.0, seed=1)      # the loss is SUM-reduced over 40 clips (TRAIN:780)
    sched = PlateauLR(tr, patience=0)
    fd, bd = torch.from_numpy(feats).to(dev), torch.from_numpy(bank).to(dev)
    rng = np.random.default_rng(0)
    perm = rng.permutation(starts)
    tr_s, va_s = np.sort(perm[:120]), np.sort(perm[120:150])
    seen = []
    hist, tracker = fit(tr, index, fd, bd, labels, tr_s, va_s, epochs=4, batch_clips=40, L=L, seed=3, scheduler=sched, log=seen.append)
    assert len(hist) == 4 and len(seen) == 4
    assert min(h["train_loss"] for h in hist[1:]) < hist[0]["train_loss"]
    assert all(np.isfinite(h["train_loss"]) for h in hist)
    assert tracker.best_state is not None and 0 <= tracker.best_epoch < 4
    assert all(0.0 <= h["val_acc"] <= 1.0 and np.isfinite(h["val_loss"]) for h in hist)
    assert any(h["best"] for h in hist)
    # validation loss of the final weights against the oracle head on the same clips
    sd_now = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    x = np.stack([feats[s:s + seq] for s in va_s])
    lf = orc.get_long_feature(va_s, orc.build_start_dict(starts.tolist()), bank, L)
    logits = orc.head(x, lf, sd_now, dtype=torch.float64)[0]
    y = torch.from_numpy(labels[va_s + seq - 1])
    ref_loss = float(Fn.cross_entropy(logits, y, reduction="sum")) / len(va_s)
    m.eval()
    hist2, _ = fit(tr, index, fd, bd, labels, tr_s[:0], va_s, epochs=1, batch_clips=40, L=L)
    assert abs(hist2[0]["val_loss"] - ref_loss) < 2e-3 * max(1.0, abs(ref_loss))


# ---------------------------------------------------------------------------------------------
# the boundary: training THROUGH the module with stock torch losses / optimisers (TRAIN:786-805, 876-887)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("use_timeconv", [True, False])
def test_module_autograd_gradients_match_reference_graph(monkeypatch, use_timeconv):
    """model(x, long_feature) records ONE autograd node; loss.backward() fills .grad of every head parameter with the
    gradient torch autograd gives over the fp64 oracle graph (dropout off: eval mode with autograd on)."""
    _patch_oracle(monkeypatch)
    dev, sd, m, x, lf, labels, cw = _setup(use_timeconv=use_timeconv)
    ref_loss, ref_logits, ref_g = _autograd_reference(sd, x, lf, labels, cw, use_timeconv)
    X, LF, Y = (torch.from_numpy(a).to(dev) for a in (x, lf, labels))
    m.eval()
    criterion = torch.nn.CrossEntropyLoss(reduction="sum", weight=torch.from_numpy(cw).to(dev))     # TRAIN:780
    outputs = m.forward(X, LF)
    assert outputs.requires_grad
    loss = criterion(outputs, Y)
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) < 1e-4 * abs(float(ref_loss))
    for k, prm in m.named_parameters():
        g, r = prm.grad.cpu().double(), ref_g[k].reshape(prm.shape)
        scale = float(r.abs().max())
        if k == "nl_block.linear2.bias":
            assert float(g.abs().max()) == 0.0
            continue
        assert float((g - r).abs().max()) <= 2e-4 * scale + 1e-7, k
    # a second backward through a fresh forward ACCUMULATES, as autograd does
    criterion(m.forward(X, LF), Y).backward()
    k, prm = next(iter(m.named_parameters()))
    assert torch.allclose(prm.grad.cpu().double(), 2 * ref_g[k].reshape(prm.shape), rtol=1e-3, atol=1e-7)
    # frozen features: inputs that require grad are refused, not silently ignored
    with pytest.raises(RuntimeError):
        m.forward(X.clone().requires_grad_(), LF)


def test_reference_loop_body_with_torch_sgd_equals_head_trainer():
    """Three iterations of the reference's loop body (TRAIN:876-887) on the module with torch.optim.SGD built like
    TRAIN:797-805 (LSTM group at lr/10, momentum 0.9, weight decay 5e-4) == three HeadTrainer.step() calls."""
    dev, sd, m, x, lf, labels, cw = _setup(B=10, C=7)
    X, LF, Y = (torch.from_numpy(a).to(dev) for a in (x, lf, labels))
    twin = tb.resnet_lstm(num_class=7, sequence_length=10)
    twin.load_state_dict(m.state_dict())
    twin = twin.to(dev)
    for mod in (m, twin):
        mod.dropout.p = 0.0
        mod.nl_block.dropout.p = 0.0
    lr = 1e-3
    optimizer = torch.optim.SGD([
        {"params": m.lstm.parameters()},
        {"params": m.time_conv.parameters(), "lr": lr},
        {"params": m.nl_block.parameters(), "lr": lr},
        {"params": m.fc_h_c.parameters(), "lr": lr},
        {"params": m.fc_c.parameters(), "lr": lr},
    ], lr=lr / 10, momentum=0.9, dampening=0, weight_decay=5e-4, nesterov=False)
    criterion_phase = torch.nn.CrossEntropyLoss(reduction="sum", weight=torch.from_numpy(cw).to(dev))
    tr = HeadTrainer(twin, lr=lr, momentum=0.9, weight_decay=5e-4, lstm_lr_scale=0.1, class_weight=cw, p_nl=0.0, p_fc=0.0)
    m.train()
    for step in range(3):
        optimizer.zero_grad()
        outputs_phase = m.forward(X, LF)
        _, preds_phase = torch.max(outputs_phase.data, 1)
        loss_phase = criterion_phase(outputs_phase, Y)
        loss_phase.backward()
        optimizer.step()
        loss2, _, pred2 = tr.step(X, LF, Y)
        assert abs(float(loss_phase.detach()) - float(loss2)) <= 1e-5 * abs(float(loss2))
        assert torch.equal(preds_phase, pred2)
        for (k, a), (_, b) in zip(m.named_parameters(), twin.named_parameters()):
            assert torch.allclose(a, b, rtol=2e-6, atol=1e-8), (step, k)
    # the inference path sees the updated weights (packed-weight caches follow the parameters' version counters)
    m.eval()
    twin.eval()
    m.math_mode = twin.math_mode = "fp32"          # (weights equal to a few ulp: compare in the exact-order mode)
    with torch.no_grad():
        a, b = m(X, LF), twin(X, LF)
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-4)
    ref_logits = orc.head(x, lf, {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()})[0]
    assert float((a.cpu() - ref_logits).abs().max()) < 2e-5 * float(ref_logits.abs().max())


@pytest.mark.parametrize("opt", ["adam", "nesterov"])
def test_adam_and_nesterov_branches_train(opt):
    """optimizer_choice == 1 (Adam, TRAIN:800-805) and use_nesterov (TRAIN:786-799) work through stock torch
    optimisers on the module's parameters; the loss on a fixed batch falls."""
    dev, sd, m, x, lf, labels, cw = _setup(B=16, C=7)
    X, LF, Y = (torch.from_numpy(a).to(dev) for a in (x, lf, labels))
    if opt == "adam":
        optimizer = torch.optim.Adam([{"params": m.lstm.parameters()}, {"params": [p for k, p in m.named_parameters()
                                                                                 if not k.startswith("lstm.")], "lr": 1e-4}], lr=1e-5)
    else:
        optimizer = torch.optim.SGD(m.parameters(), lr=1e-4, momentum=0.9, dampening=0, weight_decay=5e-4, nesterov=True)
    criterion = torch.nn.CrossEntropyLoss(reduction="sum")
    m.train()
    m.dropout.p = 0.0
    m.nl_block.dropout.p = 0.0
    losses = []
    for _ in range(6):
        optimizer.zero_grad()
        loss = criterion(m(X, LF), Y)
        loss.backward()
        optimizer.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_module_training_mode_dropout():
    dev, sd, m, x, lf, labels, cw = _setup(B=8, C=7)
    X, LF = torch.from_numpy(x).to(dev), torch.from_numpy(lf).to(dev)
    m.train()
    torch.manual_seed(11)
    m._train_calls = 0
    a = m(X, LF).detach().clone()
    b = m(X, LF).detach().clone()
    assert not torch.equal(a, b)                       # fresh masks per call
    torch.manual_seed(11)
    m._train_calls = 0
    assert torch.equal(m(X, LF).detach(), a)           # reproducible under torch.manual_seed
    m.eval()
    with torch.no_grad():
        e1, e2 = m(X, LF), m(X, LF)
    assert torch.equal(e1, e2)
    g = m(X, LF)                                       # eval + autograd: dropout off, differentiable, fp32 forward
    assert g.requires_grad and float((g.detach() - e1).abs().max()) < 2e-3 * float(e1.abs().max())
