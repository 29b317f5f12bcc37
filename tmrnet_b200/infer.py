"""Bank-level inference of the TMRNet head: the reference's test loop body
(code/eval/python/test_singlenet_phase_non-local_pretrained_2fc_copy_mutiConv6_resnest.py:470-499)
over per-frame backbone features and a memory bank that are both resident in HBM, plus the
video-sharding plan for multi-GPU inference: videos are independent units, each GPU
holds its own videos' features and bank rows (+ a halo of the previous video's tail, because the
reference window leaks across the video boundary), and no collective is needed.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, ops
from .lfb import LFBIndex
from .ops import D, F, _dev, _mode, _ptr, _stream, _ws, check


# ~40 KB of workspace per clip (projected rows, fp16 features, TimeConv tap products): 5 GB at the cap.  The
# 83 k clips of a Cholec80-sized job run as ONE batch: 4.78 ms per pass against 4.95 ms as two batches and 5.12 ms
# as three (every kernel of a batch pays a pipeline ramp and a partial last round).
MAX_BATCH_CLIPS = 131072


def _sm_count() -> int:
    if torch.cuda.is_available():
        return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return 148


def default_host_batch_clips() -> int:
    """Host-streamed pass: 64 clips x SM count (9472 on a 148-SM B200) - the persistent GEMMs work on
    CTA pairs x N tiles, so this is a whole number of rounds over the 74 CTA pairs, and each batch's
    kernels stay shorter than its H2D copy (scripts/e2e_timeline.py)."""
    return 64 * _sm_count()


class BankInference:
    """Runs the head over every clip of a (shard of a) feature bank in clip batches.

    model: tmrnet_b200.resnet_lstm (eval).  feats: CUDA (n_frames, 2048).  bank: CUDA (n_rows, 512).
    index: LFBIndex.from_lengths(video_lengths, seq) for the same frames/rows.
    """

    def __init__(self, model, index: LFBIndex, seq: int = 10, L: int = 30, batch_clips: int = None,
                 pad_mode: str = "repeat", math_mode=None, starts=None, dedup: bool = True, tail_clips: int = 0,
                 host_batch_clips: int = None, irr_from_rows: bool = True):
        self.model = model
        self.index = index
        self.seq, self.L = int(seq), int(L)
        # resident pass: as few, equally sized batches as MAX_BATCH_CLIPS allows - every kernel of a batch
        # pays a pipeline ramp (first main loop, last epilogue), measured on the bench workload (83 k
        # clips): 9.04 ms/step at 18944 clips per batch, 8.67 at 27776, 8.48 at 41600
        self.batch_clips = int(batch_clips) if batch_clips else None
        # run_host streams features H2D ahead of compute and is PCIe-bound (copies 12.5 ms, kernels
        # 10 ms per pass of the bench workload): what is left to tune is the compute still running after
        # the last copy, i.e. the size of the last batches.  Half-size batches (64 clips x SM count)
        # keep each batch's kernels shorter than its copy and end the pass ~one small batch after the
        # last copy (scripts/e2e_timeline.py).  tail_clips > 0 additionally splits a short final batch.
        self.host_batch_clips = int(host_batch_clips) if host_batch_clips else (
            max(128, self.batch_clips // 2) if self.batch_clips else default_host_batch_clips())
        self.tail_clips = int(tail_clips)
        self._host_eng = None
        self._graph = None          # (key, CUDAGraph, ...) of the resident pass, see run()
        self._last_key = None
        # irregular clips (first L of every video): TimeConv assembled from per-row tap products of the rows
        # their windows touch (True) or per-clip gather + TimeConv (False)
        self.irr_from_rows = bool(irr_from_rows)
        self._ctor = dict(pad_mode=pad_mode, math_mode=math_mode, starts=starts, dedup=dedup, tail_clips=tail_clips,
                          irr_from_rows=irr_from_rows)
        self.pad_mode = {"repeat": ops.TMR_PAD_REPEAT, "zero": ops.TMR_PAD_ZERO}[pad_mode]
        self.math_mode = math_mode
        if starts is None:       # every clip of the index; a VideoShard passes its owned clips only
            starts = np.fromiter(index.keys(), dtype=np.int64, count=len(index))
        else:                    # caller-chosen clips: each must be a valid clip start of the index (KeyError otherwise,
            index.check_starts(np.asarray(starts, dtype=np.int64))      # like the reference's dict probe)
        self.starts_host = np.sort(np.asarray(starts, dtype=np.int64))
        self._ws = None
        self._starts_dev = None
        # bank-level TimeConv dedup (tmr_head_frames_dedup_fwd): tensor-core mode, TimeConv present, L >= 6
        self.dedup = bool(dedup) and self.L >= 6 and model.time_conv is not None
        self._dedup_plan = None
        self._dedup_dev = {}

    def _use_dedup(self):
        mode = _mode(self.math_mode if self.math_mode is not None else self.model.math_mode)
        return self.dedup and mode == ops.TMR_MATH_F16

    def dedup_plan(self):
        """Per batch: which clips are regular (window = contiguous run of bank rows), the PB row range
        and the per-clip source index the attention kernel reads (see include/tmr_b200.h)."""
        if self._dedup_plan is None:
            f2r = self.index.frame2row_host.astype(np.int64)
            f2v = self.index.frame2vstart_host
            if f2v is None:
                raise ValueError("dedup needs an LFBIndex built with from_lengths() (video boundaries)")
            f2v = f2v.astype(np.int64)
            out = []
            for lo, hi, fl, fh in self.plan():
                s = self.starts_host[lo:hi]
                regular = (s - f2v[s]) >= self.L
                r0 = f2r[s] - 1                                   # bank row of slot 0
                src = np.empty(hi - lo, dtype=np.int32)
                irr = s[~regular]
                src[~regular] = -1 - np.arange(len(irr), dtype=np.int32)
                if regular.any():
                    row_base = int(r0[regular].min()) - (self.L - 1)
                    pb_rows = int(r0[regular].max()) - row_base + 1
                    src[regular] = (r0[regular] - row_base).astype(np.int32)
                else:
                    row_base, pb_rows = 0, 0
                # distinct bank rows the irregular windows touch (same rule as gather_kernel): their TimeConv
                # is assembled from per-row tap products of exactly these rows
                if len(irr) and self.irr_from_rows:
                    key = irr[:, None] - np.arange(1, self.L + 1, dtype=np.int64)[None, :]
                    rows = f2r[np.maximum(key, 0)]
                    if self.pad_mode == ops.TMR_PAD_ZERO:
                        rows = rows[key >= f2v[irr][:, None]]
                    else:
                        rows = np.where(key >= 0, rows, 0)
                    irr_rows = np.unique(rows).astype(np.int32)
                else:
                    irr_rows = np.zeros(0, dtype=np.int32)
                out.append(dict(src=src, irr=irr.astype(np.int64), irr_rows=irr_rows, row_base=row_base, pb_rows=pb_rows))
            self._dedup_plan = out
        return self._dedup_plan

    def _dedup_tensors(self, dev):
        if dev not in self._dedup_dev:
            self._dedup_dev[dev] = [(torch.from_numpy(d["src"]).to(dev),
                                     torch.from_numpy(d["irr"]).to(dev) if len(d["irr"]) else None,
                                     torch.from_numpy(d["irr_rows"]).to(dev) if len(d["irr_rows"]) else None)
                                    for d in self.dedup_plan()]
        return self._dedup_dev[dev]

    def plan(self):
        """[(clip_lo, clip_hi, frame_lo, frame_hi)] per batch; frames cover every clip of the batch."""
        out = []
        n = len(self.starts_host)
        per = self.batch_clips
        if not per:                                # even split, rounded up to whole 128-clip M tiles
            nb = max(1, -(-n // MAX_BATCH_CLIPS))
            per = max(128, -(-(-(-n // nb)) // 128) * 128)
        bounds = list(range(0, n, per)) + [n]
        if self.tail_clips > 0 and len(bounds) > 2 and bounds[-1] - bounds[-2] > 2 * self.tail_clips:
            bounds.insert(-1, n - self.tail_clips)
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            out.append((lo, hi, int(self.starts_host[lo]), int(self.starts_host[hi - 1]) + self.seq))
        return out

    def _alloc_out(self, dev, want_st):
        n, Cn = len(self.starts_host), self.model.num_class
        out = dict(logits=torch.empty((n, Cn), dtype=torch.float32, device=dev),
                   pred=torch.empty((n,), dtype=torch.int64, device=dev),
                   score=torch.empty((n,), dtype=torch.float32, device=dev))
        if want_st:
            out["St"] = torch.empty((n, D), dtype=torch.float32, device=dev)
        return out

    def _prepare(self, dev, bank):
        """Everything a pass needs that does not depend on the feature buffer."""
        lib = _lib.load()
        if self._starts_dev is None or self._starts_dev.device != dev:
            self._starts_dev = torch.from_numpy(self.starts_host).to(dev)
        plan = self.plan()
        dedup = self._use_dedup()
        dplan = dten = None
        if dedup:
            dplan, dten = self.dedup_plan(), self._dedup_tensors(dev)
            need = max((lib.tmr_head_frames_dedup_workspace_bytes(fh - fl, hi - lo, len(d["irr"]), len(d["irr_rows"]), d["pb_rows"], self.L, D)
                        for (lo, hi, fl, fh), d in zip(plan, dplan)), default=256)
        else:
            need = max((lib.tmr_head_frames_workspace_bytes(fh - fl, hi - lo, self.L, D) for lo, hi, fl, fh in plan),
                       default=256)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = _ws(need, dev)
        f2r, f2v = self.index.device_tables(dev)
        mode = _mode(self.math_mode if self.math_mode is not None else self.model.math_mode)
        return dict(lib=lib, plan=plan, dedup=dedup, dplan=dplan, dten=dten, ws=self._ws, f2r=f2r, f2v=f2v, mode=mode,
                    packs=self.model.packs(), bank=bank, starts=self._starts_dev)

    def _launch_batch(self, ctx, i, feats_ptr, out, stream, feats_f16=False):
        """Enqueue batch i; feats_ptr addresses the features of frame plan[i].frame_lo."""
        lo, hi, fl, fh = ctx["plan"][i]
        lib, packs, bank, ws = ctx["lib"], ctx["packs"], ctx["bank"], ctx["ws"]
        Cn = self.model.num_class
        st = out.get("St")
        head_in = (_ptr(packs[0]), _ptr(packs[1]), _ptr(packs[2]), _ptr(packs[3]), C.c_void_p(feats_ptr))
        common_in = (fh - fl, fl,
                     _ptr(bank), bank.shape[0], _ptr(ctx["f2r"]), _ptr(ctx["f2v"]), ctx["f2r"].numel(),
                     C.c_void_p(ctx["starts"].data_ptr() + lo * 8), hi - lo)
        common_out = (C.c_void_p(out["logits"].data_ptr() + lo * Cn * 4),
                      C.c_void_p(out["pred"].data_ptr() + lo * 8),
                      C.c_void_p(out["score"].data_ptr() + lo * 4),
                      C.c_void_p(st.data_ptr() + lo * D * 4) if st is not None else C.c_void_p(0),
                      _ptr(ws), ws.numel())
        if ctx["dedup"]:
            d, (src_dev, irr_dev, irr_rows_dev) = ctx["dplan"][i], ctx["dten"][i]
            check(lib.tmr_head_frames_dedup_fwd(*head_in, int(feats_f16), *common_in, _ptr(src_dev), _ptr(irr_dev), len(d["irr"]),
                                                _ptr(irr_rows_dev), len(d["irr_rows"]),
                                                d["row_base"], d["pb_rows"], self.seq, self.L, F, D, Cn,
                                                self.pad_mode, *common_out, stream))
        else:
            if feats_f16:
                raise TypeError("fp16 features are an input contract of the tensor-core bank-level path only (dedup=True, math 'f16')")
            check(lib.tmr_head_frames_fwd(*head_in, *common_in, self.seq, self.L, F, D, Cn, self.pad_mode,
                                          *common_out, ctx["mode"], stream))

    def run(self, feats, bank, out=None, want_st=False, graph=None):
        """All clips of the index in global clip order, features resident on the device.
        Returns dict(logits, pred, score[, St]) of device tensors.

        graph: None (default) - the second pass over the SAME buffers (features, bank, outputs, packed
        weights) is captured into a CUDA graph and later passes replay it (one launch instead of ~55, no
        per-launch tensor-map encoding on the host); True / False force or forbid that."""
        f16 = isinstance(feats, torch.Tensor) and feats.dtype == torch.float16
        feats = _dev(feats, "feats", torch.float16 if f16 else torch.float32)
        bank = _dev(bank, "bank")
        dev = feats.device
        if out is None:
            out = self._alloc_out(dev, want_st)
        ctx = self._prepare(dev, bank)
        esz = 2 if f16 else 4

        def enqueue():
            stream = _stream()
            for i, (lo, hi, fl, fh) in enumerate(ctx["plan"]):
                self._launch_batch(ctx, i, feats.data_ptr() + fl * F * esz, out, stream, f16)

        key = (feats.data_ptr(), bank.data_ptr(), tuple(sorted((k, v.data_ptr()) for k, v in out.items())),
               tuple(p.data_ptr() for p in ctx["packs"] if p is not None), ctx["ws"].data_ptr(), ctx["mode"],
               ctx["dedup"], str(dev))
        with torch.cuda.device(dev):
            if graph is False:
                enqueue()
            elif self._graph is not None and self._graph[0] == key:
                self._graph[1].replay()
            elif graph is True or self._last_key == key:
                enqueue()                                   # this pass, eagerly (also warms lazy attribute setup) ...
                g = torch.cuda.CUDAGraph()
                torch.cuda.current_stream(dev).synchronize()
                with torch.cuda.graph(g):                   # ... and its launch sequence for the following ones
                    enqueue()
                self._graph = (key, g, ctx, out, feats, bank)        # keeps every captured buffer alive
            else:
                enqueue()
        self._last_key = key
        return out

    def run_host(self, feats_host, bank, out=None, host_out=None, timeline=None):
        """Same pass with the per-frame features in (pinned) HOST memory: each batch's frames are
        copied H2D on a side stream into one of three staging buffers while earlier batches
        compute (three, not two: with two the copy of batch i+2 waits for batch i's kernels, which
        leaves the last full batch's compute exposed after the final copy); predictions and scores are copied back D2H at the end.  Returns
        (device outputs, (pred_host, score_host))."""
        bank = _dev(bank, "bank")
        dev = bank.device
        if not (isinstance(feats_host, torch.Tensor) and not feats_host.is_cuda
                and feats_host.dtype in (torch.float32, torch.float16)
                and feats_host.dim() == 2 and feats_host.shape[1] == F and feats_host.is_contiguous()):
            raise TypeError(f"feats_host must be a contiguous CPU float32 (or, optionally, float16) tensor (n_frames,{F})")
        f16 = feats_host.dtype == torch.float16
        if self.host_batch_clips != self.batch_clips:      # host streaming uses its own (smaller) batches
            if self._host_eng is None:
                self._host_eng = BankInference(self.model, self.index, self.seq, self.L, self.host_batch_clips,
                                               host_batch_clips=self.host_batch_clips, **self._ctor)
            self._host_eng.math_mode = self.math_mode
            return self._host_eng.run_host(feats_host, bank, out=out, host_out=host_out, timeline=timeline)
        if out is None:
            out = self._alloc_out(dev, False)
        ctx = self._prepare(dev, bank)
        plan = ctx["plan"]
        max_frames = max((fh - fl for _, _, fl, fh in plan), default=1)
        if (getattr(self, "_stage", None) is None or self._stage[0].shape[0] < max_frames or self._stage[0].device != dev
                or self._stage[0].dtype != feats_host.dtype):
            self._stage = [torch.empty((max_frames, F), dtype=feats_host.dtype, device=dev) for _ in range(3)]
            self._copy_stream = torch.cuda.Stream(device=dev)
        if host_out is None:
            n = len(self.starts_host)
            host_out = (torch.empty(n, dtype=torch.int64).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory())
        with torch.cuda.device(dev):
            compute = torch.cuda.current_stream()
            stream = _stream()
            timed = timeline is not None           # debug: per-batch (copy start, copy done, compute done) events
            copied = [torch.cuda.Event(enable_timing=timed) for _ in plan]
            freed = [None, None, None]
            self._copy_stream.wait_stream(compute)          # staging buffers may still be in use by earlier work
            for i, (lo, hi, fl, fh) in enumerate(plan):
                buf = self._stage[i % 3]
                with torch.cuda.stream(self._copy_stream):
                    if freed[i % 3] is not None:
                        self._copy_stream.wait_event(freed[i % 3])
                    if timed:
                        c0 = torch.cuda.Event(enable_timing=True)
                        c0.record(self._copy_stream)
                    buf[:fh - fl].copy_(feats_host[fl:fh], non_blocking=True)
                    copied[i].record(self._copy_stream)
                compute.wait_event(copied[i])
                self._launch_batch(ctx, i, buf.data_ptr(), out, stream, f16)
                freed[i % 3] = torch.cuda.Event(enable_timing=timed)
                freed[i % 3].record(compute)
                if timed:
                    timeline.append((hi - lo, c0, copied[i], freed[i % 3]))
            host_out[0].copy_(out["pred"], non_blocking=True)
            host_out[1].copy_(out["score"], non_blocking=True)
        return out, host_out

    def launches_per_run(self, feats_f16: bool = False) -> int:
        """Kernel launches of one run() (bench.py's gpu_launches), counted from the launch sequences in csrc/api.cu.
        LSTM, fp32 mode: projection + cell0 + (seq-1) steps.  Tensor-core mode: feature conversion (none for fp16
        features) + row->clip table + projection (step 0 fused) + step-0 fix-up + the recurrence (one launch, see lstm()).  Tail, fp32: gather | timeconv | q, u, attention, v,
        layernorm, out | fc_h_c, fc_c; tensor-core mode folds q, u into one GEMM and adds the fp16 conversions of the
        window, St and [St|y1]; the bank-level path replaces gather + conversion + timeconv over all clips by
        conversion(bank rows) + bankconv and, for the irregular clips of a batch, row-index gather + compact + raw
        bankconv + assemble (or gather + conversion + timeconv without the row list)."""
        mode = _mode(self.math_mode if self.math_mode is not None else self.model.math_mode)
        tc = 1 if self.model.time_conv is not None else 0
        tail = 6 + 2
        if mode != ops.TMR_MATH_F16:
            return (1 + 1 + (self.seq - 1) + 1 + tc + tail) * len(self.plan())
        tail -= 1                                      # u = W21 St + bu: one GEMM for q and u

        # one full round of the persistent recurrence grid: 2 tiles in flight x G groups of 8 CTA pairs x 256 clips
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count if torch.cuda.is_available() else 148
        rnd = 2 * ((sms // 2) // 8) * 256

        def lstm(b):
            if self.seq == 1:
                return (0 if feats_f16 else 1) + 1 + 1
            # the recurrence is ONE launch (small-batch kernel up to 512 clips, persistent kernel above), two when a
            # remainder of <= 512 clips beyond whole rounds of the persistent grid goes to the small-batch kernel
            rec = 2 if (b > rnd > 0 and 0 < b % rnd <= 512) else 1
            return (0 if feats_f16 else 1) + 1 + 1 + 1 + rec

        st_conv = 1 if self.seq == 1 else 0            # the recurrence kernels leave fp16(St) for the relation block
        if not self._use_dedup():
            # up to 512 clips the relation block + classifier are one fused launch
            return sum(lstm(hi - lo) + 1 + 1 + 2 * tc + ((1 + st_conv) if hi - lo <= 512 else (tail + 1 + st_conv))
                       for lo, hi, _, _ in self.plan())
        n = 0
        for (lo, hi, _, _), d in zip(self.plan(), self.dedup_plan()):
            n += (lstm(hi - lo) + (2 if d["pb_rows"] > 0 else 0) + (0 if not len(d["irr"]) else 4 if len(d["irr_rows"]) else 3)
                  + tail + 1 + st_conv)
        return n


# ---------------------------------------------------------------------------------------------
# multi-GPU: shard by video, no data-path collective
# ---------------------------------------------------------------------------------------------
def shard_videos(list_each_length, world_size: int):
    """Contiguous, frame-balanced partition of the videos: returns [(v_lo, v_hi)] per rank."""
    lens = np.asarray(list_each_length, dtype=np.int64)
    V = len(lens)
    cum = np.concatenate([[0], np.cumsum(lens)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        v = int(np.searchsorted(cum, target, side="left"))
        v = min(max(v, bounds[-1]), V)
        bounds.append(v)
    bounds.append(V)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


class VideoShard:
    """What one rank holds: the frames/rows of videos [v_lo, v_hi) plus the HALO needed for
    bit-exact reference windows — the first clips of video v_lo read rows from the tail of the
    previous videos (TRAIN:298-326 leak).  The halo is expressed as whole extra
    frames/rows in front of the shard so the same closed-form index applies locally."""

    def __init__(self, list_each_length, seq: int, L: int, v_lo: int, v_hi: int):
        lens = [int(v) for v in list_each_length]
        self.seq, self.L, self.v_lo, self.v_hi = seq, L, v_lo, v_hi
        cum = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        # walk back over previous videos until the keys s-1 .. s-L of the shard's first clips are covered
        h = v_lo
        need = L
        while h > 0 and need > 0:
            h -= 1
            need -= lens[h]
        self.h_lo = h                                   # first halo video
        self.frame_lo, self.frame_hi = int(cum[h]), int(cum[v_hi])
        self.own_frame_lo = int(cum[v_lo])
        self.local_lengths = lens[h:v_hi]
        rows_before = lambda v: int(sum(max(0, n - seq + 1) for n in lens[:v]))
        self.row_lo, self.row_hi = rows_before(h), rows_before(v_hi)
        self.own_row_lo = rows_before(v_lo)
        self.index = None

    def build_index(self):
        self.index = LFBIndex.from_lengths(self.local_lengths, self.seq)
        return self.index

    def own_local_starts(self):
        """Local (shard-relative) start frame ids of the clips this rank owns (halo clips excluded)."""
        from .lfb import get_useful_start_idx
        s = np.asarray(get_useful_start_idx(self.seq, self.local_lengths), dtype=np.int64)
        return s[s >= self.own_frame_lo - self.frame_lo]


# ---------------------------------------------------------------------------------------------
# upstream of the path: bank builder
# ---------------------------------------------------------------------------------------------
def build_bank(lfb_model, feats, list_each_length, seq: int = 10, batch_clips: int = None, math_mode=None):
    """The reference's LFB construction loop (train_non-local_mutiConv_resnet.py:684-756) with the
    backbone features precomputed: one LSTM pass per valid clip start, h at the last step written
    straight into the device bank (row r = r-th valid start in global order) instead of growing a
    numpy array with np.concatenate.  lfb_model: tmrnet_b200.resnet_lstm_LFB (or resnet_lstm)."""
    from .lfb import get_useful_start_idx
    feats = _dev(feats, "feats")
    starts_host = np.asarray(get_useful_start_idx(seq, list_each_length), dtype=np.int64)
    if feats.shape[0] != int(sum(int(v) for v in list_each_length)):
        raise ValueError("feats rows must equal the total number of frames")
    dev = feats.device
    packed = lfb_model.packed() if hasattr(lfb_model, "packed") else lfb_model.packs()[0]
    bank = torch.empty((len(starts_host), D), dtype=torch.float32, device=dev)
    starts = torch.from_numpy(starts_host).to(dev)
    lib = _lib.load()
    mode = _mode(math_mode if math_mode is not None else getattr(lfb_model, "math_mode", None))
    ws = None
    batch_clips = int(batch_clips) if batch_clips else MAX_BATCH_CLIPS
    with torch.cuda.device(dev):
        for lo in range(0, len(starts_host), batch_clips):
            hi = min(len(starts_host), lo + batch_clips)
            fl, fh = int(starts_host[lo]), int(starts_host[hi - 1]) + seq
            need = lib.tmr_lstm_workspace_bytes(fh - fl, hi - lo, D)
            if ws is None or ws.numel() < need:
                ws = _ws(need, dev)
            # rows are addressed relative to frame fl: pass local starts through a shifted copy
            local = starts[lo:hi] - fl
            check(lib.tmr_lstm_last_frames_fwd(_ptr(packed), C.c_void_p(feats.data_ptr() + fl * F * 4), fh - fl,
                                               _ptr(local), hi - lo, int(seq), F, D,
                                               C.c_void_p(bank.data_ptr() + lo * D * 4), _ptr(ws), ws.numel(), mode,
                                               _stream()))
    return bank
