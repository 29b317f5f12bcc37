/*
 * tmr_b200.h — C ABI of libtmr_b200.so: the TMRNet temporal-memory-relation head on B200 (sm_100a).
 *
 * The reference (lucieDLE/TMRNet, pure Python/PyTorch) has no FFI of its own; its seam for this
 * path is the nn.Module surface plus two free functions.  Every entry point
 * below names the reference code it replaces (paths relative to the reference root; NLB =
 * "code/Training TMRNet/NLBlock_MutiConv6_3.py", TRAIN =
 * "code/Training TMRNet/train_non-local_mutiConv_resnet.py", EVAL =
 * "code/eval/python/test_singlenet_phase_non-local_pretrained_2fc_copy_mutiConv6_resnest.py").
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every `const float*` / `float*` is a DEVICE pointer to fp32,
 *     contiguous, 16-byte aligned, unless the parameter name ends in `_host`;
 *   - the caller (PyTorch) owns all memory, including packed weights and workspaces, whose sizes
 *     come from the *_bytes() queries; the library keeps no mutable global state and is re-entrant
 *     per (device, stream) — the reference calls forward from one thread per GPU (TRAIN:776-778);
 *   - every function returns 0 on success, non-zero on error; tmr_last_error() returns the
 *     thread-local message of the last failure.  There is no CPU path: a missing device is an error;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, not synchronised.
 *   - D (bank row width / LSTM hidden) must be 512 and F (backbone width) 2048 as in the reference;
 *     L (memory length), seq (clip length), C (phases, <= 32) and B (clips) are runtime values.
 */
#ifndef TMR_B200_H_
#define TMR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMR_OK 0
#define TMR_ERR_ARG 1
#define TMR_ERR_CUDA 2
#define TMR_ERR_UNSUPPORTED 3

/* math mode of the GEMM-shaped stages */
#define TMR_MATH_FP32 0  /* fp32 FFMA on CUDA cores (exact-order reference path)          */
#define TMR_MATH_F16 1   /* tcgen05.mma kind::f16: operands converted to fp16 (round-to-nearest, saturating
                            at +-65504; the same 10 mantissa bits as TF32 at twice the rate and half the
                            bytes) by their producers, fp32 accumulate in TMEM; softmax / LayerNorm /
                            gates / max / final 512->C FC stay fp32                             */

/* window padding at the start of the bank / of a video */
#define TMR_PAD_REPEAT 0 /* reference semantics (TRAIN:298-326): repeat-fill, leaks into previous video */
#define TMR_PAD_ZERO 1   /* zero rows before the clip's own video start (what north_star describes)     */

const char* tmr_last_error(void);
int tmr_version(void);
/* Compute capability of the current device as major*10+minor (100 on B200); <0 on error. */
int tmr_device_arch(void);

/* ---- a1/a2: index tables --------------------------------------------------------------------
 * Replaces get_useful_start_idx (TRAIN:288-295) + the dict_*_start_idx_LFB construction
 * (TRAIN:643-644).  Host-side, closed form of the reference walk:
 *   frame2row[g] = bank row of g if g can start a clip, else the row of the smallest valid start
 *   > g (or -1 past the last one); frame2vstart[g] = first global frame of g's video.
 * lens_host[V] are the per-video frame counts.  Outputs are HOST arrays of sum(lens) int32
 * (frame2vstart_host may be NULL).  *n_rows_out receives the number of valid clip starts. */
int tmr_build_frame2row(const int64_t* lens_host, int V, int seq, int32_t* frame2row_host,
                        int32_t* frame2vstart_host, int64_t* n_rows_out);

/* ---- a3: memory-bank window gather ------------------------------------------------------------
 * Replaces get_long_feature + np.array + torch.Tensor(...).to(device) (TRAIN:298-326, 873-876).
 * out[b,k,:] = bank[row(b,k),:], row(b,k) = frame2row[starts[b]-k-1] (0 if the key is negative);
 * TMR_PAD_ZERO writes zeros where the key precedes the clip's own video (needs frame2vstart).
 * starts: device int64[B] global clip-start frame ids; frame2row/frame2vstart: device int32.
 * rows_out (nullable): device int32[B*L] receiving the gathered row ids (-1 for zero rows).
 * status (nullable): device int32, OR-ed with 1 when some starts[b] is not a valid clip start of the
 * table (the reference raises KeyError there, TRAIN:310) and with 2 when a table entry lies outside
 * [0, n_rows); such clips get an all-zero window and row ids -2.  The kernel never reads out of bounds,
 * with or without `status`; the caller zeroes *status before and reads it after (a sync). */
int tmr_gather_windows(const float* bank, int64_t n_rows, const int32_t* frame2row,
                       const int32_t* frame2vstart, int64_t n_frames, const int64_t* starts, int B,
                       int L, int D, int pad_mode, float* out, int32_t* rows_out, int32_t* status,
                       void* stream);

/* ---- weight packing (once per weight update; caller owns the packed buffers) -------------------
 * Inputs use the reference state-dict layouts. */
size_t tmr_timeconv_packed_bytes(int D);
/* time_conv.timeconv{1,2,3}.weight (D,D,{3,5,7}) + .bias (D) */
int tmr_timeconv_pack(const float* w3, const float* b3, const float* w5, const float* b5,
                      const float* w7, const float* b7, int D, void* packed, void* stream);

size_t tmr_nlblock_packed_bytes(int D);
/* nl_block.linear{1..4}.weight (D,D) + .bias (D), nl_block.layer_norm.weight/bias (1,D) */
int tmr_nlblock_pack(const float* w1, const float* b1, const float* w2, const float* b2,
                     const float* w3, const float* b3, const float* w4, const float* b4,
                     const float* ln_w, const float* ln_b, int D, void* packed, void* stream);

size_t tmr_lstm_packed_bytes(int F, int D);
/* lstm.weight_ih_l0 (4D,F), lstm.weight_hh_l0 (4D,D), lstm.bias_ih_l0 (4D), lstm.bias_hh_l0 (4D) */
int tmr_lstm_pack(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int F,
                  int D, void* packed, void* stream);

size_t tmr_classifier_packed_bytes(int D, int C);
/* fc_h_c.weight (D,2D) + bias (D), fc_c.weight (C,D) + bias (C) */
int tmr_classifier_pack(const float* w_h, const float* b_h, const float* w_c, const float* b_c,
                        int D, int C, void* packed, void* stream);

/* ---- a5: TimeConv.forward (NLB:43-79) ---------------------------------------------------------
 * x (B,L,D) -> out (B,L,D): out[b,k,c] = max(x[k], k>0 ? max(x[k],x[k-1]) : max(x[k],0),
 * conv3, conv5, conv7) with zero "same" padding inside each window.  Any L >= 1.
 * workspace >= tmr_timeconv_workspace_bytes(B,L,D) (used by TMR_MATH_F16 for the fp16 copy of x
 * that feeds the tensor cores; may be NULL in TMR_MATH_FP32). */
size_t tmr_timeconv_workspace_bytes(int B, int L, int D);
int tmr_timeconv_max_fwd(const void* packed, const float* x, int B, int L, int D, float* out,
                         void* workspace, size_t workspace_bytes, int math_mode, void* stream);

/* ---- a6: NLBlock.forward, eval mode (NLB:25-40) -----------------------------------------------
 * St (B,D), Lt (B,L,D) -> out (B,D).  workspace >= tmr_nlblock_workspace_bytes(B,D). */
/* The HBM-bound core of the relation block alone (NLB:30-34 with phi/g folded): for the folded query
 * u (B,D) = W2^T (W1 St + b1): out[b] = sum_k softmax_k((1/512)**0.5 * u[b].Lt[b,k]) Lt[b,k]  (B,D). */
int tmr_attention_fwd(const float* u, const float* Lt, int B, int L, int D, float* out, void* stream);
size_t tmr_nlblock_workspace_bytes(int B, int D);
int tmr_nlblock_fwd(const void* packed, const float* St, const float* Lt, int B, int L, int D,
                    float* out, void* workspace, size_t workspace_bytes, int math_mode,
                    void* stream);

/* ---- a7: nn.LSTM(2048,512,batch_first) from zero state, last step only (TRAIN:224,241-244) ------
 * Per-clip form: x (B,seq,F) -> out (B,D).  Also the bank builder's head half
 * (resnet_lstm_LFB.forward, TRAIN:277-285). */
size_t tmr_lstm_workspace_bytes(int64_t n_rows_x, int B, int D);
int tmr_lstm_last_fwd(const void* packed, const float* x, int B, int seq, int F, int D, float* out,
                      void* workspace, size_t workspace_bytes, int math_mode, void* stream);
/* Frame-deduplicated form: feats (n_frames,F) holds every frame once; clip b covers frames
 * starts[b] .. starts[b]+seq-1 (device int64[B]).  The input projection runs once per frame.
 * workspace >= tmr_lstm_workspace_bytes(n_frames, B, D).
 * PRECONDITION (not checked on the device): 0 <= starts[b] and starts[b]+seq <= n_frames for every b
 * (the recurrence reads projected row starts[b]+t).  Starts need not be distinct or sorted: clips that
 * share a start each get their own state (tested). */
int tmr_lstm_last_frames_fwd(const void* packed, const float* feats, int64_t n_frames,
                             const int64_t* starts, int B, int seq, int F, int D, float* out,
                             void* workspace, size_t workspace_bytes, int math_mode, void* stream);

/* Stage-1 model surface (code/models.py:38-48: LSTM over the clip, every step's h goes to the 512 -> C fc):
 * h of ALL steps, time-major out_tm (seq,B,D), fp32 CUDA-core path.  workspace >= tmr_lstm_workspace_bytes(B*seq,B,D). */
int tmr_lstm_seq_fwd(const void* packed, const float* x, int B, int seq, int F, int D, float* out_tm,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- a8 + a9: classifier and eval post-processing (TRAIN:249-252 eval mode, EVAL:122-125,491-493)
 * logits = fc_c(relu(fc_h_c([St || y1]))); score = max softmax probability; pred = first argmax.
 * logits (B,C) fp32, pred int64[B], score fp32[B] (pred/score nullable).
 * workspace >= tmr_classifier_workspace_bytes(B,D). */
size_t tmr_classifier_workspace_bytes(int B, int D);
int tmr_fc_argmax_fwd(const void* packed, const float* St, const float* y1, int B, int D, int C,
                      float* logits, int64_t* pred, float* score, void* workspace,
                      size_t workspace_bytes, int math_mode, void* stream);

/* ---- a6 + a8 + a9 in one call: everything of resnet_lstm.forward after the LSTM and the TimeConv
 * (TRAIN:245-252 eval mode: y1 = NLBlock(St, Lt); logits = fc_c(relu(fc_h_c([St || y1]))), EVAL:491-493 score/argmax).
 * St (B,D), Lt (B,L,D) -> logits (B,C), pred int64[B], score fp32[B] (pred/score nullable).
 * In TMR_MATH_F16 batches of up to 512 clips - the reference's own 120-clip calls - run as ONE launch
 * (csrc/umma_head_tail.cu); larger ones and TMR_MATH_FP32 as the launches of tmr_nlblock_fwd + tmr_fc_argmax_fwd.
 * workspace >= tmr_relation_head_workspace_bytes(B,D). */
size_t tmr_relation_head_workspace_bytes(int B, int D);
int tmr_relation_head_fwd(const void* nlblock_packed, const void* classifier_packed, const float* St,
                          const float* Lt, int B, int L, int D, int C, float* logits, int64_t* pred,
                          float* score, void* workspace, size_t workspace_bytes, int math_mode, void* stream);

/* ---- a11: whole head, resnet_lstm.forward minus `share` (TRAIN:237-253 / EVAL:110-126) --------
 * x (B,seq,F) backbone features, long_feature (B,L,D).  timeconv_packed may be NULL for the
 * NL-only wiring (train_only_non-local_pretrained.py:226-240).
 * TMR_MATH_F16, B <= 512 (the reference's own 120-clip calls): 7 launches - row table, feature conversion, input
 * projection (LSTM step 0 in its epilogue), the recurrence (csrc/umma_lstm_small.cu), window conversion, TimeConv,
 * relation block + classifier + score/argmax (csrc/umma_head_tail.cu). */
size_t tmr_head_workspace_bytes(int B, int seq, int L, int D);
int tmr_head_fwd(const void* lstm_packed, const void* timeconv_packed, const void* nlblock_packed,
                 const void* classifier_packed, const float* x, const float* long_feature, int B,
                 int seq, int L, int F, int D, int C, float* logits, int64_t* pred, float* score,
                 void* workspace, size_t workspace_bytes, int math_mode, void* stream);

/* ---- bank-level head: the eval loop body (EVAL:470-499) for clips taken from resident per-frame
 * features and a resident memory bank, without materialising per-clip copies of the frames:
 *   feats        (n_feat_frames,F): backbone features of global frames frame0 .. frame0+n_feat_frames-1,
 *                which must cover starts[b] .. starts[b]+seq-1 for every clip of the batch;
 *   bank         (n_rows,D), frame2row/frame2vstart (n_frames_total) as in tmr_gather_windows;
 *   starts       device int64[B] GLOBAL clip-start frame ids.
 * Runs: input projection once per frame -> LSTM -> window gather -> TimeConv -> NLBlock -> FCs ->
 * softmax score / argmax.  St_out (nullable, (B,D)) receives the LSTM state of each clip (the row a
 * bank builder would store, TRAIN:277-285).  workspace >= tmr_head_frames_workspace_bytes(...). */
size_t tmr_head_frames_workspace_bytes(int64_t n_feat_frames, int B, int L, int D);
int tmr_head_frames_fwd(const void* lstm_packed, const void* timeconv_packed,
                        const void* nlblock_packed, const void* classifier_packed,
                        const float* feats, int64_t n_feat_frames, int64_t frame0, const float* bank,
                        int64_t n_rows, const int32_t* frame2row, const int32_t* frame2vstart,
                        int64_t n_frames_total, const int64_t* starts, int B, int seq, int L, int F,
                        int D, int C, int pad_mode, float* logits, int64_t* pred, float* score,
                        float* St_out, void* workspace, size_t workspace_bytes, int math_mode,
                        void* stream);

/* ---- bank-level head with the TimeConv deduplicated per bank ROW (TMR_MATH_F16 only) ----------
 * Same contract and results (up to fp32 summation order) as tmr_head_frames_fwd, for clip batches
 * sorted by start frame.  A clip is REGULAR when its window is a contiguous run of bank rows, i.e.
 * it lies at least L clips into its video (starts[b] - frame2vstart[starts[b]] >= L); then slot k of
 * its window is bank row frame2row[starts[b]] - 1 - k and the TimeConv output of that slot depends
 * only on the row and on the slot's distance to the window edges, so the convolutions run once per
 * row (7.9 MFLOP/row instead of 236 MFLOP/clip) into PB[row][7 variants][D].  IRREGULAR clips (the
 * first L of every video, where the reference window repeat-fills / leaks into the previous video)
 * go through the per-clip gather + TimeConv.  The caller supplies the split, computed once per batch
 * plan on the host (tmrnet_b200.infer.BankInference does):
 *   src_idx          device int32[B]: regular clip -> (frame2row[starts[b]] - 1) - pb_row_base (>= L-1);
 *                    irregular clip -> -1 - j, j its position in irregular_starts;
 *   irregular_starts device int64[n_irregular] global start frames of the irregular clips;
 *   irregular_rows   device int32[n_irregular_rows], ascending: the DISTINCT bank rows the irregular clips'
 *                    windows touch (a few dozen around each video start).  Given them, the irregular clips'
 *                    TimeConv is assembled from per-row tap products as well (15 x D per listed row, summed
 *                    per (clip, slot) over the slot's own neighbours); with n_irregular_rows = 0 they go
 *                    through the per-clip gather + TimeConv instead;
 *   pb_row_base, pb_rows  bank row range covering slot L-1 of the first regular clip .. slot 0 of the
 *                    last one (pb_rows = 0 when the batch has no regular clip).  Needs L >= 6.
 *   feats, feats_f16 feats_f16 = 0: fp32 features (the reference's dtype; converted to fp16 once, on the device);
 *                    feats_f16 = 1: the caller already holds fp16 features (optional input contract that halves the
 *                    host->device bytes; bit-identical results, since the MMA operand is the same fp16 value).
 *   starts must be valid clip starts of the table, in range of feats, ascending (checked by BankInference on
 *   the host); they need not be distinct. */
/* The bank-level TimeConv alone: pb[(row - row_base)*7 + v][D] for bank rows row_base .. row_base+pb_rows-1,
 * v = 0 interior slot, 1..3 slot k = 0,1,2, 4..6 slot k = L-1, L-2, L-3 (see tmrnet_b200/csrc/umma_bankconv.cu).
 * pb receives FP16 values (pb_rows*7*D halves, 16-byte aligned): round-to-nearest of the fp32 TimeConv output. */
size_t tmr_bankconv_workspace_bytes(int64_t pb_rows, int D);
int tmr_bankconv_fwd(const void* timeconv_packed, const float* bank, int64_t n_rows, int64_t row_base,
                     int64_t pb_rows, int D, void* pb, void* workspace, size_t workspace_bytes, void* stream);
size_t tmr_head_frames_dedup_workspace_bytes(int64_t n_feat_frames, int B, int n_irregular, int n_irregular_rows,
                                             int64_t pb_rows, int L, int D);
int tmr_head_frames_dedup_fwd(const void* lstm_packed, const void* timeconv_packed,
                              const void* nlblock_packed, const void* classifier_packed,
                              const void* feats, int feats_f16, int64_t n_feat_frames, int64_t frame0, const float* bank,
                              int64_t n_rows, const int32_t* frame2row, const int32_t* frame2vstart,
                              int64_t n_frames_total, const int64_t* starts, int B, const int32_t* src_idx,
                              const int64_t* irregular_starts, int n_irregular, const int32_t* irregular_rows,
                              int n_irregular_rows, int64_t pb_row_base,
                              int64_t pb_rows, int seq, int L, int F, int D, int C, int pad_mode,
                              float* logits, int64_t* pred, float* score, float* St_out, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ---- a10: head training step (TRAIN:780,856-887) ----------------------------------------------
 * Forward in training mode + full backward of the head.  math_mode TMR_MATH_FP32: every GEMM in fp32 on CUDA cores
 * (gradients within 2e-4 of torch autograd over the fp64 graph); TMR_MATH_F16: the GEMMs of the forward and of the
 * backward (input and weight gradients) take fp16-rounded operands on the tensor cores with fp32 accumulation
 * (gradients within 3e-3 of each tensor's largest entry); element-wise math, reductions, the loss, saved activations,
 * gradients and parameters stay fp32 in both modes.  params/grads: 24 DEVICE pointers in
 * reference state-dict order and layouts:
 *   lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0,
 *   time_conv.timeconv{1,2,3}.{weight,bias} (NULL x6 for the NL-only wiring),
 *   nl_block.linear{1..4}.{weight,bias}, nl_block.layer_norm.{weight,bias}, fc_h_c.{weight,bias}, fc_c.{weight,bias}.
 * x (B,seq,F) backbone features and long_feature (B,L,D) carry no gradient (frozen features / bank).
 * Loss = CrossEntropyLoss(reduction='sum', weight=class_weight (nullable)) as TRAIN:780,883.
 * Dropout: p_nl on the NLBlock output (0.2, NLB:18,38), p_fc after fc_h_c (0.5, TRAIN:228,250); masks
 * come from a counter-based generator keyed by `seed` (pass 0/0 for the deterministic eval-graph grads).
 * grads are overwritten; logits (B,C), loss (1 float), pred int64[B] (nullable) are outputs. */
size_t tmr_head_train_workspace_bytes(int B, int seq, int L, int D, int F, int C);
int tmr_head_train_fwd_bwd(const float* const* params, float* const* grads, const float* x, const float* long_feature,
                           const int64_t* labels, const float* class_weight, int B, int seq, int L, int F, int D, int C,
                           float p_nl, float p_fc, uint64_t seed, float* logits, float* loss, int64_t* pred,
                           void* workspace, size_t workspace_bytes, int math_mode, void* stream);
/* The same step split at the logits, for torch.autograd: what lets the reference's loop body run unchanged -
 *   model.train(); outputs = model.forward(inputs, long_feature); loss = criterion(outputs, labels);
 *   loss.backward(); optimizer.step()                                              (TRAIN:876-887)
 * with stock torch losses and optimisers (SGD with or without nesterov, Adam; TRAIN:786-805).
 * tmr_head_train_fwd: training-mode forward, logits (B,C) out, activations saved in `workspace`.
 * tmr_head_train_bwd: given the SAME workspace (untouched in between), the same params / x / long_feature and
 * dlogits (B,C), overwrites the 24 grads.  logits_scratch: any (B,C) fp32 buffer. */
int tmr_head_train_fwd(const float* const* params, const float* x, const float* long_feature, int B, int seq, int L, int F,
                       int D, int C, float p_nl, float p_fc, uint64_t seed, float* logits, void* workspace,
                       size_t workspace_bytes, int math_mode, void* stream);
int tmr_head_train_bwd(const float* const* params, float* const* grads, const float* x, const float* long_feature,
                       const float* dlogits, int B, int seq, int L, int F, int D, int C, float* logits_scratch,
                       void* workspace, size_t workspace_bytes, int math_mode, void* stream);
/* torch.optim.SGD update on one flat fp32 tensor (momentum, weight decay, dampening 0, no nesterov),
 * TRAIN:797-805,887:  d = g + wd*p ; buf = first_step ? d : momentum*buf + d ; p -= lr*buf. */
int tmr_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                 float weight_decay, int first_step, void* stream);
/* The same update over `count` (<= 24) tensors in ONE launch, each with its own learning rate (the reference's
 * parameter groups, TRAIN:797-805).  params / grads / momentum_bufs: HOST arrays of device pointers (NULL = skip);
 * sizes, lrs: host arrays.  Element-wise identical to `count` tmr_sgd_step calls. */
int tmr_sgd_step_multi(float* const* params, const float* const* grads, float* const* momentum_bufs, const int64_t* sizes,
                       const float* lrs, int count, float momentum, float weight_decay, int first_step, void* stream);

/* ---- generic linear used by the stages above (exposed for tests) -------------------------------
 * out[M,N] = a[M,K] . w[N,K]^T + bias[N] (bias nullable), row-major, leading dims = K / K / N.
 * TMR_MATH_FP32: a, w are fp32.  TMR_MATH_F16: a, w are fp16 (K a multiple of 64); bias / out fp32. */
int tmr_linear_fwd(const void* a, const void* w, const float* bias, int64_t M, int N, int K,
                   float* out, int relu, int math_mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TMR_B200_H_ */
