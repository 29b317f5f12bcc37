// Internal launcher prototypes shared by api.cu and the kernel translation units.
#pragma once
#include "tmr_common.cuh"

namespace tmr {

// out[M,N] (ldo) = [a | a2][M,K] . w[N,K]^T (+bias[N]) (+residual[M,N]) (relu).  Columns k < k_split
// come from a (lda), the rest from a2 (lda2); a2 == nullptr means k_split == K.
struct LinearArgs {
  const float* a = nullptr;  int64_t lda = 0;
  const float* a2 = nullptr; int64_t lda2 = 0; int k_split = 0;
  const float* w = nullptr;  int64_t ldw = 0;
  const float* bias = nullptr;
  const float* residual = nullptr; int64_t ldr = 0;
  float* out = nullptr; int64_t ldo = 0;
  int64_t M = 0; int N = 0; int K = 0; int relu = 0;
  // tensor-core path: fp16 operands.  w16 = the pack's fp16 mirror (ldw elements).  A is either already fp16
  // (a16, lda elements: its producer converted it) or raw fp32 in a (| a2), converted into a_scratch [M,K] first.
  const half_t* w16 = nullptr;
  const half_t* a16 = nullptr;
  half_t* a_scratch = nullptr;
  bool a2_plus_a = false;       // with a_scratch and a2: the a2 half becomes fp16(a2 + a) (deferred residual)
  // tensor-core LSTM input projection only: row r starts clip row2clip[r] (-1: none) -> the epilogue also writes
  // step 0 of that clip's recurrence (c0 fp32, h0 fp16) and lstm_cell0_kernel is not launched
  const int32_t* row2clip = nullptr; float* c0 = nullptr; half_t* h0_16 = nullptr;
};

// ---- fp32 CUDA-core path (kernels_simt.cu) ----
int simt_linear(const LinearArgs& g, cudaStream_t st);
int simt_timeconv(const float* packed, const float* x, int B, int L, float* out, cudaStream_t st);
// One recurrent step t >= 1: gates = h_prev . Whh'^T + xp[row(m,t)], cell update into (h_out, c).
// row(m,t) = (starts ? starts[m] : m*seq) + t.
int simt_lstm_step(const float* whh, const float* xp, const int64_t* starts, int seq, int t,
                   const float* h_prev, float* h_out, float* c, int B, cudaStream_t st);

// ---- tcgen05 path, fp16 operands / fp32 accumulate (umma_*.cu); same contracts ----
int umma_linear(const LinearArgs& g, cudaStream_t st);
// x16 = fp16 copy of x (MMA operand); x = exact values for the identity / pool branches
int umma_timeconv(const float* packed, const float* x, const half_t* x16, int B, int L, float* out, cudaStream_t st);
// xp_base/xp_rows/xp_row0: the projected matrix itself ([xp_rows][4D], its row 0 = projected-row index
// xp_row0) so that warps whose 32 clips read 32 consecutive projected rows can fetch them by TMA
// h_prev is fp16; the step writes h as fp16 into h_out16 (it only feeds the next step's MMA) or, for the last
// step (h_out16 == nullptr), as fp32 into h_out.
int umma_lstm_step(const half_t* whh16, const float* xp, const int64_t* starts, int seq, int t,
                   const half_t* h_prev, half_t* h_out16, float* h_out, float* c, int B, cudaStream_t st,
                   const float* xp_base = nullptr, int64_t xp_rows = 0, int64_t xp_row0 = 0);
int umma_lstm_step_ws(const half_t* whh16, const float* xp, const int64_t* starts, int seq, int t,
                      const half_t* h_prev, half_t* h_out16, float* h_out, float* c, int B, cudaStream_t st,
                      const float* xp_base, int64_t xp_rows, int64_t xp_row0);
// all recurrent steps 1 .. seq-1 in one persistent launch (umma_lstm_persist.cu); TMR_ERR_UNSUPPORTED when the
// device cannot keep a group of 8 CTA pairs resident
int umma_lstm_persist(const half_t* whh16, const float* xp, const int64_t* starts, int seq, half_t* h16a, half_t* h16b,
                      float* h_out, const float* c0, int B, int32_t* flags, cudaStream_t st, const float* xp_base,
                      int64_t xp_rows, int64_t xp_row0);
int umma_lstm_persist_round_clips();
// the same for the reference's own batch sizes (umma_lstm_small.cu): 128-clip tiles x 32 CTAs of 64 gate columns,
// B <= umma_lstm_small_max_clips() (512 on a B200; 0 when the device cannot keep one tile's 32 CTAs resident)
int umma_lstm_small_max_clips();
int umma_lstm_small(const half_t* whh16, const float* xp, const int64_t* starts, int seq, half_t* h16a, half_t* h16b,
                    float* h_out, const float* c0, int B, int32_t* flags, cudaStream_t st);
// relation block + classifier of B <= umma_head_tail_max_clips() clips in one launch (umma_head_tail.cu);
// cls_packed == nullptr: relation block only (y1 with the residual, fp32 -> y1_out)
int umma_head_tail_max_clips();
int umma_head_tail(const float* nl_packed, const float* cls_packed, const float* St, const half_t* St16, const float* Lt,
                   int B, int L, int C, float* u, half_t* a16, half_t* y16, float* z, float* y1_out, float* logits,
                   int64_t* pred, float* score, int32_t* flags, cudaStream_t st);
// training forward of the recurrence in one launch (all steps, gates / c / h of every step saved; raw Whh layout)
int umma_lstm_train_fwd(const half_t* whh16, const float* xp, int B, int seq, float* gates, float* c, float* h,
                        half_t* h16a, half_t* h16b, int32_t* flags, cudaStream_t st);
// problems of at most one 128-row tile, cut by output column over N / 16 CTAs (umma_gemm_small.cu); called by umma_linear
int umma_linear_small(const LinearArgs& g, cudaStream_t st);
bool umma_available();
// bank-level TimeConv: pb[(row-row_base)*7 + variant][512] (fp16) for bank rows row_base .. +pb_rows-1
int umma_bankconv(const float* packed, const float* bank, const half_t* bank16, int64_t n_rows, int64_t r_lo,
                  int64_t r_cnt, int64_t row_base, int64_t pb_rows, half_t* pb, cudaStream_t st);

// ---- memory-bound kernels (kernels_mem.cu) ----
int launch_gather(const float* bank, int64_t n_rows, const int32_t* f2r, const int32_t* f2v,
                  int64_t n_frames, const int64_t* starts, int B, int L, int pad_mode, float* out,
                  int32_t* rows_out, cudaStream_t st, int32_t* status = nullptr);
// irregular clips of the bank-level path: compact + round the rows their windows touch, assemble their
// TimeConv from unshifted per-row tap products (umma_bankconv_raw)
int launch_compact_rows_half(const float* bank, const int32_t* rows, int n, half_t* out, cudaStream_t st);
int launch_irr_assemble(const float* q, const int32_t* crows, int n_c, const int32_t* wrows, const float* bank,
                        const float* b3, const float* b5, const float* b7, int n_clips, int L, float* out,
                        cudaStream_t st);
int umma_bankconv_raw(const float* packed, const half_t* rows16, int64_t n, float* q, cudaStream_t st);
// step 0 of the LSTM from zero state: c = sig(i)*tanh(g), h = sig(o)*tanh(c) from xp rows.
// h16 != nullptr: h goes out as fp16 (operand of the next step's MMA) instead of fp32 into h.
int launch_lstm_cell0(const float* xp, const int64_t* starts, int seq, float* h, half_t* h16, float* c, int B,
                      cudaStream_t st, bool fast_math = false);
// a[b,:] = sum_k softmax_k(scale * u[b].Lt[b,k]) Lt[b,k,:]; half_out: `a` receives fp16 (half_t[B,512])
int launch_attention(const float* u, const float* Lt, int B, int L, void* a, int half_out, cudaStream_t st);
// same attention over the bank-level TimeConv output (fp16 PB, see umma_bankconv.cu; irregular clips' rows are fp32)
int launch_attention_pb(const float* u, const half_t* pb, const float* lt_irr, const int32_t* src, int B, int L,
                        void* a, int half_out, cudaStream_t st);
// y = relu(layer_norm(v) * w + b) over rows of 512; half_out: y receives fp16
int launch_layernorm_relu(const float* v, const float* w, const float* b, int B, void* y,
                          int half_out, cudaStream_t st);
// row2clip[r] = clip whose first frame is projected row r (r = starts[b] - row0, or b * seq without starts), else -1
int launch_row2clip(const int64_t* starts, int seq, int B, int64_t n_rows, int64_t row0, int32_t* row2clip, cudaStream_t st);
// clips whose start collides with another clip's (lost the row2clip slot): step 0 from the projected row
int launch_lstm_cell0_fix(const float* xp, const int64_t* starts, int B, int64_t n_rows, int64_t row0,
                          const int32_t* row2clip, float* c, half_t* h16, cudaStream_t st);
// dst[n] = fp16(src[n])
int launch_to_half(const float* src, half_t* dst, int64_t n, cudaStream_t st);
// dst[M,K] = fp16([a | a2]) (columns < k_split from a (lda), the rest from a2 (lda2))
int launch_half_concat(const float* a, int64_t lda, const float* a2, int64_t lda2, int k_split, int K,
                       int64_t M, half_t* dst, cudaStream_t st, bool a2_plus_a = false);
// logits = z . Wc^T + bc ; score = max softmax prob ; pred = first argmax
int launch_fc_argmax(const float* z, const float* wc, const float* bc, int B, int C, float* logits,
                     int64_t* pred, float* score, cudaStream_t st);
int launch_pack_timeconv(const float* w3, const float* b3, const float* w5, const float* b5,
                         const float* w7, const float* b7, float* packed, cudaStream_t st);
int launch_pack_nlblock(const float* w1, const float* b1, const float* w2, const float* w3,
                        const float* b3, const float* w4, const float* b4, const float* lnw,
                        const float* lnb, float* packed, cudaStream_t st);
int launch_pack_lstm(const float* wih, const float* whh, const float* bih, const float* bhh,
                     float* packed, cudaStream_t st);
int launch_pack_classifier(const float* wh, const float* bh, const float* wc, const float* bc, int C,
                           float* packed, cudaStream_t st);

}  // namespace tmr
