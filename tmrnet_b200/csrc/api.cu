// extern "C" entry points of libtmr_b200.so (see include/tmr_b200.h for the contract and the
// reference code each one replaces).  Orchestration only: argument checks, workspace carving and
// kernel sequencing on the caller's stream.  No allocation, no synchronisation, no global state.
#include <stdarg.h>

#include <vector>
#include <mutex>

#include <stdlib.h>
#include "tmr_internal.h"

namespace tmr {

std::string& last_error_ref() {
  static thread_local std::string e;
  return e;
}
int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

static int check_mode(int math_mode) {
  TMR_CHECK_ARG(math_mode == TMR_MATH_FP32 || math_mode == TMR_MATH_F16, "unknown math_mode %d", math_mode);
  if (math_mode == TMR_MATH_F16 && !umma_available())
    return set_error(TMR_ERR_UNSUPPORTED, "TMR_MATH_F16 needs the tcgen05 kernels (sm_100a device + build)");
  return TMR_OK;
}
static int check_dims(int D, int F = kF) {
  TMR_CHECK_ARG(D == kD, "D=%d unsupported: the head is built for D=%d (reference hard-codes 512)", D, kD);
  TMR_CHECK_ARG(F == kF, "F=%d unsupported: the head is built for F=%d", F, kF);
  return TMR_OK;
}
static int do_linear(const LinearArgs& g, int mode, cudaStream_t st) {
  if (mode != TMR_MATH_F16) return simt_linear(g, st);
  LinearArgs h = g;
  if (!g.a16) {            // A operand is raw fp32: convert it to fp16 once, then feed the MMA
    TMR_CHECK_ARG(g.a_scratch, "linear: tensor-core mode needs an fp16 A operand or a scratch buffer");
    TMR_TRY(launch_half_concat(g.a, g.lda, g.a2, g.lda2, g.k_split, g.K, g.M, g.a_scratch, st, g.a2_plus_a));
    h.a16 = g.a_scratch; h.lda = g.K;
  }
  return umma_linear(h, st);
}

struct Carver {
  char* p; size_t left;
  Carver(void* ws, size_t bytes) : p((char*)ws), left(bytes) {}
  float* take(size_t n_floats) {
    const size_t b = align_up(n_floats * sizeof(float), 256);
    if (b > left) return nullptr;
    float* r = (float*)p; p += b; left -= b; return r;
  }
};
static inline size_t fbytes(size_t n_floats) { return align_up(n_floats * sizeof(float), 256); }

// ---- stage implementations on carved workspaces -------------------------------------------------
// In tensor-core mode every tensor that feeds a GEMM is converted to fp16 (round-to-nearest) by its
// producer (which then writes fp16 instead of fp32) or by a conversion pass into `scratch`; weights come
// from the fp16 mirror of the pack.  fp32 mode touches neither.  Workspace buffers keep their fp32 sizes.
struct NLWs { float* w0; float* w1; float* s; };
struct PbSrc { const half_t* pb; const float* lt_irr; const int32_t* src; };   // bank-level TimeConv output (PB in fp16)
// defer_residual (tensor-core head paths): `out` gets W4 r + b4 only; the consumer (classifier_impl with
// y1_plus_St) adds St while it converts [St || y1] for its GEMM
// St16: fp16(St) if its producer left it (the LSTM recurrence kernels do): the conversion pass of the first GEMM is skipped
static int nlblock_impl(const float* pk, const float* St, const float* Lt, int B, int L, float* out,
                        NLWs ws, int mode, cudaStream_t st, const PbSrc* pbs = nullptr, bool defer_residual = false,
                        const half_t* St16 = nullptr) {
  const bool tc = mode == TMR_MATH_F16;
  const float* w = pk;
  const half_t* w16 = mirror16<NLBlockPacked>(pk);
  half_t* w0h = reinterpret_cast<half_t*>(ws.w0);      // tensor-core mode: attention / LayerNorm write fp16 here
  LinearArgs g;
  if (tc) {
    // u = W2^T (W1 St + b1) = W21 St + bu: the query linear and the phi fold in ONE GEMM (W21, bu built at pack time)
    g = LinearArgs(); g.a = St; g.lda = kD; g.w16 = w16 + NLBlockPacked::w21_off; g.ldw = kD;
    g.bias = pk + NLBlockPacked::bu_off; g.out = ws.w1; g.ldo = kD; g.M = B; g.N = kD; g.K = kD;
    g.a_scratch = reinterpret_cast<half_t*>(ws.s);
    if (St16) { g.a16 = St16; g.lda = kD; }
    TMR_TRY(do_linear(g, mode, st));
  } else {
    // q = St W1^T + b1                                   (NLB:26-27)
    g = LinearArgs(); g.a = St; g.lda = kD; g.w = w + NLBlockPacked::w1_off; g.ldw = kD;
    g.bias = pk + NLBlockPacked::b1_off; g.out = ws.w0; g.ldo = kD; g.M = B; g.N = kD; g.K = kD;
    TMR_TRY(do_linear(g, mode, st));
    // u = W2^T q  (phi folded onto the query; b2 cancels in the softmax)     (NLB:28-30)
    g = LinearArgs(); g.a = ws.w0; g.lda = kD; g.w = w + NLBlockPacked::w2t_off; g.ldw = kD;
    g.out = ws.w1; g.ldo = kD; g.M = B; g.N = kD; g.K = kD;
    TMR_TRY(do_linear(g, mode, st));
  }
  // a = sum_k softmax(scale u.Lt_k) Lt_k                                   (NLB:30-34)
  if (pbs) TMR_TRY(launch_attention_pb(ws.w1, pbs->pb, pbs->lt_irr, pbs->src, B, L, ws.w0, tc, st));
  else TMR_TRY(launch_attention(ws.w1, Lt, B, L, ws.w0, tc, st));
  // v = W3 a + b3  (g folded after the weighted sum: sum_k p_k = 1)         (NLB:33-34)
  g = LinearArgs(); g.a = ws.w0; g.lda = kD; g.w = w + NLBlockPacked::w3_off; g.ldw = kD;
  if (tc) { g.a16 = w0h; g.w16 = w16 + NLBlockPacked::w3_off; }
  g.bias = pk + NLBlockPacked::b3_off; g.out = ws.w1; g.ldo = kD; g.M = B; g.N = kD; g.K = kD;
  TMR_TRY(do_linear(g, mode, st));
  // r = relu(LayerNorm(v))                                                 (NLB:35-36)
  TMR_TRY(launch_layernorm_relu(ws.w1, pk + NLBlockPacked::lnw_off, pk + NLBlockPacked::lnb_off, B, ws.w0, tc, st));
  // out = St + W4 r + b4   (dropout is the identity in eval)               (NLB:37-40)
  g = LinearArgs(); g.a = ws.w0; g.lda = kD; g.w = w + NLBlockPacked::w4_off; g.ldw = kD;
  if (tc) { g.a16 = w0h; g.w16 = w16 + NLBlockPacked::w4_off; }
  g.bias = pk + NLBlockPacked::b4_off; g.out = out; g.ldo = kD;
  if (!(defer_residual && tc)) { g.residual = St; g.ldr = kD; }
  g.M = B; g.N = kD; g.K = kD;
  return do_linear(g, mode, st);
}

struct LstmWs { float* xp; float* xr; float* h0; float* h1; float* c; };
// x_f16: x points to fp16 features (tensor-core mode only): the conversion pass is skipped, the MMA reads them in place
// st16: if the route taken also leaves fp16(h_T) somewhere in the workspace, *st16 points to it (else nullptr)
static int lstm_impl(const float* pk, const float* x, int64_t n_rows_x, const int64_t* starts, int B,
                     int seq, float* out, LstmWs ws, int mode, cudaStream_t st, int64_t frame0 = 0, bool x_f16 = false,
                     const half_t** st16 = nullptr) {
  const bool tc = mode == TMR_MATH_F16;
  if (st16) *st16 = nullptr;
  TMR_CHECK_ARG(!x_f16 || tc, "lstm: fp16 features need TMR_MATH_F16");
  const float* w = pk;
  const half_t* w16 = mirror16<LstmPacked>(pk);
  // input projection for every row of x once: xp = x Wih'^T + (b_ih + b_hh)', gate-interleaved
  LinearArgs g;
  g.a = x; g.lda = kF; g.w = w + LstmPacked::wih_off; g.ldw = kF; g.bias = pk + LstmPacked::bias_off;
  g.out = ws.xp; g.ldo = 4 * kD; g.M = n_rows_x; g.N = 4 * kD; g.K = kF;
  g.w16 = w16 + LstmPacked::wih_off;
  g.a_scratch = tc ? reinterpret_cast<half_t*>(ws.xr) : nullptr;
  if (x_f16) { g.a = nullptr; g.a16 = reinterpret_cast<const half_t*>(x); g.lda = kF; }
  // tensor-core mode, seq > 1: step 0 of the recurrence (zero state) rides in the projection's epilogue.  The
  // row -> clip table lives in the unused half of the fp16-feature slot (n_rows_x x 2048 x 2 of its 4 bytes used).
  const bool fuse0 = tc && seq > 1;
  if (fuse0) {
    int32_t* row2clip = reinterpret_cast<int32_t*>(ws.xr + (size_t)n_rows_x * (kF / 2));
    TMR_TRY(launch_row2clip(starts, seq, B, n_rows_x, frame0, row2clip, st));
    g.row2clip = row2clip; g.c0 = ws.c; g.h0_16 = reinterpret_cast<half_t*>(ws.h0);
  }
  TMR_TRY(do_linear(g, mode, st));
  if (fuse0 && starts)   // clips sharing a start with another clip lost the table slot: their step 0 runs here
    TMR_TRY(launch_lstm_cell0_fix(ws.xp, starts, B, n_rows_x, frame0, g.row2clip, ws.c, reinterpret_cast<half_t*>(ws.h0), st));
  const float* xp = ws.xp - frame0 * 4 * kD;   // rows addressed by GLOBAL frame id (starts[m] + t)
  // (Running the recurrence in L2-sized sub-batches - all steps of one before the next, so that the projected
  // rows consecutive clips share stay in L2 - was measured slower at every size: the extra pipeline ramps
  // cost more than the HBM re-reads; profiles/r1_final_ncu.md.)
  // t = 0 from zero state, then seq-1 recurrent steps; the last one writes `out` (always fp32).  In tensor-core
  // mode the intermediate h only feeds the next step's MMA and lives in fp16 (h0 / h1 hold half_t[B,512]).
  if (tc) {
    half_t* h16[2] = {reinterpret_cast<half_t*>(ws.h0), reinterpret_cast<half_t*>(ws.h1)};
    if (!fuse0) TMR_TRY(launch_lstm_cell0(xp, starts, seq, out, seq > 1 ? h16[0] : nullptr, ws.c, B, st, true));
    // ALL recurrent steps in one launch, c in registers, h exchanged through L2: up to 512 clips (the reference's own
    // 120-clip calls) by the latency-oriented kernel of umma_lstm_small.cu (32 CTAs of 64 gate columns per 128-clip
    // tile: 120 clips 128 -> 3x us per recurrence), larger batches by the throughput-oriented persistent kernel
    // of umma_lstm_persist.cu (CTA pairs of 256 clips x 256 columns, two tiles in flight).  The per-tile arrival
    // counters sit behind the row -> clip table in the unused part of the fp16-feature slot.
    if (fuse0 && env_int("TMR_LSTM_PERSIST", 1)) {
      const size_t spare = (size_t)n_rows_x * kF * 2;                       // bytes of ws.xr the fp16 features do not use
      const size_t table = align_up((size_t)n_rows_x * sizeof(int32_t), 256);
      const size_t nflags = 2 * (((size_t)B + 255) / 256);
      if (table + nflags * sizeof(int32_t) <= spare) {
        int32_t* flags = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws.xr) + (size_t)n_rows_x * kF * 2 + table);
        int rc = TMR_ERR_UNSUPPORTED;
        if (B <= umma_lstm_small_max_clips() && B <= env_int("TMR_LSTM_SMALL_MAX", 512)) {
          rc = umma_lstm_small(w16 + LstmPacked::whh_off, xp, starts, seq, h16[0], h16[1], out, ws.c, B, flags, st);
          if (rc == TMR_OK && st16) *st16 = h16[(seq - 1) & 1];
        } else if (B >= env_int("TMR_LSTM_PERSIST_MIN", 96)) {
          // Every round of the persistent grid (2 x G tiles of 256 clips) costs the same whether it is full or holds
          // one tile: a remainder of up to 512 clips beyond whole rounds (the 83 022-clip bench job: 78 clips = a 19th
          // round for one tile, 85 us) goes to the small-batch kernel instead (40 us), launched behind it.
          int b_main = B;
          const int round = umma_lstm_persist_round_clips();
          if (starts && round > 0 && B > round) {
            const int rem = B % round;
            const int small_max = umma_lstm_small_max_clips();
            if (rem > 0 && rem <= small_max && rem <= 512 && table + (nflags + 4) * sizeof(int32_t) <= spare) b_main = B - rem;
          }
          rc = umma_lstm_persist(w16 + LstmPacked::whh_off, xp, starts, seq, h16[0], h16[1], out, ws.c, b_main, flags, st,
                                 ws.xp, n_rows_x, frame0);
          if (rc == TMR_OK && b_main < B) {
            const size_t o = (size_t)b_main * kD;
            rc = umma_lstm_small(w16 + LstmPacked::whh_off, xp, starts + b_main, seq, h16[0] + o, h16[1] + o, out + o, ws.c + o,
                                 B - b_main, flags + nflags, st);
          }
          if (rc == TMR_OK && st16) *st16 = h16[(seq - 1) & 1];
        }
        if (rc != TMR_ERR_UNSUPPORTED) return rc;
      }
    }
    for (int t = 1; t < seq; ++t) {
      const bool last = (t == seq - 1);
      TMR_TRY(umma_lstm_step(w16 + LstmPacked::whh_off, xp, starts, seq, t, h16[(t - 1) & 1], last ? nullptr : h16[t & 1],
                             out, ws.c, B, st, ws.xp, n_rows_x, frame0));
    }
    return TMR_OK;
  }
  float* hcur = (seq == 1) ? out : ws.h0;
  TMR_TRY(launch_lstm_cell0(xp, starts, seq, hcur, nullptr, ws.c, B, st, false));
  for (int t = 1; t < seq; ++t) {
    const bool last = (t == seq - 1);
    float* hnext = last ? out : (hcur == ws.h0 ? ws.h1 : ws.h0);
    TMR_TRY(simt_lstm_step(w + LstmPacked::whh_off, xp, starts, seq, t, hcur, hnext, ws.c, B, st));
    hcur = hnext;
  }
  return TMR_OK;
}

struct ClsWs { float* z; float* s; };
static int classifier_impl(const float* pk, const float* St, const float* y1, int B, int C,
                           float* logits, int64_t* pred, float* score, ClsWs ws, int mode,
                           cudaStream_t st, bool y1_plus_St = false) {
  const bool tc = mode == TMR_MATH_F16;
  const float* w = pk;
  LinearArgs g;   // z = relu(fc_h_c([St || y1]))   (TRAIN:249-251, eval: dropout = identity)
  g.a = St; g.lda = kD; g.a2 = y1; g.lda2 = kD; g.k_split = kD; g.w = w + ClassifierPacked::wh_off;
  g.ldw = 2 * kD; g.bias = pk + ClassifierPacked::bh_off; g.out = ws.z; g.ldo = kD; g.M = B; g.N = kD;
  g.K = 2 * kD; g.relu = 1; g.a_scratch = tc ? reinterpret_cast<half_t*>(ws.s) : nullptr; g.a2_plus_a = tc && y1_plus_St;
  g.w16 = mirror16<ClassifierPacked>(pk) + ClassifierPacked::wh_off;
  TMR_TRY(do_linear(g, mode, st));
  // fc_c (512 -> C) + softmax score + argmax stay fp32 on CUDA cores
  return launch_fc_argmax(ws.z, pk + ClassifierPacked::wc_off, pk + ClassifierPacked::bc_off, B, C, logits,
                          pred, score, st);
}

// Relation block + classifier in ONE launch for the reference's own batch sizes (umma_head_tail.cu); St16 = fp16(St)
// if a producer left it, else it is converted here.  Returns TMR_ERR_UNSUPPORTED (nothing launched) when the batch
// is too large for the co-resident grid or the mode is fp32: the caller then runs the separate launches.
static int fused_tail(const float* nl_pk, const float* cls_pk, const float* St, const half_t* St16, const float* Lt,
                      int B, int L, int C, float* y1_out, float* logits, int64_t* pred, float* score, NLWs nl,
                      float* z, int32_t* flags, int mode, cudaStream_t st) {
  if (mode != TMR_MATH_F16 || B > umma_head_tail_max_clips() || !env_int("TMR_FUSED_TAIL", 1)) return TMR_ERR_UNSUPPORTED;
  half_t* s16 = reinterpret_cast<half_t*>(nl.s);        // [B][512] fp16(St) | [B][512] fp16(y)
  if (!St16) {
    TMR_TRY(launch_half_concat(St, kD, nullptr, 0, 0, kD, B, s16, st, false));
    St16 = s16;
  }
  return umma_head_tail(nl_pk, cls_pk, St, St16, Lt, B, L, C, nl.w1, reinterpret_cast<half_t*>(nl.w0), s16 + (size_t)B * kD,
                        z, y1_out, logits, pred, score, flags, st);
}

static int timeconv_impl(const float* pk, const float* x, int B, int L, float* out, float* xr, int mode,
                         cudaStream_t st) {
  if (mode != TMR_MATH_F16) return simt_timeconv(pk, x, B, L, out, st);
  half_t* x16 = reinterpret_cast<half_t*>(xr);
  TMR_TRY(launch_to_half(x, x16, (int64_t)B * L * kD, st));
  return umma_timeconv(pk, x, x16, B, L, out, st);
}

// ---- fork / join for the two independent halves of the bank-level head ------------------------------------------
// The LSTM chain (feature conversion, input projection, recurrence) and the bank-side chain (fp16 bank rows, TimeConv
// per bank row, the irregular clips' windows) do not depend on each other until the relation block.  On one stream every
// kernel's last, partly filled round leaves SMs idle (the recurrence of a 10 k-clip shard: 43 tiles over 18 slots per
// round); on two, the other chain's CTAs take those SMs.  One side stream + two events per device, created on first use
// (outside any capture: BankInference warms up before it captures); inside a stream capture the event wait forks the
// capture and the join closes it, so the graph simply has two branches.  Callers that share a device share the side
// stream: more ordering than needed, never less.
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; bool ok = false, tried = false; };
static SideStream* side_stream() {
  static SideStream per_dev[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> lock(mu);
  SideStream& ss = per_dev[dev];
  if (!ss.tried) {
    ss.tried = true;
    ss.ok = cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) == cudaSuccess;
    if (!ss.ok) cudaGetLastError();
  }
  return ss.ok ? &ss : nullptr;
}

}  // namespace tmr

using namespace tmr;

extern "C" {

const char* tmr_last_error(void) { return last_error_ref().c_str(); }
int tmr_version(void) { return 100; }

int tmr_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_error(TMR_ERR_CUDA, "no CUDA device"); return -1; }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}

int tmr_build_frame2row(const int64_t* lens_host, int V, int seq, int32_t* frame2row_host,
                        int32_t* frame2vstart_host, int64_t* n_rows_out) {
  TMR_CHECK_ARG(lens_host && frame2row_host && V >= 0 && seq >= 1, "build_frame2row: bad arguments");
  int64_t total = 0, rows = 0;
  for (int v = 0; v < V; ++v) {
    TMR_CHECK_ARG(lens_host[v] >= 0, "build_frame2row: negative video length");
    total += lens_host[v];
  }
  TMR_CHECK_ARG(total < (int64_t)INT32_MAX, "build_frame2row: more than 2^31 frames");
  // forward pass: rows of valid starts, -1 elsewhere
  int64_t g = 0;
  for (int v = 0; v < V; ++v) {
    const int64_t n = lens_host[v];
    const int64_t n_valid = n - seq + 1 > 0 ? n - seq + 1 : 0;
    for (int64_t j = 0; j < n; ++j, ++g) {
      frame2row_host[g] = (j < n_valid) ? (int32_t)(rows + j) : -1;
      if (frame2vstart_host) frame2vstart_host[g] = (int32_t)(g - j);
    }
    rows += n_valid;
  }
  // backward pass: an invalid frame takes the row of the smallest valid start above it — the row
  // the reference walk is still "remembering" when it reaches that key (TRAIN:312-323)
  int32_t nxt = -1;
  for (int64_t i = total - 1; i >= 0; --i) {
    if (frame2row_host[i] >= 0) nxt = frame2row_host[i];
    else frame2row_host[i] = nxt;
  }
  if (n_rows_out) *n_rows_out = rows;
  return TMR_OK;
}

int tmr_gather_windows(const float* bank, int64_t n_rows, const int32_t* frame2row,
                       const int32_t* frame2vstart, int64_t n_frames, const int64_t* starts, int B,
                       int L, int D, int pad_mode, float* out, int32_t* rows_out, int32_t* status,
                       void* stream) {
  TMR_TRY(check_dims(D));
  TMR_CHECK_ARG(B >= 0 && L >= 1, "gather: bad B=%d L=%d", B, L);
  TMR_CHECK_ARG(pad_mode == TMR_PAD_REPEAT || pad_mode == TMR_PAD_ZERO, "gather: bad pad_mode %d", pad_mode);
  if (B == 0) return TMR_OK;
  TMR_CHECK_ARG(bank && frame2row && starts && (out || rows_out), "gather: null pointer");
  TMR_CHECK_ARG(pad_mode != TMR_PAD_ZERO || frame2vstart, "gather: TMR_PAD_ZERO needs frame2vstart");
  TMR_CHECK_ARG(aligned16(bank) && aligned16(out), "gather: bank/out must be 16-byte aligned");
  TMR_CHECK_ARG(n_rows > 0 && n_frames > 0, "gather: empty bank");
  return launch_gather(bank, n_rows, frame2row, frame2vstart, n_frames, starts, B, L, pad_mode, out,
                       rows_out, (cudaStream_t)stream, status);
}

size_t tmr_timeconv_packed_bytes(int D) { return D == kD ? TimeConvPacked::total * sizeof(float) : 0; }
int tmr_timeconv_pack(const float* w3, const float* b3, const float* w5, const float* b5,
                      const float* w7, const float* b7, int D, void* packed, void* stream) {
  TMR_TRY(check_dims(D));
  TMR_CHECK_ARG(w3 && b3 && w5 && b5 && w7 && b7 && packed, "timeconv_pack: null pointer");
  return launch_pack_timeconv(w3, b3, w5, b5, w7, b7, (float*)packed, (cudaStream_t)stream);
}

size_t tmr_nlblock_packed_bytes(int D) { return D == kD ? NLBlockPacked::total * sizeof(float) : 0; }
int tmr_nlblock_pack(const float* w1, const float* b1, const float* w2, const float* b2,
                     const float* w3, const float* b3, const float* w4, const float* b4,
                     const float* ln_w, const float* ln_b, int D, void* packed, void* stream) {
  TMR_TRY(check_dims(D));
  (void)b2;  // cancels inside the softmax over L
  TMR_CHECK_ARG(w1 && b1 && w2 && w3 && b3 && w4 && b4 && ln_w && ln_b && packed, "nlblock_pack: null pointer");
  return launch_pack_nlblock(w1, b1, w2, w3, b3, w4, b4, ln_w, ln_b, (float*)packed, (cudaStream_t)stream);
}

size_t tmr_lstm_packed_bytes(int F, int D) { return (D == kD && F == kF) ? LstmPacked::total * sizeof(float) : 0; }
int tmr_lstm_pack(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int F,
                  int D, void* packed, void* stream) {
  TMR_TRY(check_dims(D, F));
  TMR_CHECK_ARG(w_ih && w_hh && b_ih && b_hh && packed, "lstm_pack: null pointer");
  return launch_pack_lstm(w_ih, w_hh, b_ih, b_hh, (float*)packed, (cudaStream_t)stream);
}

size_t tmr_classifier_packed_bytes(int D, int C) {
  return (D == kD && C >= 1 && C <= ClassifierPacked::kMaxC) ? ClassifierPacked::total * sizeof(float) : 0;
}
int tmr_classifier_pack(const float* w_h, const float* b_h, const float* w_c, const float* b_c,
                        int D, int C, void* packed, void* stream) {
  TMR_TRY(check_dims(D));
  TMR_CHECK_ARG(C >= 1 && C <= ClassifierPacked::kMaxC, "classifier_pack: C=%d out of range [1,%d]", C, ClassifierPacked::kMaxC);
  TMR_CHECK_ARG(w_h && b_h && w_c && b_c && packed, "classifier_pack: null pointer");
  return launch_pack_classifier(w_h, b_h, w_c, b_c, C, (float*)packed, (cudaStream_t)stream);
}

size_t tmr_timeconv_workspace_bytes(int B, int L, int D) { return fbytes((size_t)(B > 0 ? B : 1) * L * D); }
int tmr_timeconv_max_fwd(const void* packed, const float* x, int B, int L, int D, float* out,
                         void* workspace, size_t workspace_bytes, int math_mode, void* stream) {
  TMR_TRY(check_dims(D));
  TMR_TRY(check_mode(math_mode));
  TMR_CHECK_ARG(B >= 0 && L >= 1, "timeconv: bad B=%d L=%d", B, L);
  if (B == 0) return TMR_OK;
  TMR_CHECK_ARG(packed && x && out, "timeconv: null pointer");
  TMR_CHECK_ARG(aligned16(x) && aligned16(out) && aligned16(packed), "timeconv: pointers must be 16-byte aligned");
  TMR_CHECK_ARG(x != out, "timeconv: in-place not supported");
  float* xr = nullptr;
  if (math_mode == TMR_MATH_F16) {
    TMR_CHECK_ARG(workspace && aligned16(workspace), "timeconv: TMR_MATH_F16 needs a workspace");
    Carver cv(workspace, workspace_bytes);
    xr = cv.take((size_t)B * L * kD);
    TMR_CHECK_ARG(xr, "timeconv: workspace too small (%zu < %zu)", workspace_bytes, tmr_timeconv_workspace_bytes(B, L, D));
  }
  return timeconv_impl((const float*)packed, x, B, L, out, xr, math_mode, (cudaStream_t)stream);
}

int tmr_attention_fwd(const float* u, const float* Lt, int B, int L, int D, float* out, void* stream) {
  TMR_TRY(check_dims(D));
  TMR_CHECK_ARG(B >= 0 && L >= 1, "attention: bad B=%d L=%d", B, L);
  if (B == 0) return TMR_OK;
  TMR_CHECK_ARG(u && Lt && out && aligned16(u) && aligned16(Lt) && aligned16(out), "attention: null or unaligned pointer");
  return launch_attention(u, Lt, B, L, out, 0, (cudaStream_t)stream);
}

size_t tmr_nlblock_workspace_bytes(int B, int D) { return 3 * fbytes((size_t)(B > 0 ? B : 1) * D); }
int tmr_nlblock_fwd(const void* packed, const float* St, const float* Lt, int B, int L, int D,
                    float* out, void* workspace, size_t workspace_bytes, int math_mode, void* stream) {
  TMR_TRY(check_dims(D));
  TMR_TRY(check_mode(math_mode));
  TMR_CHECK_ARG(B >= 0 && L >= 1, "nlblock: bad B=%d L=%d", B, L);
  if (B == 0) return TMR_OK;
  TMR_CHECK_ARG(packed && St && Lt && out && workspace, "nlblock: null pointer");
  TMR_CHECK_ARG(aligned16(St) && aligned16(Lt) && aligned16(out) && aligned16(workspace) && aligned16(packed),
                "nlblock: pointers must be 16-byte aligned");
  Carver cv(workspace, workspace_bytes);
  NLWs ws;
  ws.w0 = cv.take((size_t)B * kD); ws.w1 = cv.take((size_t)B * kD); ws.s = cv.take((size_t)B * kD);
  TMR_CHECK_ARG(ws.w0 && ws.w1 && ws.s, "nlblock: workspace too small (%zu < %zu)", workspace_bytes, tmr_nlblock_workspace_bytes(B, D));
  {   // one launch for the reference's batch sizes; without the classifier the fp16(y) half of ws.s is free for the
      // tile arrival counters
    int32_t* flags = reinterpret_cast<int32_t*>(reinterpret_cast<half_t*>(ws.s) + (size_t)B * kD);
    const int rc = fused_tail((const float*)packed, nullptr, St, nullptr, Lt, B, L, 0, out, nullptr, nullptr, nullptr, ws,
                              nullptr, flags, math_mode, (cudaStream_t)stream);
    if (rc != TMR_ERR_UNSUPPORTED) return rc;
  }
  return nlblock_impl((const float*)packed, St, Lt, B, L, out, ws, math_mode, (cudaStream_t)stream);
}

size_t tmr_lstm_workspace_bytes(int64_t n_rows_x, int B, int D) {
  const size_t r = (size_t)(n_rows_x > 0 ? n_rows_x : 1);
  return fbytes(r * 4 * D) + fbytes(r * kF) + 3 * fbytes((size_t)(B > 0 ? B : 1) * D);
}
static bool carve_lstm(Carver& cv, int64_t n_rows_x, int B, LstmWs& ws) {
  ws.xp = cv.take((size_t)n_rows_x * 4 * kD);
  ws.xr = cv.take((size_t)n_rows_x * kF);
  ws.h0 = cv.take((size_t)B * kD); ws.h1 = cv.take((size_t)B * kD); ws.c = cv.take((size_t)B * kD);
  return ws.xp && ws.xr && ws.h0 && ws.h1 && ws.c;
}
static int lstm_entry(const void* packed, const float* x, int64_t n_rows_x, const int64_t* starts, int B,
                      int seq, int F, int D, float* out, void* workspace, size_t workspace_bytes,
                      int math_mode, void* stream) {
  TMR_TRY(check_dims(D, F));
  TMR_TRY(check_mode(math_mode));
  TMR_CHECK_ARG(B >= 0 && seq >= 1 && n_rows_x >= 0, "lstm: bad B=%d seq=%d", B, seq);
  if (B == 0) return TMR_OK;
  TMR_CHECK_ARG(packed && x && out && workspace, "lstm: null pointer");
  TMR_CHECK_ARG(aligned16(x) && aligned16(out) && aligned16(workspace) && aligned16(packed), "lstm: pointers must be 16-byte aligned");
  Carver cv(workspace, workspace_bytes);
  LstmWs ws;
  TMR_CHECK_ARG(carve_lstm(cv, n_rows_x, B, ws), "lstm: workspace too small (%zu < %zu)", workspace_bytes,
                tmr_lstm_workspace_bytes(n_rows_x, B, D));
  return lstm_impl((const float*)packed, x, n_rows_x, starts, B, seq, out, ws, math_mode, (cudaStream_t)stream);
}
int tmr_lstm_last_fwd(const void* packed, const float* x, int B, int seq, int F, int D, float* out,
                      void* workspace, size_t workspace_bytes, int math_mode, void* stream) {
  return lstm_entry(packed, x, (int64_t)B * seq, nullptr, B, seq, F, D, out, workspace, workspace_bytes,
                    math_mode, stream);
}
int tmr_lstm_last_frames_fwd(const void* packed, const float* feats, int64_t n_frames,
                             const int64_t* starts, int B, int seq, int F, int D, float* out,
                             void* workspace, size_t workspace_bytes, int math_mode, void* stream) {
  TMR_CHECK_ARG(B == 0 || starts, "lstm_frames: starts is null");
  return lstm_entry(packed, feats, n_frames, starts, B, seq, F, D, out, workspace, workspace_bytes,
                    math_mode, stream);
}

/* Stage-1 surface (code/models.py:38-48): h of EVERY step, time-major out_tm[t][b][:] (fp32 CUDA-core path). */
int tmr_lstm_seq_fwd(const void* packed, const float* x, int B, int seq, int F, int D, float* out_tm,
                     void* workspace, size_t workspace_bytes, void* stream) {
  TMR_TRY(check_dims(D, F));
  TMR_CHECK_ARG(B >= 0 && seq >= 1, "lstm_seq: bad B=%d seq=%d", B, seq);
  if (B == 0) return TMR_OK;
  TMR_CHECK_ARG(packed && x && out_tm && workspace, "lstm_seq: null pointer");
  TMR_CHECK_ARG(aligned16(x) && aligned16(out_tm) && aligned16(workspace) && aligned16(packed), "lstm_seq: pointers must be 16-byte aligned");
  Carver cv(workspace, workspace_bytes);
  LstmWs ws;
  TMR_CHECK_ARG(carve_lstm(cv, (int64_t)B * seq, B, ws), "lstm_seq: workspace too small (%zu < %zu)", workspace_bytes,
                tmr_lstm_workspace_bytes((int64_t)B * seq, B, D));
  cudaStream_t st = (cudaStream_t)stream;
  const float* pk = (const float*)packed;
  LinearArgs g;
  g.a = x; g.lda = kF; g.w = pk + LstmPacked::wih_off; g.ldw = kF; g.bias = pk + LstmPacked::bias_off;
  g.out = ws.xp; g.ldo = 4 * kD; g.M = (int64_t)B * seq; g.N = 4 * kD; g.K = kF;
  TMR_TRY(simt_linear(g, st));
  const size_t step = (size_t)B * kD;
  TMR_TRY(launch_lstm_cell0(ws.xp, nullptr, seq, out_tm, nullptr, ws.c, B, st, false));
  for (int t = 1; t < seq; ++t)
    TMR_TRY(simt_lstm_step(pk + LstmPacked::whh_off, ws.xp, nullptr, seq, t, out_tm + (t - 1) * step, out_tm + t * step, ws.c, B, st));
  return TMR_OK;
}

size_t tmr_classifier_workspace_bytes(int B, int D) { return 3 * fbytes((size_t)(B > 0 ? B : 1) * D); }
static bool carve_cls(Carver& cv, int B, ClsWs& ws) {
  ws.z = cv.take((size_t)B * kD); ws.s = cv.take((size_t)B * 2 * kD);
  return ws.z && ws.s;
}
int tmr_fc_argmax_fwd(const void* packed, const float* St, const float* y1, int B, int D, int C,
                      float* logits, int64_t* pred, float* score, void* workspace,
                      size_t workspace_bytes, int math_mode, void* stream) {
  TMR_TRY(check_dims(D));
  TMR_TRY(check_mode(math_mode));
  TMR_CHECK_ARG(C >= 1 && C <= ClassifierPacked::kMaxC, "fc_argmax: C=%d out of range", C);
  TMR_CHECK_ARG(B >= 0, "fc_argmax: bad B");
  if (B == 0) return TMR_OK;
  TMR_CHECK_ARG(packed && St && y1 && logits && workspace, "fc_argmax: null pointer");
  TMR_CHECK_ARG(aligned16(St) && aligned16(y1) && aligned16(workspace) && aligned16(packed), "fc_argmax: pointers must be 16-byte aligned");
  Carver cv(workspace, workspace_bytes);
  ClsWs ws;
  TMR_CHECK_ARG(carve_cls(cv, B, ws), "fc_argmax: workspace too small");
  return classifier_impl((const float*)packed, St, y1, B, C, logits, pred, score, ws, math_mode,
                         (cudaStream_t)stream);
}

// shared tail of the two head entry points: St, window -> logits
struct HeadWs { float* Lt; float* xr; float* St; float* y1; NLWs nl; ClsWs cls; };
static bool carve_head(Carver& cv, int B, int L, HeadWs& ws) {
  ws.Lt = cv.take((size_t)B * L * kD); ws.xr = cv.take((size_t)B * L * kD);
  ws.St = cv.take((size_t)B * kD); ws.y1 = cv.take((size_t)B * kD);
  ws.nl.w0 = cv.take((size_t)B * kD); ws.nl.w1 = cv.take((size_t)B * kD); ws.nl.s = cv.take((size_t)B * kD);
  return ws.Lt && ws.xr && ws.St && ws.y1 && ws.nl.w0 && ws.nl.w1 && ws.nl.s && carve_cls(cv, B, ws.cls);
}
static size_t head_tail_bytes(size_t b, int L, int D) {
  return 2 * fbytes(b * L * D) + 2 * fbytes(b * D) + tmr_nlblock_workspace_bytes((int)b, D) +
         tmr_classifier_workspace_bytes((int)b, D);
}
static int head_tail(const void* timeconv_packed, const void* nlblock_packed, const void* classifier_packed,
                     const float* St, const float* window, int B, int L, int C, float* logits, int64_t* pred,
                     float* score, HeadWs& ws, int mode, cudaStream_t st, const half_t* St16 = nullptr) {
  const float* Lt_in = window;     // NL-only wiring: Lt = long_feature
  if (timeconv_packed) {
    TMR_TRY(timeconv_impl((const float*)timeconv_packed, window, B, L, ws.Lt, ws.xr, mode, st));
    Lt_in = ws.Lt;
  }
  {   // the reference's own batch sizes: relation block + classifier in one launch (counters in the unused [St || y] slot)
    const int rc = fused_tail((const float*)nlblock_packed, (const float*)classifier_packed, St, St16, Lt_in, B, L, C, nullptr,
                              logits, pred, score, ws.nl, ws.cls.z, reinterpret_cast<int32_t*>(ws.cls.s), mode, st);
    if (rc != TMR_ERR_UNSUPPORTED) return rc;
  }
  TMR_TRY(nlblock_impl((const float*)nlblock_packed, St, Lt_in, B, L, ws.y1, ws.nl, mode, st, nullptr, true, St16));
  return classifier_impl((const float*)classifier_packed, St, ws.y1, B, C, logits, pred, score, ws.cls, mode, st, true);
}

size_t tmr_relation_head_workspace_bytes(int B, int D) {
  return fbytes((size_t)(B > 0 ? B : 1) * D) + tmr_nlblock_workspace_bytes(B, D) + tmr_classifier_workspace_bytes(B, D);
}
int tmr_relation_head_fwd(const void* nlblock_packed, const void* classifier_packed, const float* St, const float* Lt,
                          int B, int L, int D, int C, float* logits, int64_t* pred, float* score, void* workspace,
                          size_t workspace_bytes, int math_mode, void* stream) {
  TMR_TRY(check_dims(D));
  TMR_TRY(check_mode(math_mode));
  TMR_CHECK_ARG(B >= 0 && L >= 1, "relation_head: bad B=%d L=%d", B, L);
  TMR_CHECK_ARG(C >= 1 && C <= ClassifierPacked::kMaxC, "relation_head: C=%d out of range", C);
  if (B == 0) return TMR_OK;
  TMR_CHECK_ARG(nlblock_packed && classifier_packed && St && Lt && logits && workspace, "relation_head: null pointer");
  TMR_CHECK_ARG(aligned16(St) && aligned16(Lt) && aligned16(workspace), "relation_head: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv(workspace, workspace_bytes);
  NLWs nl; ClsWs cls;
  float* y1 = cv.take((size_t)B * kD);
  nl.w0 = cv.take((size_t)B * kD); nl.w1 = cv.take((size_t)B * kD); nl.s = cv.take((size_t)B * kD);
  TMR_CHECK_ARG(y1 && nl.w0 && nl.w1 && nl.s && carve_cls(cv, B, cls), "relation_head: workspace too small (%zu < %zu)",
                workspace_bytes, tmr_relation_head_workspace_bytes(B, D));
  const int rc = fused_tail((const float*)nlblock_packed, (const float*)classifier_packed, St, nullptr, Lt, B, L, C, nullptr,
                            logits, pred, score, nl, cls.z, reinterpret_cast<int32_t*>(cls.s), math_mode, st);
  if (rc != TMR_ERR_UNSUPPORTED) return rc;
  TMR_TRY(nlblock_impl((const float*)nlblock_packed, St, Lt, B, L, y1, nl, math_mode, st, nullptr, true));
  return classifier_impl((const float*)classifier_packed, St, y1, B, C, logits, pred, score, cls, math_mode, st, true);
}

size_t tmr_head_workspace_bytes(int B, int seq, int L, int D) {
  const size_t b = (size_t)(B > 0 ? B : 1);
  return tmr_lstm_workspace_bytes((int64_t)b * seq, (int)b, D) + head_tail_bytes(b, L, D);
}
int tmr_head_fwd(const void* lstm_packed, const void* timeconv_packed, const void* nlblock_packed,
                 const void* classifier_packed, const float* x, const float* long_feature, int B,
                 int seq, int L, int F, int D, int C, float* logits, int64_t* pred, float* score,
                 void* workspace, size_t workspace_bytes, int math_mode, void* stream) {
  TMR_TRY(check_dims(D, F));
  TMR_TRY(check_mode(math_mode));
  TMR_CHECK_ARG(B >= 0 && seq >= 1 && L >= 1, "head: bad B=%d seq=%d L=%d", B, seq, L);
  TMR_CHECK_ARG(C >= 1 && C <= ClassifierPacked::kMaxC, "head: C=%d out of range", C);
  if (B == 0) return TMR_OK;
  TMR_CHECK_ARG(lstm_packed && nlblock_packed && classifier_packed && x && long_feature && logits && workspace,
                "head: null pointer");
  TMR_CHECK_ARG(aligned16(x) && aligned16(long_feature) && aligned16(workspace), "head: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv(workspace, workspace_bytes);
  LstmWs lw; HeadWs hw;
  TMR_CHECK_ARG(carve_lstm(cv, (int64_t)B * seq, B, lw) && carve_head(cv, B, L, hw),
                "head: workspace too small (%zu < %zu)", workspace_bytes, tmr_head_workspace_bytes(B, seq, L, D));
  const half_t* st16 = nullptr;
  TMR_TRY(lstm_impl((const float*)lstm_packed, x, (int64_t)B * seq, nullptr, B, seq, hw.St, lw, math_mode, st, 0, false, &st16));
  return head_tail(timeconv_packed, nlblock_packed, classifier_packed, hw.St, long_feature, B, L, C, logits, pred,
                   score, hw, math_mode, st, st16);
}

size_t tmr_head_frames_workspace_bytes(int64_t n_feat_frames, int B, int L, int D) {
  const size_t b = (size_t)(B > 0 ? B : 1);
  return tmr_lstm_workspace_bytes(n_feat_frames, (int)b, D) + fbytes(b * L * D) + head_tail_bytes(b, L, D);
}
int tmr_head_frames_fwd(const void* lstm_packed, const void* timeconv_packed,
                        const void* nlblock_packed, const void* classifier_packed,
                        const float* feats, int64_t n_feat_frames, int64_t frame0, const float* bank,
                        int64_t n_rows, const int32_t* frame2row, const int32_t* frame2vstart,
                        int64_t n_frames_total, const int64_t* starts, int B, int seq, int L, int F,
                        int D, int C, int pad_mode, float* logits, int64_t* pred, float* score,
                        float* St_out, void* workspace, size_t workspace_bytes, int math_mode,
                        void* stream) {
  TMR_TRY(check_dims(D, F));
  TMR_TRY(check_mode(math_mode));
  TMR_CHECK_ARG(B >= 0 && seq >= 1 && L >= 1 && n_feat_frames >= 0 && frame0 >= 0, "head_frames: bad sizes");
  TMR_CHECK_ARG(C >= 1 && C <= ClassifierPacked::kMaxC, "head_frames: C=%d out of range", C);
  TMR_CHECK_ARG(pad_mode == TMR_PAD_REPEAT || pad_mode == TMR_PAD_ZERO, "head_frames: bad pad_mode");
  if (B == 0) return TMR_OK;
  TMR_CHECK_ARG(lstm_packed && nlblock_packed && classifier_packed && feats && bank && frame2row && starts &&
                logits && workspace, "head_frames: null pointer");
  TMR_CHECK_ARG(pad_mode != TMR_PAD_ZERO || frame2vstart, "head_frames: TMR_PAD_ZERO needs frame2vstart");
  TMR_CHECK_ARG(aligned16(feats) && aligned16(bank) && aligned16(workspace), "head_frames: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv(workspace, workspace_bytes);
  LstmWs lw; HeadWs hw;
  float* win = nullptr;
  const bool ok = carve_lstm(cv, n_feat_frames, B, lw) && (win = cv.take((size_t)B * L * kD)) && carve_head(cv, B, L, hw);
  TMR_CHECK_ARG(ok, "head_frames: workspace too small (%zu < %zu)", workspace_bytes,
                tmr_head_frames_workspace_bytes(n_feat_frames, B, L, D));
  float* St = St_out ? St_out : hw.St;
  const half_t* st16 = nullptr;
  TMR_TRY(lstm_impl((const float*)lstm_packed, feats, n_feat_frames, starts, B, seq, St, lw, math_mode, st, frame0, false, &st16));
  TMR_TRY(launch_gather(bank, n_rows, frame2row, frame2vstart, n_frames_total, starts, B, L, pad_mode, win,
                        nullptr, st));
  return head_tail(timeconv_packed, nlblock_packed, classifier_packed, St, win, B, L, C, logits, pred, score, hw,
                   math_mode, st, st16);
}

size_t tmr_bankconv_workspace_bytes(int64_t pb_rows, int D) { return fbytes((size_t)((pb_rows > 0 ? pb_rows : 1) + 8) * D); }
int tmr_bankconv_fwd(const void* timeconv_packed, const float* bank, int64_t n_rows, int64_t row_base,
                     int64_t pb_rows, int D, void* pb, void* workspace, size_t workspace_bytes, void* stream) {
  TMR_TRY(check_dims(D));
  TMR_TRY(check_mode(TMR_MATH_F16));
  TMR_CHECK_ARG(pb_rows >= 0 && row_base >= 0 && row_base + pb_rows <= n_rows, "bankconv: row range outside the bank");
  if (pb_rows == 0) return TMR_OK;
  TMR_CHECK_ARG(timeconv_packed && bank && pb && workspace, "bankconv: null pointer");
  TMR_CHECK_ARG(aligned16(bank) && aligned16(pb) && aligned16(workspace), "bankconv: pointers must be 16-byte aligned");
  Carver cv(workspace, workspace_bytes);
  float* bank_r = cv.take((size_t)(pb_rows + 8) * kD);
  TMR_CHECK_ARG(bank_r, "bankconv: workspace too small");
  const int64_t r_lo = row_base - 3 > 0 ? row_base - 3 : 0;
  const int64_t r_hi = row_base + pb_rows + 4 < n_rows ? row_base + pb_rows + 4 : n_rows;
  cudaStream_t st = (cudaStream_t)stream;
  half_t* bank16 = reinterpret_cast<half_t*>(bank_r);
  TMR_TRY(launch_to_half(bank + r_lo * kD, bank16, (r_hi - r_lo) * kD, st));
  return umma_bankconv((const float*)timeconv_packed, bank, bank16, n_rows, r_lo, r_hi - r_lo, row_base, pb_rows,
                       reinterpret_cast<half_t*>(pb), st);
}

size_t tmr_head_frames_dedup_workspace_bytes(int64_t n_feat_frames, int B, int n_irregular, int n_irregular_rows,
                                             int64_t pb_rows, int L, int D) {
  const size_t b = (size_t)(B > 0 ? B : 1);
  const size_t ni = (size_t)(n_irregular > 0 ? n_irregular : 1);
  const size_t pr = (size_t)(pb_rows > 0 ? pb_rows : 1);
  const size_t nr = (size_t)(n_irregular_rows > 0 ? n_irregular_rows : 0);
  const size_t irr = (n_irregular > 0 && nr > 0) ? fbytes(nr * 15 * D) + fbytes(nr * D) + fbytes(ni * L)
                                                 : 2 * fbytes(ni * L * D);
  return tmr_lstm_workspace_bytes(n_feat_frames, (int)b, D) + fbytes((pr + 8) * D) + fbytes((pr * 7 * D + 1) / 2) +
         fbytes(ni * L * D) + irr + 2 * fbytes(b * D) + tmr_nlblock_workspace_bytes((int)b, D) +
         tmr_classifier_workspace_bytes((int)b, D);
}
int tmr_head_frames_dedup_fwd(const void* lstm_packed, const void* timeconv_packed,
                              const void* nlblock_packed, const void* classifier_packed,
                              const void* feats, int feats_f16, int64_t n_feat_frames, int64_t frame0, const float* bank,
                              int64_t n_rows, const int32_t* frame2row, const int32_t* frame2vstart,
                              int64_t n_frames_total, const int64_t* starts, int B, const int32_t* src_idx,
                              const int64_t* irregular_starts, int n_irregular, const int32_t* irregular_rows,
                              int n_irregular_rows, int64_t pb_row_base,
                              int64_t pb_rows, int seq, int L, int F, int D, int C, int pad_mode,
                              float* logits, int64_t* pred, float* score, float* St_out, void* workspace,
                              size_t workspace_bytes, void* stream) {
  const int mode = TMR_MATH_F16;
  TMR_TRY(check_dims(D, F));
  TMR_TRY(check_mode(mode));
  TMR_CHECK_ARG(B >= 0 && seq >= 1 && n_feat_frames >= 0 && frame0 >= 0 && n_irregular >= 0 && pb_rows >= 0,
                "head_frames_dedup: bad sizes");
  TMR_CHECK_ARG(L >= 6, "head_frames_dedup: L=%d < 6 has no interior slots; use tmr_head_frames_fwd", L);
  TMR_CHECK_ARG(C >= 1 && C <= ClassifierPacked::kMaxC, "head_frames_dedup: C=%d out of range", C);
  TMR_CHECK_ARG(pad_mode == TMR_PAD_REPEAT || pad_mode == TMR_PAD_ZERO, "head_frames_dedup: bad pad_mode");
  if (B == 0) return TMR_OK;
  TMR_CHECK_ARG(lstm_packed && timeconv_packed && nlblock_packed && classifier_packed && feats && bank && frame2row &&
                starts && src_idx && logits && workspace, "head_frames_dedup: null pointer");
  TMR_CHECK_ARG(n_irregular == 0 || irregular_starts, "head_frames_dedup: irregular_starts is null");
  TMR_CHECK_ARG(n_irregular_rows >= 0 && (n_irregular_rows == 0 || irregular_rows), "head_frames_dedup: irregular_rows is null");
  const bool irr_rows = n_irregular > 0 && n_irregular_rows > 0;   // assemble from per-row tap products
  TMR_CHECK_ARG(pad_mode != TMR_PAD_ZERO || frame2vstart, "head_frames_dedup: TMR_PAD_ZERO needs frame2vstart");
  TMR_CHECK_ARG(pb_row_base >= 0 && pb_row_base + pb_rows <= n_rows, "head_frames_dedup: PB row range outside the bank");
  TMR_CHECK_ARG(aligned16(feats) && aligned16(bank) && aligned16(workspace), "head_frames_dedup: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv(workspace, workspace_bytes);
  LstmWs lw;
  bool ok = carve_lstm(cv, n_feat_frames, B, lw);
  const int64_t r_lo = pb_row_base - 3 > 0 ? pb_row_base - 3 : 0;                       // fp16 bank slice
  const int64_t r_hi = pb_row_base + pb_rows + 4 < n_rows ? pb_row_base + pb_rows + 4 : n_rows;
  float* bank_r = cv.take((size_t)(pb_rows + 8) * kD);
  half_t* pb = reinterpret_cast<half_t*>(cv.take(((size_t)(pb_rows > 0 ? pb_rows : 1) * 7 * kD + 1) / 2));   // fp16
  const size_t ni = (size_t)(n_irregular > 0 ? n_irregular : 1);
  float* lt_i = cv.take(ni * L * kD);
  float *win_i = nullptr, *xr_i = nullptr, *q_i = nullptr, *xc_i = nullptr;
  int32_t* wrows_i = nullptr;
  if (irr_rows) {
    q_i = cv.take((size_t)n_irregular_rows * 15 * kD); xc_i = cv.take((size_t)n_irregular_rows * kD);
    wrows_i = reinterpret_cast<int32_t*>(cv.take(ni * L));
    ok = ok && q_i && xc_i && wrows_i;
  } else {
    win_i = cv.take(ni * L * kD); xr_i = cv.take(ni * L * kD);
    ok = ok && win_i && xr_i;
  }
  float* St_ws = cv.take((size_t)B * kD);
  float* y1 = cv.take((size_t)B * kD);
  NLWs nl; nl.w0 = cv.take((size_t)B * kD); nl.w1 = cv.take((size_t)B * kD); nl.s = cv.take((size_t)B * kD);
  ClsWs cls;
  ok = ok && bank_r && pb && lt_i && St_ws && y1 && nl.w0 && nl.w1 && nl.s && carve_cls(cv, B, cls);
  TMR_CHECK_ARG(ok, "head_frames_dedup: workspace too small (%zu < %zu)", workspace_bytes,
                tmr_head_frames_dedup_workspace_bytes(n_feat_frames, B, n_irregular, n_irregular_rows, pb_rows, L, D));
  float* St = St_out ? St_out : St_ws;
  const half_t* st16 = nullptr;
  // fork: the bank-side chain runs on the side stream `sb` beside the LSTM chain on `st` (see side_stream())
  cudaStream_t sb = st;
  SideStream* ss = env_int("TMR_FORK_BANK", 1) ? side_stream() : nullptr;
  if (ss) {
    TMR_CUDA(cudaEventRecord(ss->fork, st));
    TMR_CUDA(cudaStreamWaitEvent(ss->s, ss->fork, 0));
    sb = ss->s;
  }
  TMR_TRY(lstm_impl((const float*)lstm_packed, (const float*)feats, n_feat_frames, starts, B, seq, St, lw, mode, st, frame0,
                    feats_f16 != 0, &st16));
  {
  cudaStream_t st = sb;            // everything in this block is bank-side work
  if (pb_rows > 0) {
    // fp16 copy of the bank rows the convolutions touch, then one pass of tap products per row
    half_t* bank16 = reinterpret_cast<half_t*>(bank_r);
    TMR_TRY(launch_to_half(bank + r_lo * kD, bank16, (r_hi - r_lo) * kD, st));
    TMR_TRY(umma_bankconv((const float*)timeconv_packed, bank, bank16, n_rows, r_lo, r_hi - r_lo, pb_row_base,
                          pb_rows, pb, st));
  }
  if (irr_rows) {
    // clips whose window crosses a video start: tap products once per DISTINCT row their windows touch,
    // then each (clip, slot) sums the taps of its own neighbours (no per-clip GEMM)
    const float* tp = (const float*)timeconv_packed;
    TMR_TRY(launch_gather(bank, n_rows, frame2row, frame2vstart, n_frames_total, irregular_starts, n_irregular, L,
                          pad_mode, nullptr, wrows_i, st));
    half_t* xc16 = reinterpret_cast<half_t*>(xc_i);
    TMR_TRY(launch_compact_rows_half(bank, irregular_rows, n_irregular_rows, xc16, st));
    TMR_TRY(umma_bankconv_raw(tp, xc16, n_irregular_rows, q_i, st));
    TMR_TRY(launch_irr_assemble(q_i, irregular_rows, n_irregular_rows, wrows_i, bank, tp + TimeConvPacked::b3_off,
                                tp + TimeConvPacked::b5_off, tp + TimeConvPacked::b7_off, n_irregular, L, lt_i, st));
  } else if (n_irregular > 0) {      // no row list given: per-clip gather + TimeConv
    TMR_TRY(launch_gather(bank, n_rows, frame2row, frame2vstart, n_frames_total, irregular_starts, n_irregular, L,
                          pad_mode, win_i, nullptr, st));
    TMR_TRY(timeconv_impl((const float*)timeconv_packed, win_i, n_irregular, L, lt_i, xr_i, mode, st));
  }
  }
  if (ss) {                        // join: the relation block needs St and PB / the irregular windows
    TMR_CUDA(cudaEventRecord(ss->join, sb));
    TMR_CUDA(cudaStreamWaitEvent(st, ss->join, 0));
  }
  PbSrc pbs{pb, lt_i, src_idx};
  TMR_TRY(nlblock_impl((const float*)nlblock_packed, St, nullptr, B, L, y1, nl, mode, st, &pbs, true, st16));
  return classifier_impl((const float*)classifier_packed, St, y1, B, C, logits, pred, score, cls, mode, st, true);
}

int tmr_linear_fwd(const void* a, const void* w, const float* bias, int64_t M, int N, int K,
                   float* out, int relu, int math_mode, void* stream) {
  // Exposed for tests of the GEMM engines.  TMR_MATH_F16: `a` [M,K] and `w` [N,K] point to fp16 data (the
  // caller converts; this entry point has no workspace), fp32 accumulate, fp32 bias / out.
  TMR_TRY(check_mode(math_mode));
  TMR_CHECK_ARG(M >= 0 && N >= 1 && K >= 1, "linear: bad sizes");
  if (M == 0) return TMR_OK;
  TMR_CHECK_ARG(a && w && out, "linear: null pointer");
  TMR_CHECK_ARG(aligned16(a) && aligned16(w) && aligned16(out), "linear: pointers must be 16-byte aligned");
  LinearArgs g;
  g.lda = K; g.ldw = K; g.bias = bias; g.out = out; g.ldo = N; g.M = M; g.N = N; g.K = K;
  g.relu = relu;
  if (math_mode == TMR_MATH_F16) { g.a16 = reinterpret_cast<const half_t*>(a); g.w16 = reinterpret_cast<const half_t*>(w); }
  else { g.a = reinterpret_cast<const float*>(a); g.w = reinterpret_cast<const float*>(w); }
  return do_linear(g, math_mode, (cudaStream_t)stream);
}

}  // extern "C"
