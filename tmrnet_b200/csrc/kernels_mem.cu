// HBM-bound kernels of the head: window gather, attention over the L memory slots (warp-shuffle
// online softmax), LayerNorm+ReLU, the 7-way FC + softmax score + argmax, the LSTM step-0 cell and
// the one-off weight repacks.  All use 128-bit accesses with a warp per 512-float row.
#include "tmr_internal.h"
#include "row_ops.cuh"

namespace tmr {

// -------------------------------------------------------------------------------------------
// a3: window gather.  One warp per (clip, slot) row of 512 floats: 4 x 128-bit loads per lane.
// row(b,k) = frame2row[starts[b]-k-1], 0 for negative keys (repeat-fill/leak semantics are baked
// into frame2row, see tmr_build_frame2row); TMR_PAD_ZERO zeroes slots before the clip's video.
// -------------------------------------------------------------------------------------------
constexpr int kGatherWarps = 8;

// Memory safety: a start that is not a valid clip start of the table (outside [0, n_frames), or a frame whose
// table entry is only the "row the reference walk is still remembering": f2r[s] == f2r[s+1]) makes the
// reference raise KeyError (TRAIN:310).  Here such a clip gets an all-zero window, row indices -2, and bit 0 of
// *status; a table entry outside [0, n_rows) sets bit 1.  Nothing is ever read out of bounds.
__global__ void __launch_bounds__(kGatherWarps * 32)
gather_kernel(const float* __restrict__ bank, int64_t n_rows, const int32_t* __restrict__ f2r,
              const int32_t* __restrict__ f2v, int64_t n_frames, const int64_t* __restrict__ starts, int64_t total,
              int L, int pad_mode, float* __restrict__ out, int32_t* __restrict__ rows_out,
              int32_t* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int64_t wid0 = (int64_t)blockIdx.x * kGatherWarps + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * kGatherWarps;
  for (int64_t w = wid0; w < total; w += stride) {
    const int64_t b = w / L;
    const int k = (int)(w - b * L);
    const int64_t s = starts[b];
    const int64_t key = s - k - 1;
    int64_t row = -2;
    int bad = 0;
    if (s < 0 || s >= n_frames) {
      bad = 1;
    } else {
      const int32_t own = f2r[s];
      if (own < 0 || (s + 1 < n_frames && f2r[s + 1] == own)) bad = 1;
    }
    if (!bad) {
      if (pad_mode == TMR_PAD_ZERO) row = (key >= (int64_t)f2v[s]) ? (int64_t)f2r[key] : -1;
      else row = (key >= 0) ? (int64_t)f2r[key] : 0;
      if (row >= n_rows || row < -1) { bad = 2; row = -2; }
    }
    if (bad && status && lane == 0) atomicOr(status, bad);
    if (out) {                                 // out == nullptr: window ROW INDICES only
      float4 v[4];
      if (row >= 0) {
        const float4* src = reinterpret_cast<const float4*>(bank + row * kD);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = ldg_nc(src + i * 32 + lane);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float4* dst = reinterpret_cast<float4*>(out + w * kD);
#pragma unroll
      for (int i = 0; i < 4; ++i) stg_na(dst + i * 32 + lane, v[i]);
    }
    if (rows_out && lane == 0) rows_out[w] = (int32_t)row;
  }
}

int launch_gather(const float* bank, int64_t n_rows, const int32_t* f2r, const int32_t* f2v,
                  int64_t n_frames, const int64_t* starts, int B, int L, int pad_mode, float* out,
                  int32_t* rows_out, cudaStream_t st, int32_t* status) {
  const int64_t total = (int64_t)B * L;
  if (total == 0) return TMR_OK;
  // grid: enough warps to cover the rows, capped at a multiple of the 148 SMs (8 CTAs each)
  int64_t blocks = (total + kGatherWarps - 1) / kGatherWarps;
  const int64_t cap = 148 * 8 * 4;
  if (blocks > cap) blocks = cap;
  gather_kernel<<<(unsigned)blocks, kGatherWarps * 32, 0, st>>>(bank, n_rows, f2r, f2v, n_frames, starts, total,
                                                               L, pad_mode, out, rows_out, status);
  TMR_LAUNCH_CHECK("gather_kernel");
  return TMR_OK;
}

// -------------------------------------------------------------------------------------------
// Irregular clips of the bank-level path (windows that repeat-fill / leak across a video start): their
// TimeConv is assembled from UNSHIFTED per-row tap products instead of a per-clip implicit GEMM.
//   compact_rows_half : xc[i] = fp16(bank[rows[i]])              (the distinct rows those windows touch)
//   umma_bankconv_raw : q[i][tap] = W_tap . xc[i]                (15 taps x 512 per row, once per row)
//   irr_assemble      : Lt[b,k] = max(x[k], k>0 ? x[k-1] : 0, conv3, conv5, conv7),
//                       conv_K[k] = b_K + sum_{t, 0<=k+t<L} q[row(slot k+t)][tap(K,t)]      (NLB:55-68)
// A clip's 236 MFLOP become ~34 KB of L2 reads per slot; only fp32 summation order changes.
// -------------------------------------------------------------------------------------------
__global__ void compact_rows_half_kernel(const float* __restrict__ bank, const int32_t* __restrict__ rows, int n,
                                         half_t* __restrict__ out) {
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // float4 index
  if (i4 >= (int64_t)n * (kD / 4)) return;
  const int i = (int)(i4 / (kD / 4));
  const int c4 = (int)(i4 - (int64_t)i * (kD / 4));
  float4 v = __ldg(reinterpret_cast<const float4*>(bank + (int64_t)rows[i] * kD) + c4);
  reinterpret_cast<uint2*>(out)[i4] = pack_h4(v);
}
int launch_compact_rows_half(const float* bank, const int32_t* rows, int n, half_t* out, cudaStream_t st) {
  if (n <= 0) return TMR_OK;
  const int64_t total = (int64_t)n * (kD / 4);
  compact_rows_half_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(bank, rows, n, out);
  TMR_LAUNCH_CHECK("compact_rows_half_kernel");
  return TMR_OK;
}

constexpr int kAsmWarps = 8;
__global__ void __launch_bounds__(kAsmWarps * 32)
irr_assemble_kernel(const float* __restrict__ q, const int32_t* __restrict__ crows, int n_c,
                    const int32_t* __restrict__ wrows, const float* __restrict__ bank,
                    const float* __restrict__ b3, const float* __restrict__ b5, const float* __restrict__ b7,
                    int64_t total, int L, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * kAsmWarps + (threadIdx.x >> 5);    // (clip, slot)
  if (w >= total) return;
  const int64_t b = w / L;
  const int k = (int)(w - b * L);
  // lanes 0..6 resolve the window rows of slots k-3 .. k+3 to positions in the compact row list
  int row_t = -1, ci = -1;
  if (lane < 7) {
    const int kk = k + lane - 3;
    if (kk >= 0 && kk < L) {
      row_t = wrows[b * L + kk];
      if (row_t >= 0) {
        int lo = 0, hi = n_c - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (crows[mid] < row_t) lo = mid + 1; else hi = mid; }
        ci = (crows[lo] == row_t) ? lo : -1;
      }
    }
  }
  float4 a7[4], a5[4], a3[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a7[i] = __ldg(reinterpret_cast<const float4*>(b7) + i * 32 + lane);
    a5[i] = __ldg(reinterpret_cast<const float4*>(b5) + i * 32 + lane);
    a3[i] = __ldg(reinterpret_cast<const float4*>(b3) + i * 32 + lane);
  }
  auto add4 = [](float4& a, const float4 v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; };
#pragma unroll
  for (int t = -3; t <= 3; ++t) {
    const int c = __shfl_sync(0xffffffffu, ci, t + 3);
    if (c < 0) continue;                                  // slot outside the window, or a zero-padded slot
    const float4* base = reinterpret_cast<const float4*>(q + (int64_t)c * (15 * kD));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      add4(a7[i], __ldg(base + (t + 3) * (kD / 4) + i * 32 + lane));
      if (t >= -2 && t <= 2) add4(a5[i], __ldg(base + (7 + t + 2) * (kD / 4) + i * 32 + lane));
      if (t >= -1 && t <= 1) add4(a3[i], __ldg(base + (12 + t + 1) * (kD / 4) + i * 32 + lane));
    }
  }
  const int row_k = __shfl_sync(0xffffffffu, row_t, 3);
  const int row_p = __shfl_sync(0xffffffffu, row_t, 2);   // slot k-1 (-1 when k == 0: the pool sees the zero pad)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 x0 = row_k >= 0 ? __ldg(reinterpret_cast<const float4*>(bank + (int64_t)row_k * kD) + i * 32 + lane) : z;
    const float4 x1 = row_p >= 0 ? __ldg(reinterpret_cast<const float4*>(bank + (int64_t)row_p * kD) + i * 32 + lane) : z;
    float4 o;
    o.x = fmaxf(fmaxf(fmaxf(a7[i].x, a5[i].x), a3[i].x), fmaxf(x0.x, x1.x));
    o.y = fmaxf(fmaxf(fmaxf(a7[i].y, a5[i].y), a3[i].y), fmaxf(x0.y, x1.y));
    o.z = fmaxf(fmaxf(fmaxf(a7[i].z, a5[i].z), a3[i].z), fmaxf(x0.z, x1.z));
    o.w = fmaxf(fmaxf(fmaxf(a7[i].w, a5[i].w), a3[i].w), fmaxf(x0.w, x1.w));
    reinterpret_cast<float4*>(out + w * kD)[i * 32 + lane] = o;
  }
}
int launch_irr_assemble(const float* q, const int32_t* crows, int n_c, const int32_t* wrows, const float* bank,
                        const float* b3, const float* b5, const float* b7, int n_clips, int L, float* out,
                        cudaStream_t st) {
  const int64_t total = (int64_t)n_clips * L;
  if (total == 0) return TMR_OK;
  irr_assemble_kernel<<<(unsigned)((total + kAsmWarps - 1) / kAsmWarps), kAsmWarps * 32, 0, st>>>(
      q, crows, n_c, wrows, bank, b3, b5, b7, total, L, out);
  TMR_LAUNCH_CHECK("irr_assemble_kernel");
  return TMR_OK;
}

// -------------------------------------------------------------------------------------------
// LSTM step 0 from zero state: gates = xp row (bias already folded), c = sig(i) tanh(g).
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }
__device__ __forceinline__ float ex2f_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// One thread per (clip, 4 units): 64 contiguous bytes of the projected row in, 128-bit c / h out.
// FAST (tensor-core mode): MUFU ex2/rcp gates like the recurrent-step epilogue; fp32 mode keeps expf/tanhf.
template <bool FAST>
__global__ void lstm_cell0_kernel(const float* __restrict__ xp, const int64_t* __restrict__ starts,
                                  int seq, float* __restrict__ h, half_t* __restrict__ h16, float* __restrict__ c,
                                  int64_t total4) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (clip, unit quad)
  if (idx >= total4) return;
  const int64_t m = idx / (kD / 4);
  const int u4 = (int)(idx - m * (kD / 4)) * 4;
  const int64_t xr = starts ? starts[m] : m * seq;
  const float4* src = reinterpret_cast<const float4*>(xp + xr * (4 * kD) + u4 * 4);
  float cn[4], hn[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 p = __ldg(src + k);                   // (i, f, g, o) of unit u4 + k; f * 0 drops out
    if (FAST) {
      const float a = ex2f_approx(fminf(-1.4426950408889634f * p.x, 40.f));
      const float d = ex2f_approx(fminf(-2.8853900817779268f * p.z, 40.f));
      const float e = ex2f_approx(fminf(-1.4426950408889634f * p.w, 40.f));
      cn[k] = (1.f - d) * rcpf_approx((1.f + a) * (1.f + d));          // sigmoid(i) tanh(g)
      const float f2 = ex2f_approx(fminf(-2.8853900817779268f * cn[k], 40.f));
      hn[k] = (1.f - f2) * rcpf_approx((1.f + e) * (1.f + f2));        // sigmoid(o) tanh(c)
    } else {
      cn[k] = sigmoidf_(p.x) * tanhf(p.z);
      hn[k] = sigmoidf_(p.w) * tanhf(cn[k]);
    }
  }
  *reinterpret_cast<float4*>(c + m * kD + u4) = make_float4(cn[0], cn[1], cn[2], cn[3]);
  if (h16) *reinterpret_cast<uint2*>(h16 + m * kD + u4) = pack_h4(hn[0], hn[1], hn[2], hn[3]);
  else *reinterpret_cast<float4*>(h + m * kD + u4) = make_float4(hn[0], hn[1], hn[2], hn[3]);
}

int launch_lstm_cell0(const float* xp, const int64_t* starts, int seq, float* h, half_t* h16, float* c, int B,
                      cudaStream_t st, bool fast_math) {
  const int64_t total4 = (int64_t)B * (kD / 4);
  if (total4 == 0) return TMR_OK;
  const unsigned blocks = (unsigned)((total4 + 255) / 256);
  if (fast_math) lstm_cell0_kernel<true><<<blocks, 256, 0, st>>>(xp, starts, seq, h, h16, c, total4);
  else lstm_cell0_kernel<false><<<blocks, 256, 0, st>>>(xp, starts, seq, h, h16, c, total4);
  TMR_LAUNCH_CHECK("lstm_cell0_kernel");
  return TMR_OK;
}

// -------------------------------------------------------------------------------------------
// Attention over the memory slots (NLB:30-34 with phi/g folded):
//   s_k = scale * (u . Lt_k),  p = softmax_k(s),  a = sum_k p_k Lt_k.
// One warp per clip; each lane owns 16 channels (4 float4); slots are streamed once in chunks of
// KB rows with an online (running-max) softmax, dots reduced with warp shuffles.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttnWarps * 32)
attention_kernel(const float* __restrict__ u, const float* __restrict__ Lt, int B, int L, float scale,
                 void* __restrict__ a, int half_out) {
  const int b = blockIdx.x * kAttnWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const float4* base = reinterpret_cast<const float4*>(Lt + (int64_t)b * L * kD);
  attention_body(u, b, L, scale, a, half_out, [&](int k) { return base + (int64_t)k * (kD / 4); });
}

// 16 consecutive fp16 (one 32-byte sector) -> 8 packed fp32 pairs
__device__ __forceinline__ void ldg_h16(const half_t* p, uint64_t (&x2)[8]) {
  uint32_t v[8];
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(v[i]));
    x2[i] = pk2(lo, hi);
  }
}

// Attention over the bank-level TimeConv output PB[row][7][512] (fp16, umma_bankconv.cu) for a regular clip: the
// window is the contiguous run of bank rows whose slot 0 is PB row s; slot k reads PB row s-k in the variant its
// distance to the window edges selects.  A lane owns 16 CONSECUTIVE channels (one 32-byte sector of an fp16 row, one
// 256-bit load per slot; loads allocate in L1, where the four warps of a CTA - four consecutive clips - share the
// interior rows of their windows).
__device__ __forceinline__ void attention_body_pb16(const float* __restrict__ u, const half_t* __restrict__ pb, int b, int s,
                                                    int L, float scale, void* __restrict__ a, int half_out) {
  const int lane = threadIdx.x & 31;
  uint64_t u2[8], acc2[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(u + (int64_t)b * kD + lane * 16) + i);
    u2[2 * i] = pk2(q.x, q.y); u2[2 * i + 1] = pk2(q.z, q.w);
    acc2[2 * i] = acc2[2 * i + 1] = pk2(0.f, 0.f);
  }
  float run_max = -INFINITY, run_sum = 0.f;
  for (int k0 = 0; k0 < L; k0 += KB) {
    uint64_t x2[KB][8];
    float d[KB];
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      const int k = min(k0 + kk, L - 1);                 // slots past L re-read the last one and are masked below
      const int v = (k <= 2) ? k + 1 : ((L - 1 - k <= 2) ? 4 + (L - 1 - k) : 0);
      ldg_h16(pb + ((int64_t)(s - k) * 7 + v) * kD + lane * 16, x2[kk]);
    }
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      uint64_t p2 = fmul2(x2[kk][0], u2[0]);
#pragma unroll
      for (int i = 1; i < 8; ++i) p2 = ffma2(x2[kk][i], u2[i], p2);
      float pe, po;
      upk2(p2, pe, po);
      d[kk] = pe + po;
    }
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) d[kk] = warp_sum(d[kk]);
    float cmax = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      d[kk] = (k0 + kk < L) ? d[kk] * scale : -INFINITY;
      cmax = fmaxf(cmax, d[kk]);
    }
    const float new_max = fmaxf(run_max, cmax);
    const float corr = expf(run_max - new_max);          // 0 on the first chunk (run_max = -inf)
    run_sum *= corr;
    const uint64_t corr2 = pk2(corr, corr);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc2[i] = fmul2(acc2[i], corr2);
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      const float p = expf(d[kk] - new_max);             // exp(-inf) = 0 for masked slots
      run_sum += p;
      const uint64_t pp = pk2(p, p);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc2[i] = ffma2(pp, x2[kk][i], acc2[i]);
    }
    run_max = new_max;
  }
  const float inv = 1.f / run_sum;
  float o[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) { upk2(acc2[i], o[2 * i], o[2 * i + 1]); o[2 * i] *= inv; o[2 * i + 1] *= inv; }
  if (half_out) {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<half_t*>(a) + (int64_t)b * kD + lane * 16);
    dst[0] = make_uint4(pack_h2(o[0], o[1]), pack_h2(o[2], o[3]), pack_h2(o[4], o[5]), pack_h2(o[6], o[7]));
    dst[1] = make_uint4(pack_h2(o[8], o[9]), pack_h2(o[10], o[11]), pack_h2(o[12], o[13]), pack_h2(o[14], o[15]));
  } else {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a) + (int64_t)b * kD + lane * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
  }
}

// src[b] >= 0: a regular clip, window read from PB (above).  src[b] < 0: an irregular clip (window crosses a video
// start) whose TimeConv output lt_irr[-1-src[b]] (fp32) was assembled per clip.
__global__ void __launch_bounds__(kAttnWarps * 32)
attention_pb_kernel(const float* __restrict__ u, const half_t* __restrict__ pb, const float* __restrict__ lt_irr,
                    const int32_t* __restrict__ src, int B, int L, float scale, void* __restrict__ a,
                    int half_out) {
  const int b = blockIdx.x * kAttnWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const int s = src[b];
  if (s >= 0) {
    attention_body_pb16(u, pb, b, s, L, scale, a, half_out);
  } else {
    const float4* base = reinterpret_cast<const float4*>(lt_irr + (int64_t)(-1 - s) * L * kD);
    attention_body(u, b, L, scale, a, half_out, [&](int k) { return base + (int64_t)k * (kD / 4); });
  }
}

int launch_attention(const float* u, const float* Lt, int B, int L, void* a, int half_out, cudaStream_t st) {
  if (B == 0) return TMR_OK;
  const float scale = (float)0.044194173824159216;   // (1/512)**0.5 as python computes it (NLB:31)
  attention_kernel<<<(B + kAttnWarps - 1) / kAttnWarps, kAttnWarps * 32, 0, st>>>(u, Lt, B, L, scale, a, half_out);
  TMR_LAUNCH_CHECK("attention_kernel");
  return TMR_OK;
}

int launch_attention_pb(const float* u, const half_t* pb, const float* lt_irr, const int32_t* src, int B, int L,
                        void* a, int half_out, cudaStream_t st) {
  if (B == 0) return TMR_OK;
  const float scale = (float)0.044194173824159216;
  attention_pb_kernel<<<(B + kAttnWarps - 1) / kAttnWarps, kAttnWarps * 32, 0, st>>>(u, pb, lt_irr, src, B, L, scale, a, half_out);
  TMR_LAUNCH_CHECK("attention_pb_kernel");
  return TMR_OK;
}

// -------------------------------------------------------------------------------------------
// LayerNorm([1,512]) (biased variance, eps 1e-5) + ReLU (NLB:35-36).  One warp per row.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
layernorm_relu_kernel(const float* __restrict__ v, const float* __restrict__ w,
                      const float* __restrict__ bsh, int B, void* __restrict__ y, int half_out) {
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  layernorm_relu_row<false>(v, w, bsh, b, y, half_out);
}

int launch_layernorm_relu(const float* v, const float* w, const float* b, int B, void* y, int half_out,
                          cudaStream_t st) {
  if (B == 0) return TMR_OK;
  layernorm_relu_kernel<<<(B + 3) / 4, 128, 0, st>>>(v, w, b, B, y, half_out);
  TMR_LAUNCH_CHECK("layernorm_relu_kernel");
  return TMR_OK;
}

// -------------------------------------------------------------------------------------------
// fc_c (512 -> C) + Softmax + torch.max (TRAIN:252, EVAL:491-493).  One warp per clip.
// pred = first index of the maximum logit (ties -> lowest index); score = 1 / sum exp(l - lmax).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
fc_argmax_kernel(const float* __restrict__ z, const float* __restrict__ wc, const float* __restrict__ bc,
                 int B, int C, float* __restrict__ logits, int64_t* __restrict__ pred,
                 float* __restrict__ score) {
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  fc_argmax_row<false>(z, wc, bc, b, C, logits, pred, score);
}

// Large batches, C <= 8: re-loading the C weight rows for every clip made the kernel L1-wavefront-bound (28 LDG.128 of
// weights per clip against 4 of activations: 69 us per 83 k clips, 2.6x the HBM time of z; this form: 59 us - with 8
// warps per SM it is bound by the latency of its own reduction chains; computing max / denominator locally on every
// lane after the butterflies was slower, 81 us).  Here a warp keeps its
// lanes' share of Wc in registers (C x 16 floats) and walks clips grid-stride, the next clip's z row in flight while
// the current one is reduced.  Same products, same order, same reductions as fc_argmax_row: bit-identical.
template <int CMAX>
__global__ void __launch_bounds__(256)
fc_argmax_regs_kernel(const float* __restrict__ z, const float* __restrict__ wc, const float* __restrict__ bc,
                      int B, int C, float* __restrict__ logits, int64_t* __restrict__ pred, float* __restrict__ score) {
  const int lane = threadIdx.x & 31;
  const int warp0 = blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = gridDim.x * 8;
  float4 w[CMAX][4];
  float bias[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    const bool on = c < C;
    bias[c] = on ? __ldg(bc + c) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      w[c][i] = on ? __ldg(reinterpret_cast<const float4*>(wc + (int64_t)c * kD) + i * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // two rows in flight per warp behind the one being reduced (8 warps per SM x 2 KB x 2: enough bytes for HBM)
  auto load_row = [&](int b, float4 (&r)[4]) {
    if (b < B) {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = ldg_nc(reinterpret_cast<const float4*>(z + (int64_t)b * kD) + i * 32 + lane);
    }
  };
  float4 x[4], x1[4], x2[4];
  load_row(warp0, x1);
  load_row(warp0 + nwarps, x2);
  for (int b = warp0; b < B; b += nwarps) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = x1[i]; x1[i] = x2[i]; }
    load_row(b + 2 * nwarps, x2);
    // all CMAX dot products (rows past C are zero) and their butterflies side by side: no branch between them, so the
    // eight 5-level reductions overlap instead of running back to back (2 warps per scheduler cannot hide them)
    float pc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      float p = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        p = fmaf(x[i].x, w[c][i].x, p); p = fmaf(x[i].y, w[c][i].y, p); p = fmaf(x[i].z, w[c][i].z, p); p = fmaf(x[i].w, w[c][i].w, p);
      }
      pc[c] = p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c) pc[c] += __shfl_xor_sync(0xffffffffu, pc[c], o);
    }
    float mine = -INFINITY;                    // lane c keeps logit c
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (lane == c && c < C) mine = pc[c] + bias[c];
    if (lane < C) logits[(int64_t)b * C + lane] = mine;
    const float mx = warp_max(mine);
    const unsigned hit = __ballot_sync(0xffffffffu, lane < C && mine == mx);
    const float e = (lane < C) ? expf(mine - mx) : 0.f;
    const float den = warp_sum(e);
    if (lane == 0) {
      if (pred) pred[b] = (int64_t)(__ffs(hit) - 1);
      if (score) score[b] = 1.f / den;
    }
  }
}

int launch_fc_argmax(const float* z, const float* wc, const float* bc, int B, int C, float* logits,
                     int64_t* pred, float* score, cudaStream_t st) {
  if (B == 0) return TMR_OK;
  if (C <= 8 && B >= 4096) {
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    fc_argmax_regs_kernel<8><<<sms, 256, 0, st>>>(z, wc, bc, B, C, logits, pred, score);     // 8 warps per SM
    TMR_LAUNCH_CHECK("fc_argmax_regs_kernel");
    return TMR_OK;
  }
  fc_argmax_kernel<<<(B + 3) / 4, 128, 0, st>>>(z, wc, bc, B, C, logits, pred, score);
  TMR_LAUNCH_CHECK("fc_argmax_kernel");
  return TMR_OK;
}

// -------------------------------------------------------------------------------------------
// weight repacks (run once per weight update)
// -------------------------------------------------------------------------------------------
__global__ void pack_conv_kernel(const float* __restrict__ w, int taps, float* __restrict__ dst) {
  // dst[o][tap][c] = w[o][c][tap]
  const int64_t total = (int64_t)kD * taps * kD;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % kD);
    const int tap = (int)((i / kD) % taps);
    const int o = (int)(i / ((int64_t)kD * taps));
    dst[i] = w[((int64_t)o * kD + c) * taps + tap];
  }
}
__global__ void copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  // dst[j][i] = src[i][j], n x n
  __shared__ float t[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = src[(int64_t)(by + r) * n + bx + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) dst[(int64_t)(bx + r) * n + by + threadIdx.x] = t[threadIdx.x][r];
}
// dst row (unit*4+gate) = src row (gate*D+unit); cols wide
__global__ void interleave_gates_kernel(const float* __restrict__ src, float* __restrict__ dst, int cols) {
  const int64_t total = (int64_t)4 * kD * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % cols);
    const int r = (int)(i / cols);
    const int unit = r >> 2, gate = r & 3;
    dst[i] = src[((int64_t)gate * kD + unit) * cols + col];
  }
}
__global__ void interleave_bias_kernel(const float* __restrict__ bih, const float* __restrict__ bhh, float* __restrict__ dst) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= 4 * kD) return;
  const int unit = r >> 2, gate = r & 3;
  dst[r] = bih[gate * kD + unit] + bhh[gate * kD + unit];
}

__global__ void row2clip_kernel(const int64_t* __restrict__ starts, int seq, int B, int64_t n_rows, int64_t row0,
                                int32_t* __restrict__ row2clip) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t r = starts ? starts[b] - row0 : (int64_t)b * seq;
  if (r >= 0 && r < n_rows) row2clip[r] = b;
}
// After the projection whose epilogue ran step 0 for the clips the table names: a clip that LOST its table slot
// (two clips with the same start: the table holds one clip per projected row) still needs c0 / h0.  One thread per
// clip compares row2clip[start] with its own id; losers compute the 512 cells from the projected row with the same
// formula as the fused epilogue (bit-identical c0 / h0).  Distinct starts: one 4-byte load per clip.
__global__ void lstm_cell0_fix_kernel(const float* __restrict__ xp, const int64_t* __restrict__ starts, int B,
                                      int64_t n_rows, int64_t row0, const int32_t* __restrict__ row2clip,
                                      float* __restrict__ c, half_t* __restrict__ h16) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t r = starts[b] - row0;
  const bool in_range = r >= 0 && r < n_rows;
  if (in_range && row2clip[r] == b) return;
  const float4* src = reinterpret_cast<const float4*>(xp + r * (4 * kD));
  for (int u = 0; u < kD; ++u) {
    float cn = 0.f, hn = 0.f;
    if (in_range) {
      const float4 p = __ldg(src + u);
      const float a = ex2f_approx(fminf(-1.4426950408889634f * p.x, 40.f));
      const float d = ex2f_approx(fminf(-2.8853900817779268f * p.z, 40.f));
      const float e = ex2f_approx(fminf(-1.4426950408889634f * p.w, 40.f));
      cn = (1.f - d) * rcpf_approx((1.f + a) * (1.f + d));
      const float f2 = ex2f_approx(fminf(-2.8853900817779268f * cn, 40.f));
      hn = (1.f - f2) * rcpf_approx((1.f + e) * (1.f + f2));
    }
    c[(int64_t)b * kD + u] = cn;
    h16[(int64_t)b * kD + u] = (half_t)(pack_h2(hn, 0.f) & 0xffffu);
  }
}
int launch_lstm_cell0_fix(const float* xp, const int64_t* starts, int B, int64_t n_rows, int64_t row0,
                          const int32_t* row2clip, float* c, half_t* h16, cudaStream_t st) {
  if (B == 0 || !starts) return TMR_OK;           // without starts clip b owns row b * seq: no collisions
  lstm_cell0_fix_kernel<<<(B + 255) / 256, 256, 0, st>>>(xp, starts, B, n_rows, row0, row2clip, c, h16);
  TMR_LAUNCH_CHECK("lstm_cell0_fix_kernel");
  return TMR_OK;
}

int launch_row2clip(const int64_t* starts, int seq, int B, int64_t n_rows, int64_t row0, int32_t* row2clip, cudaStream_t st) {
  if (n_rows == 0) return TMR_OK;
  TMR_CUDA(cudaMemsetAsync(row2clip, 0xff, sizeof(int32_t) * n_rows, st));          // -1 everywhere
  if (B > 0) row2clip_kernel<<<(B + 255) / 256, 256, 0, st>>>(starts, seq, B, n_rows, row0, row2clip);
  TMR_LAUNCH_CHECK("row2clip_kernel");
  return TMR_OK;
}

__global__ void to_half_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = pack_h4(__ldg(src + i));
}
int launch_to_half(const float* src, half_t* dst, int64_t n, cudaStream_t st) {
  if (n == 0) return TMR_OK;
  TMR_CHECK_ARG(n % 4 == 0, "to_half: n must be a multiple of 4");
  int64_t b = (n / 4 + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  to_half_kernel<<<(unsigned)b, 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst), n / 4);
  TMR_LAUNCH_CHECK("to_half_kernel");
  return TMR_OK;
}
// dst[m] = fp16([a[m] || a2[m]]); a2_plus_a: the second half is fp16(a2[m] + a[m]) (the non-local block's
// residual St + W4 r, added here instead of in the GEMM epilogue, which pays ~30 us per batch for it).
// A thread converts 8 consecutive floats (two 128-bit loads, one 128-bit store).
__global__ void half_concat_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ a2, int64_t lda2,
                                   int k_split, int K, int64_t M, half_t* __restrict__ dst, int a2_plus_a) {
  const int64_t n8 = M * (K / 8);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (!a2) {
    // plain conversion (the features of a batch: the one large instance): a thread converts one float4 per
    // piece - fully coalesced 128-bit loads and 64-bit stores - eight independent pieces per iteration, all loads
    // issued before the first store, so enough bytes are in flight per SM
    const int64_t n4 = M * (K / 4);
    int64_t i4 = i;
    for (; i4 + 7 * stride < n4; i4 += 8 * stride) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t ij = i4 + j * stride;
        const int64_t m = ij / (K / 4);
        v[j] = ldg_nc(reinterpret_cast<const float4*>(a + m * lda) + (ij - m * (K / 4)));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) reinterpret_cast<uint2*>(dst)[i4 + j * stride] = pack_h4(v[j]);
    }
    for (; i4 < n4; i4 += stride) {
      const int64_t m = i4 / (K / 4);
      reinterpret_cast<uint2*>(dst)[i4] = pack_h4(ldg_nc(reinterpret_cast<const float4*>(a + m * lda) + (i4 - m * (K / 4))));
    }
    return;
  }
  for (; i < n8; i += stride) {
    const int64_t m = i / (K / 8);
    const int k = (int)(i - m * (K / 8)) * 8;
    const float* src = (k < k_split) ? a + m * lda + k : a2 + m * lda2 + (k - k_split);
    float4 v = __ldg(reinterpret_cast<const float4*>(src));
    float4 w = __ldg(reinterpret_cast<const float4*>(src) + 1);
    if (a2_plus_a && k >= k_split) {
      const float4 e = __ldg(reinterpret_cast<const float4*>(a + m * lda + (k - k_split)));
      const float4 f = __ldg(reinterpret_cast<const float4*>(a + m * lda + (k - k_split)) + 1);
      v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
      w.x += f.x; w.y += f.y; w.z += f.z; w.w += f.w;
    }
    const uint2 lo = pack_h4(v), hi = pack_h4(w);
    *reinterpret_cast<uint4*>(dst + m * K + k) = make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
}
int launch_half_concat(const float* a, int64_t lda, const float* a2, int64_t lda2, int k_split, int K,
                       int64_t M, half_t* dst, cudaStream_t st, bool a2_plus_a) {
  if (M == 0) return TMR_OK;
  TMR_CHECK_ARG(K % 8 == 0 && (!a2 || k_split % 8 == 0), "half_concat: K and the split must be multiples of 8");
  int64_t b = (M * (K / 8) + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  half_concat_kernel<<<(unsigned)b, 256, 0, st>>>(a, lda, a2, lda2, a2 ? k_split : K, K, M, dst,
                                                  (a2 && a2_plus_a && K == 2 * k_split) ? 1 : 0);
  TMR_LAUNCH_CHECK("half_concat_kernel");
  return TMR_OK;
}

static inline unsigned blocks_for(int64_t n) { int64_t b = (n + 255) / 256; return (unsigned)(b > 4096 ? 4096 : b); }

int launch_pack_timeconv(const float* w3, const float* b3, const float* w5, const float* b5,
                         const float* w7, const float* b7, float* packed, cudaStream_t st) {
  pack_conv_kernel<<<blocks_for((int64_t)kD * 3 * kD), 256, 0, st>>>(w3, 3, packed + TimeConvPacked::w3_off);
  pack_conv_kernel<<<blocks_for((int64_t)kD * 5 * kD), 256, 0, st>>>(w5, 5, packed + TimeConvPacked::w5_off);
  pack_conv_kernel<<<blocks_for((int64_t)kD * 7 * kD), 256, 0, st>>>(w7, 7, packed + TimeConvPacked::w7_off);
  copy_kernel<<<2, 256, 0, st>>>(b3, packed + TimeConvPacked::b3_off, kD);
  copy_kernel<<<2, 256, 0, st>>>(b5, packed + TimeConvPacked::b5_off, kD);
  copy_kernel<<<2, 256, 0, st>>>(b7, packed + TimeConvPacked::b7_off, kD);
  TMR_LAUNCH_CHECK("pack_timeconv");
  return launch_to_half(packed, reinterpret_cast<half_t*>(packed + TimeConvPacked::fp32_total), TimeConvPacked::fp32_total, st);
}

int launch_pack_nlblock(const float* w1, const float* b1, const float* w2, const float* w3,
                        const float* b3, const float* w4, const float* b4, const float* lnw,
                        const float* lnb, float* packed, cudaStream_t st) {
  const int64_t n2 = (int64_t)kD * kD;
  copy_kernel<<<blocks_for(n2), 256, 0, st>>>(w1, packed + NLBlockPacked::w1_off, n2);
  transpose_kernel<<<dim3(kD / 32, kD / 32), dim3(32, 8), 0, st>>>(w2, packed + NLBlockPacked::w2t_off, kD);
  copy_kernel<<<blocks_for(n2), 256, 0, st>>>(w3, packed + NLBlockPacked::w3_off, n2);
  copy_kernel<<<blocks_for(n2), 256, 0, st>>>(w4, packed + NLBlockPacked::w4_off, n2);
  copy_kernel<<<2, 256, 0, st>>>(b1, packed + NLBlockPacked::b1_off, kD);
  copy_kernel<<<2, 256, 0, st>>>(b3, packed + NLBlockPacked::b3_off, kD);
  copy_kernel<<<2, 256, 0, st>>>(b4, packed + NLBlockPacked::b4_off, kD);
  copy_kernel<<<2, 256, 0, st>>>(lnw, packed + NLBlockPacked::lnw_off, kD);
  copy_kernel<<<2, 256, 0, st>>>(lnb, packed + NLBlockPacked::lnb_off, kD);
  TMR_LAUNCH_CHECK("pack_nlblock");
  {
    // W21[j][i] = sum_k W2T[j][k] W1[k][i] and bu[j] = sum_k W2T[j][k] b1[k], fp32; W1^T is staged in the
    // mirror half, which the fp16 conversion below overwrites
    float* w1t = packed + NLBlockPacked::fp32_total;
    transpose_kernel<<<dim3(kD / 32, kD / 32), dim3(32, 8), 0, st>>>(w1, w1t, kD);
    LinearArgs g;
    g.a = packed + NLBlockPacked::w2t_off; g.lda = kD; g.w = w1t; g.ldw = kD;
    g.out = packed + NLBlockPacked::w21_off; g.ldo = kD; g.M = kD; g.N = kD; g.K = kD;
    TMR_TRY(simt_linear(g, st));
    LinearArgs b;
    b.a = packed + NLBlockPacked::b1_off; b.lda = kD; b.w = packed + NLBlockPacked::w2t_off; b.ldw = kD;
    b.out = packed + NLBlockPacked::bu_off; b.ldo = kD; b.M = 1; b.N = kD; b.K = kD;
    TMR_TRY(simt_linear(b, st));
  }
  return launch_to_half(packed, reinterpret_cast<half_t*>(packed + NLBlockPacked::fp32_total), NLBlockPacked::fp32_total, st);
}

int launch_pack_lstm(const float* wih, const float* whh, const float* bih, const float* bhh,
                     float* packed, cudaStream_t st) {
  interleave_gates_kernel<<<blocks_for((int64_t)4 * kD * kF), 256, 0, st>>>(wih, packed + LstmPacked::wih_off, kF);
  interleave_gates_kernel<<<blocks_for((int64_t)4 * kD * kD), 256, 0, st>>>(whh, packed + LstmPacked::whh_off, kD);
  interleave_bias_kernel<<<(4 * kD + 255) / 256, 256, 0, st>>>(bih, bhh, packed + LstmPacked::bias_off);
  TMR_LAUNCH_CHECK("pack_lstm");
  return launch_to_half(packed, reinterpret_cast<half_t*>(packed + LstmPacked::fp32_total), LstmPacked::fp32_total, st);
}

int launch_pack_classifier(const float* wh, const float* bh, const float* wc, const float* bc, int C,
                           float* packed, cudaStream_t st) {
  TMR_CUDA(cudaMemsetAsync(packed + ClassifierPacked::wc_off, 0,
                           sizeof(float) * (ClassifierPacked::fp32_total - ClassifierPacked::wc_off), st));
  const int64_t nh = (int64_t)kD * 2 * kD;
  copy_kernel<<<blocks_for(nh), 256, 0, st>>>(wh, packed + ClassifierPacked::wh_off, nh);
  copy_kernel<<<2, 256, 0, st>>>(bh, packed + ClassifierPacked::bh_off, kD);
  copy_kernel<<<blocks_for((int64_t)C * kD), 256, 0, st>>>(wc, packed + ClassifierPacked::wc_off, (int64_t)C * kD);
  copy_kernel<<<1, 256, 0, st>>>(bc, packed + ClassifierPacked::bc_off, C);
  TMR_LAUNCH_CHECK("pack_classifier");
  return launch_to_half(packed, reinterpret_cast<half_t*>(packed + ClassifierPacked::fp32_total), ClassifierPacked::fp32_total, st);
}

}  // namespace tmr
