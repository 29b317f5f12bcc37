// TMR_MATH_FP32 path: the GEMM-shaped stages of the head on CUDA cores (fp32 FFMA, fp32 order close
// to the reference's).  Used for exact-order parity, tiny batches and as the on-device check of the
// tcgen05 path.  See sgemm_simt.cuh for the main loop.
#include "sgemm_simt.cuh"
#include "tmr_internal.h"

namespace tmr {
using namespace simt;

// -------------------------------------------------------------------------------------------
// linear: out = [a|a2] . w^T (+bias) (+residual) (relu)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) linear_kernel(LinearArgs g) {
  __shared__ Smem s;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto afn = [&](int r, int kt) -> const float* {
    const int64_t m = m0 + r;
    if (m >= g.M) return nullptr;
    const int k = kt * BK;
    return (k < g.k_split) ? g.a + m * g.lda + k : g.a2 + m * g.lda2 + (k - g.k_split);
  };
  auto bfn = [&](int r, int kt) -> const float* {
    const int n = n0 + r;
    return (n < g.N) ? g.w + (int64_t)n * g.ldw + kt * BK : nullptr;
  };
  mainloop(acc, afn, bfn, g.K / BK, s);

  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + tile_row(ty, i);
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tile_col(tx, j);
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += __ldg(g.bias + n);
      if (g.residual) v += __ldg(g.residual + m * g.ldr + n);
      if (g.relu) v = fmaxf(v, 0.f);
      g.out[m * g.ldo + n] = v;
    }
  }
}

int simt_linear(const LinearArgs& g, cudaStream_t st) {
  TMR_CHECK_ARG(g.K % BK == 0 && g.K > 0, "linear: K=%d must be a positive multiple of %d", g.K, BK);
  TMR_CHECK_ARG(g.a2 == nullptr ? true : (g.k_split % BK == 0), "linear: k_split must be a multiple of %d", BK);
  if (g.M == 0 || g.N == 0) return TMR_OK;
  LinearArgs h = g;
  if (!h.a2) h.k_split = h.K;
  dim3 grid((h.N + BN - 1) / BN, (unsigned)((h.M + BM - 1) / BM));
  linear_kernel<<<grid, NT, 0, st>>>(h);
  TMR_LAUNCH_CHECK("linear_kernel");
  return TMR_OK;
}

// -------------------------------------------------------------------------------------------
// TimeConv (NLB:43-79): three zero-padded temporal convolutions as implicit GEMMs over the window
// rows, folded into a running max, then max with x[k] and the left-padded pool term.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) timeconv_kernel(const float* __restrict__ packed,
                                                      const float* __restrict__ x, int B, int L,
                                                      float* __restrict__ out) {
  __shared__ Smem s;
  const int64_t M = (int64_t)B * L;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;

  float best[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) best[i][j] = -INFINITY;

  const size_t w_off[3] = {TimeConvPacked::w3_off, TimeConvPacked::w5_off, TimeConvPacked::w7_off};
  const size_t b_off[3] = {TimeConvPacked::b3_off, TimeConvPacked::b5_off, TimeConvPacked::b7_off};
#pragma unroll 1
  for (int conv = 0; conv < 3; ++conv) {
    const int taps = 3 + 2 * conv;
    const int half = taps / 2;
    const float* w = packed + w_off[conv];
    const int64_t ldw = (int64_t)taps * kD;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    auto afn = [&](int r, int kt) -> const float* {
      const int64_t m = m0 + r;
      if (m >= M) return nullptr;
      const int tap = kt / (kD / BK);
      const int c0 = (kt % (kD / BK)) * BK;
      const int64_t b = m / L;
      const int kk = (int)(m - b * L) + tap - half;
      if (kk < 0 || kk >= L) return nullptr;          // zero "same" padding inside the window
      return x + (b * L + kk) * kD + c0;
    };
    auto bfn = [&](int r, int kt) -> const float* { return w + (int64_t)(n0 + r) * ldw + kt * BK; };
    mainloop(acc, afn, bfn, taps * (kD / BK), s);

    const float* bias = packed + b_off[conv];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float bj = __ldg(bias + n0 + tile_col(tx, j));
#pragma unroll
      for (int i = 0; i < 8; ++i) best[i][j] = fmaxf(best[i][j], acc[i][j] + bj);
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + tile_row(ty, i);
    if (m >= M) continue;
    const int k = (int)(m % L);
#pragma unroll
    for (int jq = 0; jq < 2; ++jq) {
      const int n = n0 + tile_col(tx, jq * 4);
      const float4 xc = __ldg(reinterpret_cast<const float4*>(x + m * kD + n));
      // F.pad(x,(1,0)) + MaxPool1d(2,1): the zero pad takes part in the max at k == 0 (NLB:67-68)
      const float4 xp = (k > 0) ? __ldg(reinterpret_cast<const float4*>(x + (m - 1) * kD + n))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 o;
      o.x = fmaxf(fmaxf(best[i][jq * 4 + 0], xc.x), xp.x);
      o.y = fmaxf(fmaxf(best[i][jq * 4 + 1], xc.y), xp.y);
      o.z = fmaxf(fmaxf(best[i][jq * 4 + 2], xc.z), xp.z);
      o.w = fmaxf(fmaxf(best[i][jq * 4 + 3], xc.w), xp.w);
      *reinterpret_cast<float4*>(out + m * kD + n) = o;
    }
  }
}

int simt_timeconv(const float* packed, const float* x, int B, int L, float* out, cudaStream_t st) {
  if (B == 0) return TMR_OK;
  const int64_t M = (int64_t)B * L;
  dim3 grid(kD / BN, (unsigned)((M + BM - 1) / BM));
  timeconv_kernel<<<grid, NT, 0, st>>>(packed, x, B, L, out);
  TMR_LAUNCH_CHECK("timeconv_kernel");
  return TMR_OK;
}

// -------------------------------------------------------------------------------------------
// LSTM recurrent step (torch.nn.LSTM semantics, TRAIN:224,241-244) on gate-interleaved weights:
// column n = unit*4 + gate, so each thread's float4 of accumulators is (i,f,g,o) of one unit.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void __launch_bounds__(NT) lstm_step_kernel(const float* __restrict__ whh,
                                                       const float* __restrict__ xp,
                                                       const int64_t* __restrict__ starts, int seq,
                                                       int t, const float* __restrict__ h_prev,
                                                       float* __restrict__ h_out,
                                                       float* __restrict__ c, int B) {
  __shared__ Smem s;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;            // 128 gate columns = 32 hidden units
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  auto afn = [&](int r, int kt) -> const float* {
    const int64_t m = m0 + r;
    return (m < B) ? h_prev + m * kD + kt * BK : nullptr;
  };
  auto bfn = [&](int r, int kt) -> const float* { return whh + (int64_t)(n0 + r) * kD + kt * BK; };
  mainloop(acc, afn, bfn, kD / BK, s);

  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + tile_row(ty, i);
    if (m >= B) continue;
    const int64_t xr = (starts ? starts[m] : m * seq) + t;
#pragma unroll
    for (int jq = 0; jq < 2; ++jq) {
      const int n = n0 + tile_col(tx, jq * 4);          // multiple of 4: gates of unit n/4
      const int unit = n >> 2;
      const float4 p = __ldg(reinterpret_cast<const float4*>(xp + xr * (4 * kD) + n));
      const float gi = acc[i][jq * 4 + 0] + p.x, gf = acc[i][jq * 4 + 1] + p.y;
      const float gg = acc[i][jq * 4 + 2] + p.z, go = acc[i][jq * 4 + 3] + p.w;
      const float cn = sigmoidf_(gf) * c[m * kD + unit] + sigmoidf_(gi) * tanhf(gg);
      c[m * kD + unit] = cn;
      h_out[m * kD + unit] = sigmoidf_(go) * tanhf(cn);
    }
  }
}

int simt_lstm_step(const float* whh, const float* xp, const int64_t* starts, int seq, int t,
                   const float* h_prev, float* h_out, float* c, int B, cudaStream_t st) {
  if (B == 0) return TMR_OK;
  dim3 grid(4 * kD / BN, (unsigned)((B + BM - 1) / BM));
  lstm_step_kernel<<<grid, NT, 0, st>>>(whh, xp, starts, seq, t, h_prev, h_out, c, B);
  TMR_LAUNCH_CHECK("lstm_step_kernel");
  return TMR_OK;
}

}  // namespace tmr
