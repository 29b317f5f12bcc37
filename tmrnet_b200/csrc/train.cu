// Head training step: forward in training mode + full backward of the TMRNet
// head, fp32 on CUDA cores (training batches are tens to a few hundred clips: launch-bound, so the
// register-tiled FFMA GEMM of sgemm_simt.cuh is used for every product, with explicit transposes for
// the dX = dY.W and dW = dY^T.X forms).  Reference: resnet_lstm.forward in training mode
// (train_non-local_mutiConv_resnet.py:237-253), CrossEntropyLoss(reduction='sum', weight) (:780,883),
// loss.backward() (:886); NLBlock / TimeConv (NLBlock_MutiConv6_3.py:10-79); torch.nn.LSTM.
// Backbone features and the memory bank are inputs without gradient (frozen features).
//
// Parameter / gradient order (24 tensors, reference state-dict layouts):
//   0 lstm.weight_ih_l0 (4D,F)  1 lstm.weight_hh_l0 (4D,D)  2 lstm.bias_ih_l0  3 lstm.bias_hh_l0
//   4 timeconv1.weight (D,D,3)  5 .bias   6 timeconv2.weight (D,D,5)  7 .bias   8 timeconv3.weight (D,D,7)  9 .bias
//   10 nl.linear1.weight 11 .bias 12 nl.linear2.weight 13 .bias 14 nl.linear3.weight 15 .bias 16 nl.linear4.weight 17 .bias
//   18 nl.layer_norm.weight (1,D) 19 .bias   20 fc_h_c.weight (D,2D) 21 .bias   22 fc_c.weight (C,D) 23 .bias
#include "tmr_internal.h"

namespace tmr {
namespace train {

static inline size_t fb(size_t n) { return align_up(n * sizeof(float), 256); }
static inline int64_t pad16(int64_t v) { return (v + 63) / 64 * 64; }   // 64: a whole k-step of the tensor-core GEMM

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
#define GRID_STRIDE(i, n) for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)
static inline unsigned nblk(int64_t n) { int64_t b = (n + 255) / 256; return (unsigned)(b < 1 ? 1 : (b > 8192 ? 8192 : b)); }

// dst[c][r] = src[r][c] for r < R, 0 for R <= r < Rpad.   src: R x Ccols (ld), dst: Ccols x Rpad
__global__ void transpose_pad_kernel(const float* __restrict__ src, int64_t R, int Ccols, int64_t ld, float* __restrict__ dst, int64_t Rpad) {
  __shared__ float t[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32;
  const int c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i; const int c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < R && c < Ccols) ? src[r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i; const int64_t r = r0 + threadIdx.x;
    if (c < Ccols && r < Rpad) dst[(int64_t)c * Rpad + r] = t[threadIdx.x][i];
  }
}
static int transpose_pad(const float* src, int64_t R, int Ccols, int64_t ld, float* dst, int64_t Rpad, cudaStream_t st) {
  dim3 grid((Ccols + 31) / 32, (unsigned)((Rpad + 31) / 32));
  transpose_pad_kernel<<<grid, dim3(32, 8), 0, st>>>(src, R, Ccols, ld, dst, Rpad);
  TMR_LAUNCH_CHECK("transpose_pad_kernel");
  return TMR_OK;
}

// out[c] (+)= sum_r X[r][c] (* Y[r][c])
__global__ void colsum_kernel(const float* __restrict__ X, const float* __restrict__ Y, int64_t R, int Ccols, int64_t ld,
                              float* __restrict__ out, int accumulate) {
  // 32 columns x 32 row lanes per CTA, four independent partial sums per thread: the bias gradients of the convolutions
  // sum 1200 rows, and with 8 row lanes and one dependent chain each of those launches took ~25 us
  const int c = blockIdx.x * 32 + threadIdx.x;
  __shared__ float part[32][33];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < Ccols) {
    const int64_t step = blockDim.y;
    int64_t r = threadIdx.y;
    for (; r + 3 * step < R; r += 4 * step) {
      const int64_t i0 = r * ld + c, i1 = (r + step) * ld + c, i2 = (r + 2 * step) * ld + c, i3 = (r + 3 * step) * ld + c;
      s0 += Y ? X[i0] * Y[i0] : X[i0]; s1 += Y ? X[i1] * Y[i1] : X[i1];
      s2 += Y ? X[i2] * Y[i2] : X[i2]; s3 += Y ? X[i3] * Y[i3] : X[i3];
    }
    for (; r < R; r += step) s0 += Y ? X[r * ld + c] * Y[r * ld + c] : X[r * ld + c];
  }
  part[threadIdx.y][threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (threadIdx.y == 0 && c < Ccols) {
    float tsum = 0.f;
    for (int i = 0; i < 32; ++i) tsum += part[i][threadIdx.x];
    out[c] = accumulate ? out[c] + tsum : tsum;
  }
}
static int colsum(const float* X, const float* Y, int64_t R, int Ccols, int64_t ld, float* out, int accumulate, cudaStream_t st) {
  colsum_kernel<<<(Ccols + 31) / 32, dim3(32, 32), 0, st>>>(X, Y, R, Ccols, ld, out, accumulate);
  TMR_LAUNCH_CHECK("colsum_kernel");
  return TMR_OK;
}

__global__ void add_vec_kernel(const float* a, const float* b, float* o, int64_t n) { GRID_STRIDE(i, n) o[i] = a[i] + b[i]; }
__global__ void copy_vec_kernel(const float* a, float* o, int64_t n) { GRID_STRIDE(i, n) o[i] = a[i]; }
__global__ void zero_kernel(float* o, int64_t n) { GRID_STRIDE(i, n) o[i] = 0.f; }
// (B,S,F) -> (S,B,F)
__global__ void permute_bsf_kernel(const float* __restrict__ x, int B, int S, int Fw, float* __restrict__ o) {
  const int64_t n = (int64_t)B * S * Fw;
  GRID_STRIDE(i, n) {
    const int f = (int)(i % Fw); const int64_t r = i / Fw; const int b = (int)(r % B); const int t = (int)(r / B);
    o[i] = x[((int64_t)b * S + t) * Fw + f];
  }
}

__device__ __forceinline__ float sigm(float v) { return 1.f / (1.f + expf(-v)); }

// LSTM cell, training forward: pre = xp (+ hh); standard gate order i,f,g,o in 4 blocks of D columns.
__global__ void lstm_cell_fwd_kernel(const float* __restrict__ xp, const float* __restrict__ hh, const float* __restrict__ c_prev,
                                     float* __restrict__ gates, float* __restrict__ c, float* __restrict__ h, int B) {
  const int64_t n = (int64_t)B * kD;
  GRID_STRIDE(idx, n) {
    const int64_t b = idx / kD; const int u = (int)(idx % kD);
    const float* xr = xp + b * 4 * kD; const float* hr = hh ? hh + b * 4 * kD : nullptr;
    const float pi = xr[u] + (hr ? hr[u] : 0.f), pf = xr[kD + u] + (hr ? hr[kD + u] : 0.f);
    const float pg = xr[2 * kD + u] + (hr ? hr[2 * kD + u] : 0.f), po = xr[3 * kD + u] + (hr ? hr[3 * kD + u] : 0.f);
    const float gi = sigm(pi), gf = sigm(pf), gg = tanhf(pg), go = sigm(po);
    const float cn = gf * (c_prev ? c_prev[idx] : 0.f) + gi * gg;
    float* gr = gates + b * 4 * kD;
    gr[u] = gi; gr[kD + u] = gf; gr[2 * kD + u] = gg; gr[3 * kD + u] = go;
    c[idx] = cn; h[idx] = go * tanhf(cn);
  }
}
// backward of one cell: dh, dc (in: from t+1, out: for t-1) -> dpre (B,4D)
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ dh, float* __restrict__ dc, const float* __restrict__ gates,
                                     const float* __restrict__ c, const float* __restrict__ c_prev, float* __restrict__ dpre, int B) {
  const int64_t n = (int64_t)B * kD;
  GRID_STRIDE(idx, n) {
    const int64_t b = idx / kD; const int u = (int)(idx % kD);
    const float* gr = gates + b * 4 * kD;
    const float gi = gr[u], gf = gr[kD + u], gg = gr[2 * kD + u], go = gr[3 * kD + u];
    const float tc = tanhf(c[idx]);
    const float dhv = dh[idx];
    const float dcv = dc[idx] + dhv * go * (1.f - tc * tc);
    float* dr = dpre + b * 4 * kD;
    dr[u] = dcv * gg * gi * (1.f - gi);
    dr[kD + u] = dcv * (c_prev ? c_prev[idx] : 0.f) * gf * (1.f - gf);
    dr[2 * kD + u] = dcv * gi * (1.f - gg * gg);
    dr[3 * kD + u] = dhv * tc * go * (1.f - go);
    dc[idx] = dcv * gf;
  }
}

// TimeConv training forward on precomputed conv outputs (bias included): Lt = max(x, pool, c3, c5, c7);
// branch = 0 if a non-parameter branch (identity / pool) wins, else 1,2,3 for conv3/5/7.  torch's
// max over cat((y0,y1,y2,y3,y4)) sends the gradient to the FIRST maximal entry (y0 = x first).
__global__ void timeconv_max_train_kernel(const float* __restrict__ x, const float* __restrict__ c3, const float* __restrict__ c5,
                                          const float* __restrict__ c7, int L, int64_t n, float* __restrict__ Lt,
                                          uint8_t* __restrict__ branch) {
  GRID_STRIDE(i, n) {
    const int64_t row = i / kD; const int k = (int)(row % L);
    const float x0 = x[i];
    const float pool = fmaxf(x0, k > 0 ? x[i - kD] : 0.f);
    float best = x0; uint8_t br = 0;
    if (c3[i] > best) { best = c3[i]; br = 1; }
    if (c5[i] > best) { best = c5[i]; br = 2; }
    if (c7[i] > best) { best = c7[i]; br = 3; }
    if (pool > best) { best = pool; br = 0; }
    Lt[i] = best; branch[i] = br;
  }
}
__global__ void conv_route_kernel(const float* __restrict__ dLt, const uint8_t* __restrict__ branch, int64_t n,
                                  float* __restrict__ d3, float* __restrict__ d5, float* __restrict__ d7) {
  GRID_STRIDE(i, n) {
    const float g = dLt[i]; const uint8_t b = branch[i];
    d3[i] = b == 1 ? g : 0.f; d5[i] = b == 2 ? g : 0.f; d7[i] = b == 3 ? g : 0.f;
  }
}
// out[(b,k)][c] = x[(b,k+d)][c] or 0 outside the window
__global__ void shift_rows_kernel(const float* __restrict__ x, int L, int d, int64_t n, float* __restrict__ o) {
  GRID_STRIDE(i, n) {
    const int64_t row = i / kD; const int k = (int)(row % L);
    const int kk = k + d;
    o[i] = (kk >= 0 && kk < L) ? x[i + (int64_t)d * kD] : 0.f;
  }
}
// dW[(o*D + c)*K + j] = tmp[o*D + c]
__global__ void scatter_tap_kernel(const float* __restrict__ tmp, int K, int j, float* __restrict__ dW) {
  const int64_t n = (int64_t)kD * kD;
  GRID_STRIDE(i, n) dW[i * K + j] = tmp[i];
}
// direct conv forward for training: out[(b,k)][o] = bias[o] + sum_j sum_c w[o][c][j] x[(b,k+j-h)][c] is done as
// K GEMMs on shifted copies (see conv_forward()).
//
// TMR_MATH_F16 batches the taps of a convolution into ONE GEMM each way (15 + 15 tap GEMMs with their conversions,
// shifts and scatters cost ~1.3 ms of launch-bound time per step): the seven shifted copies of the window rows sit side
// by side in xcat[r][(d+3)*512 + c] = x[(b,k+d)][c] (0 outside the window), d = -3..3, as fp16; conv_K reads the 512*K
// columns of its taps (K-dimension of the GEMM) against wcat_K[o][j*512 + c] = w_K[o][c][j]; the weight gradient is
// dcat_K[o][j*512 + c] = sum_r dconv_K[r][o] xcat[r][(j-h+3)*512 + c], permuted back to the (o, c, j) layout.
constexpr int kTaps = 7;
__global__ void build_xcat16_kernel(const float* __restrict__ x, int L, int64_t R, half_t* __restrict__ xcat) {
  const int64_t n4 = R * kTaps * (kD / 4);
  GRID_STRIDE(i, n4) {
    const int c4 = (int)(i % (kD / 4));
    const int dj = (int)((i / (kD / 4)) % kTaps);
    const int64_t row = i / ((int64_t)kTaps * (kD / 4));
    const int kk = (int)(row % L) + dj - 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kk >= 0 && kk < L) v = __ldg(reinterpret_cast<const float4*>(x + (row + dj - 3) * kD) + c4);
    reinterpret_cast<uint2*>(xcat + (row * kTaps + dj) * kD)[c4] = pack_h4(v);
  }
}
// wcat[o][j*512 + c] = fp16(w[o][c][j]),  w: (512, 512, K)
__global__ void permute_wcat16_kernel(const float* __restrict__ w, int K, half_t* __restrict__ wcat) {
  const int64_t n = (int64_t)kD * kD * K;
  GRID_STRIDE(i, n) {
    const int c = (int)(i % kD); const int j = (int)((i / kD) % K); const int64_t o = i / ((int64_t)kD * K);
    const float v = w[(o * kD + c) * K + j];
    wcat[i] = (half_t)(pack_h2(v, 0.f) & 0xffffu);
  }
}
// dW[(o*512 + c)*K + j] = dcat[o][j*512 + c]
__global__ void permute_dwcat_kernel(const float* __restrict__ dcat, int K, float* __restrict__ dW) {
  const int64_t n = (int64_t)kD * kD * K;
  GRID_STRIDE(i, n) {
    const int j = (int)(i % K); const int c = (int)((i / K) % kD); const int64_t o = i / ((int64_t)kD * K);
    dW[i] = dcat[(o * K + j) * kD + c];
  }
}
// dst[c][r] = fp16(src[r][c]) for r < R, 0 for R <= r < Rpad.  SRC = float or half_t
template <class SRC>
__global__ void transpose_pad16_kernel(const SRC* __restrict__ src, int64_t R, int Ccols, int64_t ld, half_t* __restrict__ dst, int64_t Rpad) {
  __shared__ half_t t[32][34];
  const int64_t r0 = (int64_t)blockIdx.y * 32;
  const int c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i; const int c = c0 + threadIdx.x;
    half_t v = 0;
    if (r < R && c < Ccols) {
      if constexpr (sizeof(SRC) == 4) v = (half_t)(pack_h2((float)src[r * ld + c], 0.f) & 0xffffu);
      else v = (half_t)src[r * ld + c];
    }
    t[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i; const int64_t r = r0 + threadIdx.x;
    if (c < Ccols && r < Rpad) dst[(int64_t)c * Rpad + r] = t[threadIdx.x][i];
  }
}
template <class SRC>
static int transpose_pad16(const SRC* src, int64_t R, int Ccols, int64_t ld, half_t* dst, int64_t Rpad, cudaStream_t st) {
  dim3 grid((Ccols + 31) / 32, (unsigned)((Rpad + 31) / 32));
  transpose_pad16_kernel<SRC><<<grid, dim3(32, 8), 0, st>>>(src, R, Ccols, ld, dst, Rpad);
  TMR_LAUNCH_CHECK("transpose_pad16_kernel");
  return TMR_OK;
}

// fp16 copies of BOTH operands of a tensor-core GEMM in one launch (a: M x K with row stride lda, w: N x K with ldw;
// K % 4 == 0): a training step converts ~25 operand pairs, and a launch costs more than either conversion.
__global__ void half_pair_kernel(const float* __restrict__ a, int64_t lda, int64_t M, const float* __restrict__ w, int64_t ldw,
                                 int64_t N, int K, half_t* __restrict__ a16, half_t* __restrict__ w16) {
  const int k4 = K / 4;
  const int64_t na = M * k4, n = na + N * k4;
  GRID_STRIDE(i, n) {
    const bool first = i < na;
    const int64_t j = first ? i : i - na;
    const int64_t r = j / k4; const int c = (int)(j - r * k4);
    const float4 v = __ldg(reinterpret_cast<const float4*>((first ? a + r * lda : w + r * ldw)) + c);
    reinterpret_cast<uint2*>(first ? a16 : w16)[j] = pack_h4(v);
  }
}

// attention, training forward: ONE CTA of 8 warps per clip (a training batch is 40 clips: a warp per clip left the
// GPU with 40 warps walking 30 slots one global round trip at a time, 103 us).  Warp w takes slots w, w + 8, ..; a lane
// owns 16 channels as four coalesced float4; scores go through shared memory, the softmax is computed redundantly by
// every thread, and the weighted sum runs one float2 of channels per thread over all slots (rows now in L1 / L2).
// Saves the softmax p (B,L).
constexpr int kAttnTrainWarps = 8;
constexpr int kAttnTrainMaxL = 512;
__device__ __forceinline__ float dot16(const float4 (&a)[4], const float4* __restrict__ row, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = __ldg(row + i * 32 + lane);
    s = fmaf(a[i].x, v.x, s); s = fmaf(a[i].y, v.y, s); s = fmaf(a[i].z, v.z, s); s = fmaf(a[i].w, v.w, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}
__global__ void __launch_bounds__(kAttnTrainWarps * 32)
attention_train_fwd_kernel(const float* __restrict__ u, const float* __restrict__ Lt, int B, int L, float scale,
                           float* __restrict__ p, float* __restrict__ abar) {
  __shared__ float sc[kAttnTrainMaxL];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const float4* Lb = reinterpret_cast<const float4*>(Lt + (int64_t)b * L * kD);
  float4 uq[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) uq[i] = __ldg(reinterpret_cast<const float4*>(u + (int64_t)b * kD) + i * 32 + lane);
  for (int k = warp; k < L; k += kAttnTrainWarps) {
    const float s = dot16(uq, Lb + (int64_t)k * (kD / 4), lane) * scale;
    if (lane == 0) sc[k] = s;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int k = 0; k < L; ++k) mx = fmaxf(mx, sc[k]);
  float den = 0.f;
  for (int k = 0; k < L; ++k) den += expf(sc[k] - mx);
  __syncthreads();                                       // every thread has read the raw scores
  for (int k = threadIdx.x; k < L; k += blockDim.x) {
    const float pk = expf(sc[k] - mx) / den;
    sc[k] = pk;
    p[(int64_t)b * L + k] = pk;
  }
  __syncthreads();
  // abar[c] = sum_k p_k Lt[k][c]: thread t owns channels 2t, 2t+1
  const float2* L2 = reinterpret_cast<const float2*>(Lt + (int64_t)b * L * kD);
  float2 a = make_float2(0.f, 0.f);
  for (int k = 0; k < L; ++k) {
    const float2 v = __ldg(L2 + (int64_t)k * (kD / 2) + threadIdx.x);
    a.x = fmaf(sc[k], v.x, a.x); a.y = fmaf(sc[k], v.y, a.y);
  }
  reinterpret_cast<float2*>(abar + (int64_t)b * kD)[threadIdx.x] = a;
}
// attention backward: dabar -> du, dLt.  dp_k = dabar.Lt_k ; ds = p (dp - sum p dp) ; du = scale sum ds_k Lt_k ;
// dLt_k = p_k dabar + scale ds_k u.  Same CTA-per-clip shape as the forward.
__global__ void __launch_bounds__(kAttnTrainWarps * 32)
attention_train_bwd_kernel(const float* __restrict__ u, const float* __restrict__ Lt, const float* __restrict__ p,
                           const float* __restrict__ dabar, int B, int L, float scale, float* __restrict__ ds_buf,
                           float* __restrict__ du, float* __restrict__ dLt) {
  __shared__ float sd[kAttnTrainMaxL];
  __shared__ float sp[kAttnTrainMaxL];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const float4* Lb = reinterpret_cast<const float4*>(Lt + (int64_t)b * L * kD);
  const float* pb = p + (int64_t)b * L;
  float4 dq[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) dq[i] = __ldg(reinterpret_cast<const float4*>(dabar + (int64_t)b * kD) + i * 32 + lane);
  for (int k = warp; k < L; k += kAttnTrainWarps) {
    const float s = dot16(dq, Lb + (int64_t)k * (kD / 4), lane);
    if (lane == 0) { sd[k] = s; sp[k] = pb[k]; }
  }
  __syncthreads();
  float dot = 0.f;
  for (int k = 0; k < L; ++k) dot = fmaf(sp[k], sd[k], dot);
  __syncthreads();
  for (int k = threadIdx.x; k < L; k += blockDim.x) {
    const float ds = sp[k] * (sd[k] - dot);
    sd[k] = ds;
    ds_buf[(int64_t)b * L + k] = ds;
  }
  __syncthreads();
  const float2* L2 = reinterpret_cast<const float2*>(Lt + (int64_t)b * L * kD);
  const float2 da = __ldg(reinterpret_cast<const float2*>(dabar + (int64_t)b * kD) + threadIdx.x);
  const float2 ub = __ldg(reinterpret_cast<const float2*>(u + (int64_t)b * kD) + threadIdx.x);
  float2 acc = make_float2(0.f, 0.f);
  for (int k = 0; k < L; ++k) {
    const float2 v = __ldg(L2 + (int64_t)k * (kD / 2) + threadIdx.x);
    acc.x = fmaf(sd[k], v.x, acc.x); acc.y = fmaf(sd[k], v.y, acc.y);
    if (dLt)
      reinterpret_cast<float2*>(dLt + ((int64_t)b * L + k) * kD)[threadIdx.x] =
          make_float2(sp[k] * da.x + scale * sd[k] * ub.x, sp[k] * da.y + scale * sd[k] * ub.y);
  }
  reinterpret_cast<float2*>(du + (int64_t)b * kD)[threadIdx.x] = make_float2(scale * acc.x, scale * acc.y);
}

// LayerNorm([1,D]) forward (saves xhat, rstd) + ReLU
__global__ void layernorm_train_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ be, int B,
                                           float* __restrict__ xhat, float* __restrict__ rstd, float* __restrict__ nout, float* __restrict__ r) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* vb = v + (int64_t)b * kD;
  float s = 0.f;
  for (int c = lane; c < kD; c += 32) s += vb[c];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / kD;
  float q = 0.f;
  for (int c = lane; c < kD; c += 32) { const float a = vb[c] - mean; q += a * a; }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rs = rsqrtf(q / kD + 1e-5f);
  if (lane == 0) rstd[b] = rs;
  for (int c = lane; c < kD; c += 32) {
    const float xh = (vb[c] - mean) * rs;
    const float nn = xh * g[c] + be[c];
    xhat[(int64_t)b * kD + c] = xh; nout[(int64_t)b * kD + c] = nn; r[(int64_t)b * kD + c] = fmaxf(nn, 0.f);
  }
}
// dn (already masked by relu) -> dv
__global__ void layernorm_train_bwd_kernel(const float* __restrict__ dn, const float* __restrict__ xhat, const float* __restrict__ rstd,
                                           const float* __restrict__ g, int B, float* __restrict__ dv) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* d = dn + (int64_t)b * kD; const float* xh = xhat + (int64_t)b * kD;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < kD; c += 32) { const float dx = d[c] * g[c]; s1 += dx; s2 += dx * xh[c]; }
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  s1 /= kD; s2 /= kD;
  const float rs = rstd[b];
  for (int c = lane; c < kD; c += 32) dv[(int64_t)b * kD + c] = rs * (d[c] * g[c] - s1 - xh[c] * s2);
}

// counter-based dropout mask: keep with probability 1-p; out = in * keep / (1-p)
__device__ __forceinline__ uint32_t hash32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return (uint32_t)x;
}
__global__ void dropout_kernel(const float* __restrict__ in, float p, uint64_t seed, int64_t n, float* __restrict__ out, float* __restrict__ mask) {
  const float inv = p < 1.f ? 1.f / (1.f - p) : 0.f;
  GRID_STRIDE(i, n) {
    const float keep = (p <= 0.f) ? 1.f : ((hash32(seed * 0x9E3779B97F4A7C15ULL + (uint64_t)i) * (1.0f / 4294967296.0f)) >= p ? 1.f : 0.f);
    const float m = keep * inv;
    mask[i] = m; out[i] = in[i] * m;
  }
}
__global__ void mul_mask_kernel(const float* a, const float* m, float* o, int64_t n) { GRID_STRIDE(i, n) o[i] = a[i] * m[i]; }
__global__ void relu_fwd_kernel(const float* a, float* o, int64_t n) { GRID_STRIDE(i, n) o[i] = fmaxf(a[i], 0.f); }
__global__ void relu_bwd_kernel(const float* d, const float* pre, float* o, int64_t n) { GRID_STRIDE(i, n) o[i] = pre[i] > 0.f ? d[i] : 0.f; }
// cat = [St || y1], y1 = St + o'
__global__ void make_cat_kernel(const float* St, const float* od, int B, float* cat) {
  const int64_t n = (int64_t)B * kD;
  GRID_STRIDE(i, n) { const int64_t b = i / kD; const int c = (int)(i % kD); cat[b * 2 * kD + c] = St[i]; cat[b * 2 * kD + kD + c] = St[i] + od[i]; }
}
// dSt = dcat[:, :D] + dcat[:, D:] ; do' = dcat[:, D:]
__global__ void split_dcat_kernel(const float* dcat, int B, float* dSt, float* dod) {
  const int64_t n = (int64_t)B * kD;
  GRID_STRIDE(i, n) { const int64_t b = i / kD; const int c = (int)(i % kD); const float a = dcat[b * 2 * kD + c], d = dcat[b * 2 * kD + kD + c]; dSt[i] = a + d; dod[i] = d; }
}
__global__ void axpy_kernel(const float* x, float* y, int64_t n) { GRID_STRIDE(i, n) y[i] += x[i]; }

// CrossEntropyLoss(reduction='sum', weight=w): loss = sum_b w[y_b] (-log softmax(logits_b)[y_b]);
// dlogits (ld 16, zero padded) = w[y_b] (softmax - onehot).  One thread per clip; loss accumulated atomically.
__global__ void ce_loss_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ cw, int B, int C,
                               float* __restrict__ dlog16, float* __restrict__ loss, int64_t* __restrict__ pred) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* l = logits + (int64_t)b * C;
  float mx = l[0]; int am = 0;
  for (int c = 1; c < C; ++c) if (l[c] > mx) { mx = l[c]; am = c; }
  float den = 0.f;
  for (int c = 0; c < C; ++c) den += expf(l[c] - mx);
  const int64_t yl = labels[b];
  if (yl < 0 || yl >= C) {                 // torch raises here; on the device: no out-of-bounds read, NaN loss / gradient
    const float nan = __int_as_float(0x7fc00000);
    for (int c = 0; c < 16; ++c) dlog16[(int64_t)b * 16 + c] = c < C ? nan : 0.f;
    atomicAdd(loss, nan);
    if (pred) pred[b] = am;
    return;
  }
  const int y = (int)yl;
  const float w = cw ? cw[y] : 1.f;
  for (int c = 0; c < 16; ++c) dlog16[(int64_t)b * 16 + c] = c < C ? w * (expf(l[c] - mx) / den - (c == y ? 1.f : 0.f)) : 0.f;
  atomicAdd(loss, w * (logf(den) + mx - l[y]));
  if (pred) pred[b] = am;
}

// SGD with momentum and weight decay exactly as torch.optim.SGD (dampening 0, no nesterov):
//   d = g + wd * p ; buf = first ? d : mu * buf + d ; p -= lr * buf          (TRAIN:797-805, 887)
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, int64_t n, float lr, float mu,
                           float wd, int first) {
  GRID_STRIDE(i, n) {
    const float d = g[i] + wd * p[i];
    const float bv = first ? d : mu * buf[i] + d;
    buf[i] = bv;
    p[i] -= lr * bv;
  }
}

// the same update over up to 24 tensors in ONE launch (blockIdx.y = tensor; each with its own learning rate: the
// reference's parameter groups)
struct SgdTable { float* p[24]; const float* g[24]; float* buf[24]; int64_t n[24]; float lr[24]; };
__global__ void sgd_multi_kernel(SgdTable t, float mu, float wd, int first) {
  const int k = blockIdx.y;
  float* __restrict__ p = t.p[k]; const float* __restrict__ g = t.g[k]; float* __restrict__ buf = t.buf[k];
  const int64_t n = t.n[k];
  const float lr = t.lr[k];
  if (!p) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = g[i] + wd * p[i];
    const float bv = first ? d : mu * buf[i] + d;
    buf[i] = bv;
    p[i] -= lr * bv;
  }
}

// ---------------------------------------------------------------------------------------------
// GEMM helpers on top of simt_linear (out = a . w^T, K-major operands)
// ---------------------------------------------------------------------------------------------

// fp16 operand scratch of the tensor-core GEMMs: the largest [rows][K] operand of the step
static inline size_t tc_scratch_halves(size_t T, size_t Tp, size_t R, size_t Rp) {
  size_t rows = kF;                                   // Wih (4D x F), x^T (F x Tp), dpre^T (4D x Tp) ...
  size_t k = kF > Tp ? kF : Tp;
  if (Rp > k) k = Rp;
  size_t a = rows * k, b = (T > R ? T : R) * (size_t)kF;
  return a > b ? a : b;
}

struct Ws {
  char* p; size_t left;
  float* f(size_t n) { size_t b = fb(n); if (b > left) return nullptr; float* r = (float*)p; p += b; left -= b; return r; }
};

}  // namespace train
}  // namespace tmr

using namespace tmr;
using namespace tmr::train;

extern "C" {

size_t tmr_head_train_workspace_bytes(int B, int seq, int L, int D, int F, int C) {
  (void)C;
  const size_t b = (size_t)(B > 0 ? B : 1), S = seq, R = b * L, Rp = pad16(R), Bp = pad16(b), T = b * S, Tp = pad16(T);
  size_t n = 0;
  n += fb(T * F) + fb(4 * D) + fb(T * 4 * D) + fb(b * 4 * D) + fb(T * 4 * D) + 2 * fb(T * D);        // LSTM forward
  n += 5 * fb(R * D) + fb(R * D / 4 + 64) + fb((size_t)D * D);                                        // TimeConv forward
  n += 4 * fb((size_t)D * D) + fb((size_t)2 * D * D) + fb((size_t)D * 4 * D) + fb((size_t)D * 16);    // weight transposes
  n += 26 * fb(b * D) + 2 * fb(b * 2 * D) + 2 * fb(b * L) + fb(b) + fb(b * 16);                       // activations + their grads
  n += 2 * fb((size_t)2 * D * Bp) + fb((size_t)16 * Bp);                                              // small transposes
  n += 4 * fb(R * D) + 2 * fb((size_t)D * Rp) + fb((size_t)D * D);                                    // TimeConv backward
  n += fb(T * 4 * D) + fb((size_t)4 * D * Tp) + fb((size_t)F * Tp);                                   // BPTT
  n += 2 * fb(tc_scratch_halves(T, Tp, R, Rp) / 2) + 2 * fb((size_t)4 * D * D / 2);                   // fp16 operands (TMR_MATH_F16)
  n += fb(R * 7 * D / 2) + fb((size_t)7 * D * Rp / 2) + fb((size_t)15 * D * D / 2) + fb((size_t)D * Rp / 2) + fb((size_t)7 * D * D);   // batched taps
  return n + (1 << 16);
}

}  // extern "C"

/* Forward (training mode) + loss + backward of the head for one batch, as phases over ONE workspace layout
 * (the carving below is identical whatever the phase, so activations saved by a forward-only call are found
 * again by a later backward-only call on the same workspace):
 *   PH_FWD   features, window -> logits (saves gates / c / h of all steps, TimeConv branch winners, softmax p,
 *            LayerNorm xhat / rstd, dropout masks)
 *   PH_LOSS  CrossEntropyLoss(reduction='sum'[, weight]) and its dlogits
 *   PH_BWD   dlogits -> grads of the 24 parameters (overwritten, not accumulated)
 * dlogits_ext (B,C): the caller's gradient w.r.t. the logits (autograd) instead of the built-in loss. */
enum { PH_FWD = 1, PH_LOSS = 2, PH_BWD = 4 };

// TMR_MATH_F16: the backward runs on gradients scaled by 2^10 (exact in fp32), so that the small entries of dpre / dLt
// stay inside fp16's normal range when they become GEMM operands; every parameter gradient is linear in dlogits and is
// scaled back by one launch over the 24 tensors at the end.
constexpr float kGradScale = 1024.f;
__global__ void scale_vec_kernel(float* p, int64_t n, float sc) { GRID_STRIDE(i, n) p[i] *= sc; }
struct GradTable { float* p[24]; int64_t n[24]; };
__global__ void unscale_grads_kernel(GradTable t, float sc) {
  float* p = t.p[blockIdx.y];
  const int64_t n = t.n[blockIdx.y];
  if (!p) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {           // 128-bit body, scalar tail
    float4* p4 = reinterpret_cast<float4*>(p);
    for (int64_t i = i0; i < n / 4; i += stride) { float4 v = p4[i]; v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; p4[i] = v; }
    for (int64_t i = (n / 4) * 4 + i0; i < n; i += stride) p[i] *= sc;
  } else {
    for (int64_t i = i0; i < n; i += stride) p[i] *= sc;
  }
}

__global__ void pad_dlogits_kernel(const float* __restrict__ d, int B, int C, float* __restrict__ d16) {
  const int64_t n = (int64_t)B * 16;
  GRID_STRIDE(i, n) { const int64_t b = i / 16; const int c = (int)(i % 16); d16[i] = c < C ? d[b * C + c] : 0.f; }
}

static int train_impl(int phase, const float* const* params, float* const* grads, const float* x, const float* long_feature,
                      const int64_t* labels, const float* class_weight, const float* dlogits_ext, int B, int seq, int L,
                      int F, int D, int C, float p_nl, float p_fc, uint64_t seed, float* logits, float* loss, int64_t* pred,
                      void* workspace, size_t workspace_bytes, int math_mode, void* stream) {
  TMR_CHECK_ARG(D == kD && F == kF, "train: D/F unsupported");
  TMR_CHECK_ARG(math_mode == TMR_MATH_FP32 || math_mode == TMR_MATH_F16, "train: unknown math_mode %d", math_mode);
  if (math_mode == TMR_MATH_F16 && !umma_available())
    return set_error(TMR_ERR_UNSUPPORTED, "train: TMR_MATH_F16 needs the tcgen05 kernels (sm_100a device + build)");
  const bool tcm = math_mode == TMR_MATH_F16;
  TMR_CHECK_ARG(B >= 1 && seq >= 1 && L >= 1 && L <= kAttnTrainMaxL && C >= 1 && C <= 16, "train: bad sizes (L <= 512, C <= 16)");
  TMR_CHECK_ARG(params && x && long_feature && logits && workspace, "train: null pointer");
  TMR_CHECK_ARG(!(phase & PH_LOSS) || (labels && loss), "train: loss phase needs labels and loss");
  TMR_CHECK_ARG(!(phase & PH_BWD) || grads, "train: backward phase needs grads");
  TMR_CHECK_ARG(!(phase & PH_BWD) || (phase & PH_LOSS) || dlogits_ext, "train: backward phase needs the loss phase or dlogits");
  const bool do_fwd = phase & PH_FWD, do_loss = phase & PH_LOSS, do_bwd = phase & PH_BWD;
  cudaStream_t st = (cudaStream_t)stream;
  const bool has_tc = params[4] != nullptr;
  const int S = seq;
  const int64_t R = (int64_t)B * L, Rp = pad16(R), Bp = pad16(B), T = (int64_t)S * B, Tp = pad16(T);
  const float scale = (float)0.044194173824159216;
  Ws w{(char*)workspace, workspace_bytes};
#define TAKE(name, n) float* name = w.f(n); TMR_CHECK_ARG(name, "train: workspace too small at " #name)
  const float *Wih = params[0], *Whh = params[1];
  // ---- the whole workspace layout, phase-independent ----
  TAKE(x_tm, (size_t)T * kF); TAKE(bsum, 4 * kD); TAKE(xp, (size_t)T * 4 * kD); TAKE(hh, (size_t)B * 4 * kD);
  TAKE(gates, (size_t)T * 4 * kD); TAKE(cst, (size_t)T * kD); TAKE(hst, (size_t)T * kD);
  float *c3 = nullptr, *c5 = nullptr, *c7 = nullptr, *Ltb = nullptr, *shiftb = nullptr, *wtap = nullptr;
  uint8_t* branch = nullptr;
  if (has_tc) {
    c3 = w.f((size_t)R * kD); c5 = w.f((size_t)R * kD); c7 = w.f((size_t)R * kD); Ltb = w.f((size_t)R * kD);
    shiftb = w.f((size_t)R * kD); branch = (uint8_t*)w.f((size_t)R * kD / 4 + 64); wtap = w.f((size_t)kD * kD);
    TMR_CHECK_ARG(c3 && c5 && c7 && Ltb && shiftb && branch && wtap, "train: workspace too small (timeconv)");
  }
  TAKE(W1T, (size_t)kD * kD); TAKE(W2T, (size_t)kD * kD); TAKE(W3T, (size_t)kD * kD); TAKE(W4T, (size_t)kD * kD);
  TAKE(WhT, (size_t)2 * kD * kD); TAKE(WhhT, (size_t)kD * 4 * kD); TAKE(WcT, (size_t)kD * 16);
  TAKE(q, (size_t)B * kD); TAKE(u, (size_t)B * kD); TAKE(pbuf, (size_t)B * L); TAKE(abar, (size_t)B * kD);
  TAKE(v, (size_t)B * kD); TAKE(xhat, (size_t)B * kD); TAKE(rstd, B); TAKE(nrm, (size_t)B * kD); TAKE(r, (size_t)B * kD);
  TAKE(o, (size_t)B * kD); TAKE(od, (size_t)B * kD); TAKE(m1, (size_t)B * kD); TAKE(cat, (size_t)B * 2 * kD);
  TAKE(z0, (size_t)B * kD); TAKE(z1, (size_t)B * kD); TAKE(m2, (size_t)B * kD); TAKE(z, (size_t)B * kD);
  TAKE(dlog16, (size_t)B * 16);
  TAKE(tA, (size_t)2 * kD * Bp); TAKE(tB, (size_t)2 * kD * Bp); TAKE(tC, (size_t)16 * Bp);
  TAKE(dz, (size_t)B * kD); TAKE(dz0, (size_t)B * kD); TAKE(dcat, (size_t)B * 2 * kD); TAKE(dSt, (size_t)B * kD);
  TAKE(dod, (size_t)B * kD); TAKE(dr, (size_t)B * kD); TAKE(dv, (size_t)B * kD); TAKE(dabar, (size_t)B * kD);
  TAKE(du, (size_t)B * kD); TAKE(dq, (size_t)B * kD); TAKE(dsb, (size_t)B * L); TAKE(tmpBD, (size_t)B * kD);
  float *dLt = nullptr, *d3 = nullptr, *d5 = nullptr, *d7 = nullptr, *dcT = nullptr, *shT = nullptr, *tap = nullptr;
  if (has_tc) {
    dLt = w.f((size_t)R * kD); d3 = w.f((size_t)R * kD); d5 = w.f((size_t)R * kD); d7 = w.f((size_t)R * kD);
    dcT = w.f((size_t)kD * Rp); shT = w.f((size_t)kD * Rp); tap = w.f((size_t)kD * kD);
    TMR_CHECK_ARG(dLt && d3 && d5 && d7 && dcT && shT && tap, "train: workspace too small (timeconv backward)");
  }
  TAKE(dpre, (size_t)T * 4 * kD); TAKE(dc, (size_t)B * kD); TAKE(dh, (size_t)B * kD);
  TAKE(dpT, (size_t)4 * kD * Tp); TAKE(opT, (size_t)kF * Tp);
  // fp16 operand scratch of the tensor-core GEMMs (carved in either mode: the layout is phase- and mode-independent)
  const size_t tc_halves = tc_scratch_halves((size_t)T, (size_t)Tp, (size_t)R, (size_t)Rp);
  TAKE(a16f, tc_halves / 2); TAKE(w16f, tc_halves / 2); TAKE(whh16f, (size_t)4 * kD * kD / 2); TAKE(whhT16f, (size_t)4 * kD * kD / 2);
  half_t *xcat16 = nullptr, *xcatT16 = nullptr, *wcat16 = nullptr, *dcT16 = nullptr; float* dwcat = nullptr;
  if (has_tc) {     // batched-tap TimeConv of TMR_MATH_F16 (carved in either mode)
    xcat16 = reinterpret_cast<half_t*>(w.f((size_t)R * kTaps * kD / 2)); xcatT16 = reinterpret_cast<half_t*>(w.f((size_t)kTaps * kD * Rp / 2));
    wcat16 = reinterpret_cast<half_t*>(w.f((size_t)15 * kD * kD / 2)); dcT16 = reinterpret_cast<half_t*>(w.f((size_t)kD * Rp / 2));
    dwcat = w.f((size_t)kTaps * kD * kD);
    TMR_CHECK_ARG(xcat16 && xcatT16 && wcat16 && dcT16 && dwcat, "train: workspace too small (batched taps)");
  }
  half_t* a16 = reinterpret_cast<half_t*>(a16f); half_t* w16 = reinterpret_cast<half_t*>(w16f);
  half_t* Whh16 = reinterpret_cast<half_t*>(whh16f); half_t* WhhT16 = reinterpret_cast<half_t*>(whhT16f);
  // out[M,N] (ldo) = a[M,K] . w[N,K]^T (+ bias) (+ residual).  TMR_MATH_F16: both operands are rounded to fp16 (one
  // conversion launch each; w16_pre = an operand converted earlier in this call) and multiplied on the tensor cores with
  // fp32 accumulation; shapes the tcgen05 engine does not take (K not a multiple of 64: the 16-wide padded class
  // dimension) stay on the fp32 CUDA-core GEMM.
  auto gemm_nt = [&](const float* a, int64_t lda, const float* wgt, int64_t ldw, const float* bias, float* out, int64_t ldo,
                     int64_t M, int N, int K, cudaStream_t s2, const float* residual = nullptr, int64_t ldr = 0,
                     const half_t* w16_pre = nullptr) -> int {
    LinearArgs g; g.a = a; g.lda = lda; g.w = wgt; g.ldw = ldw; g.bias = bias; g.out = out; g.ldo = ldo; g.M = M; g.N = N; g.K = K;
    g.residual = residual; g.ldr = ldr;
    if (!tcm || K % 64 != 0 || N % 4 != 0 || ldo % 4 != 0) return simt_linear(g, s2);
    TMR_CHECK_ARG((size_t)M * K <= tc_halves && (size_t)N * K <= tc_halves, "train: fp16 scratch too small");
    if (w16_pre || (lda % 4) || (ldw % 4) || !aligned16(a) || !aligned16(wgt)) {
      TMR_TRY(launch_half_concat(a, lda, nullptr, 0, K, K, M, a16, s2));
      if (!w16_pre) TMR_TRY(launch_half_concat(wgt, ldw, nullptr, 0, K, K, N, w16, s2));
    } else {                                   // both operands in one launch
      const int64_t n4 = (M + N) * (K / 4);
      half_pair_kernel<<<nblk(n4), 256, 0, s2>>>(a, lda, M, wgt, ldw, N, K, a16, w16);
    }
    g.a16 = a16; g.lda = K; g.w16 = w16_pre ? w16_pre : w16; g.ldw = K;
    return umma_linear(g, s2);
  };
  const float* St = hst + (size_t)(S - 1) * B * kD;
  const float* Lt = has_tc ? Ltb : long_feature;
  float* convs[3] = {c3, c5, c7};

  if (do_fwd) {
    // ---------------- forward: LSTM ----------------
    permute_bsf_kernel<<<nblk(T * kF), 256, 0, st>>>(x, B, S, kF, x_tm);
    add_vec_kernel<<<nblk(4 * kD), 256, 0, st>>>(params[2], params[3], bsum, 4 * kD);
    TMR_TRY(gemm_nt(x_tm, kF, Wih, kF, bsum, xp, 4 * kD, T, 4 * kD, kF, st));
    if (tcm && S > 1) TMR_TRY(launch_half_concat(Whh, kD, nullptr, 0, kD, kD, 4 * kD, Whh16, st));
    // TMR_MATH_F16, batches the co-resident grid holds (<= 512 clips): ALL steps in one launch (umma_lstm_small.cu:
    // c in registers, h exchanged through L2, every step's gates / c / h saved) instead of a conversion, a GEMM and a
    // cell launch per step; same fp16 rounding of h, same K order: the same bits.  Its fp16 exchange buffers and
    // arrival counters sit in the (idle) fp16 operand scratch.
    int rec = TMR_ERR_UNSUPPORTED;
    if (tcm && S > 1 && B <= umma_lstm_small_max_clips() && (size_t)2 * B * kD + 64 <= tc_halves)
      rec = umma_lstm_train_fwd(Whh16, xp, B, S, gates, cst, hst, a16, a16 + (size_t)B * kD,
                                reinterpret_cast<int32_t*>(a16 + (size_t)2 * B * kD), st);
    if (rec != TMR_OK && rec != TMR_ERR_UNSUPPORTED) return rec;
    for (int t = 0; rec == TMR_ERR_UNSUPPORTED && t < S; ++t) {
      if (t > 0) TMR_TRY(gemm_nt(hst + (size_t)(t - 1) * B * kD, kD, Whh, kD, nullptr, hh, 4 * kD, B, 4 * kD, kD, st, nullptr, 0,
                                 tcm ? Whh16 : nullptr));
      lstm_cell_fwd_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(xp + (size_t)t * B * 4 * kD, t > 0 ? hh : nullptr,
                                                                 t > 0 ? cst + (size_t)(t - 1) * B * kD : nullptr,
                                                                 gates + (size_t)t * B * 4 * kD, cst + (size_t)t * B * kD,
                                                                 hst + (size_t)t * B * kD, B);
    }
    // ---------------- forward: TimeConv ----------------
    if (has_tc && tcm) {
      build_xcat16_kernel<<<nblk(R * kTaps * (kD / 4)), 256, 0, st>>>(long_feature, L, R, xcat16);
      size_t woff = 0;
      for (int ci = 0; ci < 3; ++ci) {         // conv_K = bias + xcat[:, taps of K] . wcat_K^T: one GEMM per convolution
        const int K = 3 + 2 * ci, h = K / 2;
        permute_wcat16_kernel<<<nblk((int64_t)kD * kD * K), 256, 0, st>>>(params[4 + 2 * ci], K, wcat16 + woff);
        LinearArgs g; g.a16 = xcat16 + (size_t)(3 - h) * kD; g.lda = kTaps * kD; g.w16 = wcat16 + woff; g.ldw = (int64_t)K * kD;
        g.bias = params[5 + 2 * ci]; g.out = convs[ci]; g.ldo = kD; g.M = R; g.N = kD; g.K = K * kD;
        TMR_TRY(umma_linear(g, st));
        woff += (size_t)kD * kD * K;
      }
      timeconv_max_train_kernel<<<nblk(R * kD), 256, 0, st>>>(long_feature, c3, c5, c7, L, R * kD, Ltb, branch);
    } else if (has_tc) {
      for (int ci = 0; ci < 3; ++ci) {         // conv_K = bias + sum_j shift_j(x) . W_K[:,:,j]^T  (tap matrices gathered from (D,D,K))
        const int K = 3 + 2 * ci, h = K / 2;
        const float* Wk = params[4 + 2 * ci];
        for (int j = 0; j < K; ++j) {
          // wtap[o][c] = Wk[o][c][j]  (strided gather = transpose_pad of a (D*D) x K matrix column j)
          transpose_pad_kernel<<<dim3(1, (unsigned)((kD * kD + 31) / 32)), dim3(32, 8), 0, st>>>(Wk + j, (int64_t)kD * kD, 1, K, wtap, (int64_t)kD * kD);
          shift_rows_kernel<<<nblk(R * kD), 256, 0, st>>>(long_feature, L, j - h, R * kD, shiftb);
          TMR_TRY(gemm_nt(shiftb, kD, wtap, kD, j == 0 ? params[5 + 2 * ci] : nullptr, convs[ci], kD, R, kD, kD, st,
                          j == 0 ? nullptr : convs[ci], kD));
        }
      }
      timeconv_max_train_kernel<<<nblk(R * kD), 256, 0, st>>>(long_feature, c3, c5, c7, L, R * kD, Ltb, branch);
    }
    // ---------------- forward: NLBlock ----------------
    TMR_TRY(transpose_pad(params[12], kD, kD, kD, W2T, kD, st));
    TMR_TRY(gemm_nt(St, kD, params[10], kD, params[11], q, kD, B, kD, kD, st));            // q = St W1^T + b1
    TMR_TRY(gemm_nt(q, kD, W2T, kD, nullptr, u, kD, B, kD, kD, st));                       // u = W2^T q
    attention_train_fwd_kernel<<<B, kAttnTrainWarps * 32, 0, st>>>(u, Lt, B, L, scale, pbuf, abar);
    TMR_TRY(gemm_nt(abar, kD, params[14], kD, params[15], v, kD, B, kD, kD, st));          // v = W3 abar + b3
    layernorm_train_fwd_kernel<<<(B + 3) / 4, 128, 0, st>>>(v, params[18], params[19], B, xhat, rstd, nrm, r);
    TMR_TRY(gemm_nt(r, kD, params[16], kD, params[17], o, kD, B, kD, kD, st));             // o = W4 r + b4
    dropout_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(o, p_nl, seed * 2 + 1, (int64_t)B * kD, od, m1);
    make_cat_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(St, od, B, cat);
    // ---------------- forward: classifier ----------------
    TMR_TRY(gemm_nt(cat, 2 * kD, params[20], 2 * kD, params[21], z0, kD, B, kD, 2 * kD, st));
    dropout_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(z0, p_fc, seed * 2 + 2, (int64_t)B * kD, z1, m2);
    relu_fwd_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(z1, z, (int64_t)B * kD);
    TMR_TRY(gemm_nt(z, kD, params[22], kD, params[23], logits, C, B, C, kD, st));
    TMR_LAUNCH_CHECK("train forward");
  }
  if (do_loss) {
    TMR_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    ce_loss_kernel<<<(B + 127) / 128, 128, 0, st>>>(logits, labels, class_weight, B, C, dlog16, loss, pred);
    TMR_LAUNCH_CHECK("train loss");
  }
  if (!do_bwd) return TMR_OK;
  if (!do_loss) pad_dlogits_kernel<<<nblk((int64_t)B * 16), 256, 0, st>>>(dlogits_ext, B, C, dlog16);
  if (tcm) scale_vec_kernel<<<nblk((int64_t)B * 16), 256, 0, st>>>(dlog16, (int64_t)B * 16, kGradScale);

  // weight transposes the backward GEMMs read (recomputed here: the weights cannot have changed since the forward,
  // autograd's version counters check that)
  TMR_TRY(transpose_pad(params[10], kD, kD, kD, W1T, kD, st));
  TMR_TRY(transpose_pad(params[14], kD, kD, kD, W3T, kD, st));
  TMR_TRY(transpose_pad(params[16], kD, kD, kD, W4T, kD, st));
  TMR_TRY(transpose_pad(params[20], kD, 2 * kD, 2 * kD, WhT, kD, st));       // (D,2D) -> (2D, D)
  TMR_TRY(transpose_pad(Whh, 4 * kD, kD, kD, WhhT, 4 * kD, st));             // (4D,D) -> (D, 4D)
  TMR_TRY(transpose_pad(params[22], C, kD, kD, WcT, 16, st));                // (C,D)  -> (D, 16) zero padded
  // ---------------- backward: classifier ----------------
  // dWc = dlogits^T z ; dbc ; dz = dlogits Wc
  TMR_TRY(transpose_pad(dlog16, B, 16, 16, tC, Bp, st));                                  // (16, Bp)
  TMR_TRY(transpose_pad(z, B, kD, kD, tA, Bp, st));                                       // (D, Bp)
  TMR_TRY(gemm_nt(tC, Bp, tA, Bp, nullptr, grads[22], kD, C, kD, (int)Bp, st));
  TMR_TRY(colsum(dlog16, nullptr, B, C, 16, grads[23], 0, st));
  TMR_TRY(gemm_nt(dlog16, 16, WcT, 16, nullptr, dz, kD, B, kD, 16, st));
  relu_bwd_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(dz, z1, dz0, (int64_t)B * kD);
  mul_mask_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(dz0, m2, dz0, (int64_t)B * kD);
  // dWh = dz0^T cat ; dbh ; dcat = dz0 Wh
  TMR_TRY(transpose_pad(dz0, B, kD, kD, tA, Bp, st));                                     // (D, Bp)
  TMR_TRY(transpose_pad(cat, B, 2 * kD, 2 * kD, tB, Bp, st));                             // (2D, Bp)
  TMR_TRY(gemm_nt(tA, Bp, tB, Bp, nullptr, grads[20], 2 * kD, kD, 2 * kD, (int)Bp, st));
  TMR_TRY(colsum(dz0, nullptr, B, kD, kD, grads[21], 0, st));
  TMR_TRY(gemm_nt(dz0, kD, WhT, kD, nullptr, dcat, 2 * kD, B, 2 * kD, kD, st));
  split_dcat_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(dcat, B, dSt, dod);
  // ---------------- backward: NLBlock ----------------
  mul_mask_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(dod, m1, dod, (int64_t)B * kD);   // do
  TMR_TRY(transpose_pad(dod, B, kD, kD, tA, Bp, st));
  TMR_TRY(transpose_pad(r, B, kD, kD, tB, Bp, st));
  TMR_TRY(gemm_nt(tA, Bp, tB, Bp, nullptr, grads[16], kD, kD, kD, (int)Bp, st));           // dW4 = do^T r
  TMR_TRY(colsum(dod, nullptr, B, kD, kD, grads[17], 0, st));
  TMR_TRY(gemm_nt(dod, kD, W4T, kD, nullptr, dr, kD, B, kD, kD, st));                      // dr = do W4
  relu_bwd_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(dr, nrm, dr, (int64_t)B * kD);     // dn
  TMR_TRY(colsum(dr, xhat, B, kD, kD, grads[18], 0, st));                                   // dgamma
  TMR_TRY(colsum(dr, nullptr, B, kD, kD, grads[19], 0, st));                                // dbeta
  layernorm_train_bwd_kernel<<<(B + 3) / 4, 128, 0, st>>>(dr, xhat, rstd, params[18], B, dv);
  TMR_TRY(transpose_pad(dv, B, kD, kD, tA, Bp, st));
  TMR_TRY(transpose_pad(abar, B, kD, kD, tB, Bp, st));
  TMR_TRY(gemm_nt(tA, Bp, tB, Bp, nullptr, grads[14], kD, kD, kD, (int)Bp, st));           // dW3 = dv^T abar
  TMR_TRY(colsum(dv, nullptr, B, kD, kD, grads[15], 0, st));
  TMR_TRY(gemm_nt(dv, kD, W3T, kD, nullptr, dabar, kD, B, kD, kD, st));                    // dabar = dv W3
  attention_train_bwd_kernel<<<B, kAttnTrainWarps * 32, 0, st>>>(u, Lt, pbuf, dabar, B, L, scale, dsb, du, dLt);
  // u = W2^T q : dq = du W2^T (out[b,i] = sum_j du[b,j] W2[i][j]) ; dW2 = q^T du ; db2 = 0
  TMR_TRY(gemm_nt(du, kD, params[12], kD, nullptr, dq, kD, B, kD, kD, st));
  TMR_TRY(transpose_pad(q, B, kD, kD, tA, Bp, st));
  TMR_TRY(transpose_pad(du, B, kD, kD, tB, Bp, st));
  TMR_TRY(gemm_nt(tA, Bp, tB, Bp, nullptr, grads[12], kD, kD, kD, (int)Bp, st));
  zero_kernel<<<2, 256, 0, st>>>(grads[13], kD);
  // q = St W1^T + b1 : dW1 = dq^T St ; db1 ; dSt += dq W1
  TMR_TRY(transpose_pad(dq, B, kD, kD, tA, Bp, st));
  TMR_TRY(transpose_pad(St, B, kD, kD, tB, Bp, st));
  TMR_TRY(gemm_nt(tA, Bp, tB, Bp, nullptr, grads[10], kD, kD, kD, (int)Bp, st));
  TMR_TRY(colsum(dq, nullptr, B, kD, kD, grads[11], 0, st));
  TMR_TRY(gemm_nt(dq, kD, W1T, kD, nullptr, tmpBD, kD, B, kD, kD, st));
  axpy_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(tmpBD, dSt, (int64_t)B * kD);
  // ---------------- backward: TimeConv (weights only; the bank has no gradient) ----------------
  if (has_tc && tcm) {
    conv_route_kernel<<<nblk(R * kD), 256, 0, st>>>(dLt, branch, R * kD, d3, d5, d7);
    float* dcs[3] = {d3, d5, d7};
    // xcat may come from a forward in the other math mode (or not at all): rebuild it, it is one launch
    build_xcat16_kernel<<<nblk(R * kTaps * (kD / 4)), 256, 0, st>>>(long_feature, L, R, xcat16);
    TMR_TRY(transpose_pad16<half_t>(xcat16, R, kTaps * kD, kTaps * kD, xcatT16, Rp, st));          // (7*512, Rp)
    for (int ci = 0; ci < 3; ++ci) {
      const int K = 3 + 2 * ci, h = K / 2;
      TMR_TRY(colsum(dcs[ci], nullptr, R, kD, kD, grads[5 + 2 * ci], 0, st));
      TMR_TRY(transpose_pad16<float>(dcs[ci], R, kD, kD, dcT16, Rp, st));                          // (D_out, Rp)
      LinearArgs g; g.a16 = dcT16; g.lda = Rp; g.w16 = xcatT16 + (size_t)(3 - h) * kD * Rp; g.ldw = Rp;
      g.out = dwcat; g.ldo = (int64_t)K * kD; g.M = kD; g.N = K * kD; g.K = (int)Rp;
      TMR_TRY(umma_linear(g, st));                                                                  // dcat[o][j*512 + c]
      permute_dwcat_kernel<<<nblk((int64_t)kD * kD * K), 256, 0, st>>>(dwcat, K, grads[4 + 2 * ci]);
    }
  } else if (has_tc) {
    conv_route_kernel<<<nblk(R * kD), 256, 0, st>>>(dLt, branch, R * kD, d3, d5, d7);
    float* dcs[3] = {d3, d5, d7};
    for (int ci = 0; ci < 3; ++ci) {
      const int K = 3 + 2 * ci, h = K / 2;
      TMR_TRY(colsum(dcs[ci], nullptr, R, kD, kD, grads[5 + 2 * ci], 0, st));
      TMR_TRY(transpose_pad(dcs[ci], R, kD, kD, dcT, Rp, st));                              // (D_out, Rp)
      for (int j = 0; j < K; ++j) {
        shift_rows_kernel<<<nblk(R * kD), 256, 0, st>>>(long_feature, L, j - h, R * kD, shiftb);
        TMR_TRY(transpose_pad(shiftb, R, kD, kD, shT, Rp, st));                             // (D_in, Rp)
        TMR_TRY(gemm_nt(dcT, Rp, shT, Rp, nullptr, tap, kD, kD, kD, (int)Rp, st));          // tap[o][c]
        scatter_tap_kernel<<<nblk((int64_t)kD * kD), 256, 0, st>>>(tap, K, j, grads[4 + 2 * ci]);
      }
    }
  }
  // ---------------- backward: LSTM (BPTT) ----------------
  zero_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(dc, (int64_t)B * kD);
  copy_vec_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(dSt, dh, (int64_t)B * kD);
  if (tcm && S > 1) TMR_TRY(launch_half_concat(WhhT, 4 * kD, nullptr, 0, 4 * kD, 4 * kD, kD, WhhT16, st));
  for (int t = S - 1; t >= 0; --t) {
    lstm_cell_bwd_kernel<<<nblk((int64_t)B * kD), 256, 0, st>>>(dh, dc, gates + (size_t)t * B * 4 * kD, cst + (size_t)t * B * kD,
                                                               t > 0 ? cst + (size_t)(t - 1) * B * kD : nullptr,
                                                               dpre + (size_t)t * B * 4 * kD, B);
    if (t > 0) TMR_TRY(gemm_nt(dpre + (size_t)t * B * 4 * kD, 4 * kD, WhhT, 4 * kD, nullptr, dh, kD, B, kD, 4 * kD, st, nullptr, 0,
                               tcm ? WhhT16 : nullptr));   // dh_{t-1} = dpre_t Whh
  }
  TMR_TRY(transpose_pad(dpre, T, 4 * kD, 4 * kD, dpT, Tp, st));                              // (4D, Tp)
  TMR_TRY(transpose_pad(x_tm, T, kF, kF, opT, Tp, st));                                      // (F, Tp)
  TMR_TRY(gemm_nt(dpT, Tp, opT, Tp, nullptr, grads[0], kF, 4 * kD, kF, (int)Tp, st));        // dWih = dpre^T X
  TMR_TRY(colsum(dpre, nullptr, T, 4 * kD, 4 * kD, grads[2], 0, st));
  copy_vec_kernel<<<nblk(4 * kD), 256, 0, st>>>(grads[2], grads[3], 4 * kD);
  if (S > 1) {   // dWhh = sum_{t>=1} dpre_t^T h_{t-1}: rows B.. of dpre against rows 0..T-B of the h history
    const int64_t T1 = T - B, T1p = pad16(T1);
    TMR_TRY(transpose_pad(dpre + (size_t)B * 4 * kD, T1, 4 * kD, 4 * kD, dpT, T1p, st));
    TMR_TRY(transpose_pad(hst, T1, kD, kD, opT, T1p, st));
    TMR_TRY(gemm_nt(dpT, T1p, opT, T1p, nullptr, grads[1], kD, 4 * kD, kD, (int)T1p, st));
  } else {
    zero_kernel<<<nblk((int64_t)4 * kD * kD), 256, 0, st>>>(grads[1], (int64_t)4 * kD * kD);
  }
  if (tcm) {
    GradTable gt{};
    const int64_t sizes[24] = {(int64_t)4 * kD * kF, (int64_t)4 * kD * kD, 4 * kD, 4 * kD,
                               (int64_t)kD * kD * 3, kD, (int64_t)kD * kD * 5, kD, (int64_t)kD * kD * 7, kD,
                               (int64_t)kD * kD, kD, (int64_t)kD * kD, kD, (int64_t)kD * kD, kD, (int64_t)kD * kD, kD, kD, kD,
                               (int64_t)kD * 2 * kD, kD, (int64_t)C * kD, C};
    for (int i = 0; i < 24; ++i) { gt.p[i] = (has_tc || i < 4 || i >= 10) ? grads[i] : nullptr; gt.n[i] = sizes[i]; }
    unscale_grads_kernel<<<dim3(296, 24), 256, 0, st>>>(gt, 1.f / kGradScale);
  }
  TMR_LAUNCH_CHECK("train backward");
#undef TAKE
  return TMR_OK;
}

extern "C" {

/* One call: forward (training mode) + CrossEntropyLoss(sum[, weight]) + backward.  params/grads: 24 device pointers in
 * the order documented at the top of train.cu (timeconv entries may be NULL for the NL-only wiring).
 * x (B,seq,F) features, long_feature (B,L,D), labels int64[B], class_weight float[C] or NULL.
 * Dropout p_nl (NLBlock, 0.2 in the reference) and p_fc (0.5) with a counter-based mask from `seed`.
 * Outputs: grads (overwritten), logits (B,C), loss (1 float, sum-reduced), pred int64[B] (nullable). */
int tmr_head_train_fwd_bwd(const float* const* params, float* const* grads, const float* x, const float* long_feature,
                           const int64_t* labels, const float* class_weight, int B, int seq, int L, int F, int D, int C,
                           float p_nl, float p_fc, uint64_t seed, float* logits, float* loss, int64_t* pred,
                           void* workspace, size_t workspace_bytes, int math_mode, void* stream) {
  return train_impl(PH_FWD | PH_LOSS | PH_BWD, params, grads, x, long_feature, labels, class_weight, nullptr, B, seq, L, F, D, C,
                    p_nl, p_fc, seed, logits, loss, pred, workspace, workspace_bytes, math_mode, stream);
}

/* The same step split at the logits, for torch.autograd (tmrnet_b200.modules: model.train(); out = model(x, lf);
 * loss = criterion(out, y); loss.backward(); optimizer.step() - the reference's loop body, TRAIN:876-887):
 * tmr_head_train_fwd saves its activations in `workspace`; tmr_head_train_bwd, given the SAME workspace (untouched
 * in between), the same params / x / long_feature and dlogits (B,C), overwrites grads. */
int tmr_head_train_fwd(const float* const* params, const float* x, const float* long_feature, int B, int seq, int L, int F,
                       int D, int C, float p_nl, float p_fc, uint64_t seed, float* logits, void* workspace,
                       size_t workspace_bytes, int math_mode, void* stream) {
  return train_impl(PH_FWD, params, nullptr, x, long_feature, nullptr, nullptr, nullptr, B, seq, L, F, D, C, p_nl, p_fc, seed,
                    logits, nullptr, nullptr, workspace, workspace_bytes, math_mode, stream);
}
int tmr_head_train_bwd(const float* const* params, float* const* grads, const float* x, const float* long_feature,
                       const float* dlogits, int B, int seq, int L, int F, int D, int C, float* logits_scratch,
                       void* workspace, size_t workspace_bytes, int math_mode, void* stream) {
  TMR_CHECK_ARG(dlogits, "train_bwd: dlogits is null");
  return train_impl(PH_BWD, params, grads, x, long_feature, nullptr, nullptr, dlogits, B, seq, L, F, D, C, 0.f, 0.f, 0,
                    logits_scratch, nullptr, nullptr, workspace, workspace_bytes, math_mode, stream);
}

/* torch.optim.SGD step on one flat tensor (momentum mu, weight decay wd, dampening 0): see sgd_kernel. */
int tmr_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float mu, float wd,
                 int first_step, void* stream) {
  TMR_CHECK_ARG(param && grad && momentum_buf && n >= 0, "sgd: bad arguments");
  if (n == 0) return TMR_OK;
  sgd_kernel<<<nblk(n), 256, 0, (cudaStream_t)stream>>>(param, grad, momentum_buf, n, lr, mu, wd, first_step);
  TMR_LAUNCH_CHECK("sgd_kernel");
  return TMR_OK;
}

/* The same update over `count` (<= 24) tensors in one launch: params / grads / momentum_bufs are HOST arrays of device
 * pointers (NULL entries are skipped), sizes and lrs host arrays; element-wise identical to `count` tmr_sgd_step calls. */
int tmr_sgd_step_multi(float* const* params, const float* const* grads, float* const* momentum_bufs, const int64_t* sizes,
                       const float* lrs, int count, float mu, float wd, int first_step, void* stream) {
  TMR_CHECK_ARG(params && grads && momentum_bufs && sizes && lrs && count >= 0 && count <= 24, "sgd_multi: bad arguments");
  if (count == 0) return TMR_OK;
  SgdTable t{};
  for (int i = 0; i < count; ++i) {
    TMR_CHECK_ARG(!params[i] || (grads[i] && momentum_bufs[i] && sizes[i] >= 0), "sgd_multi: tensor %d incomplete", i);
    t.p[i] = params[i]; t.g[i] = grads[i]; t.buf[i] = momentum_bufs[i]; t.n[i] = params[i] ? sizes[i] : 0; t.lr[i] = lrs[i];
  }
  sgd_multi_kernel<<<dim3(296, (unsigned)count), 256, 0, (cudaStream_t)stream>>>(t, mu, wd, first_step);
  TMR_LAUNCH_CHECK("sgd_multi_kernel");
  return TMR_OK;
}

}  // extern "C"
