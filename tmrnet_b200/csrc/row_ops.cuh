// Row operations of the head shared by the stand-alone HBM-bound kernels (kernels_mem.cu: one warp per 512-float row)
// and the fused small-batch relation + classifier kernel (umma_head_tail.cu), so both compute the same bits:
// attention over the L memory slots, LayerNorm + ReLU, the C-way FC + softmax score + argmax.
// kCoherent = true: the row was written EARLIER IN THE SAME LAUNCH by another CTA (the caller has acquired the flag
// that publishes it): read it through L2 (ld.global.cg), never through the read-only / L1 path.
#pragma once
#include "tmr_internal.h"

namespace tmr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// streaming 128-bit load/store: the bank is read through the read-only path; gathered windows are
// written once and not re-read by this kernel.
__device__ __forceinline__ float4 ldg_nc(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_na(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool kCoherent>
__device__ __forceinline__ float4 ld_row4(const float4* p) {
  if constexpr (kCoherent) {
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
  } else {
    return __ldg(p);
  }
}

// -------------------------------------------------------------------------------------------
// Attention over the memory slots (NLB:30-34 with phi/g folded):
//   s_k = scale * (u . Lt_k),  p = softmax_k(s),  a = sum_k p_k Lt_k.
// One warp per clip; each lane owns 16 channels (4 float4); slots are streamed once in chunks of
// KB rows with an online (running-max) softmax, dots reduced with warp shuffles.
// -------------------------------------------------------------------------------------------
constexpr int kAttnWarps = 4;
constexpr int KB = 6;

// Packed fp32 pairs (FFMA2 / FMUL2 on sm_100): the attention kernels are bound by instruction issue, and the two
// passes over a slot (dot product, weighted accumulation) are pure FMA streams - one instruction per two lanes of
// a float4 halves them.  Packing a pair of adjacent registers is free.
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// Body shared by the two attention kernels; rowptr(k) yields the 512-float row of memory slot k.
template <bool kCoherent = false, class RowPtr>
__device__ __forceinline__ void attention_body(const float* __restrict__ u, int b, int L, float scale,
                                               void* __restrict__ a, int half_out, RowPtr rowptr) {
  const int lane = threadIdx.x & 31;
  float4 uq[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) uq[i] = ld_row4<kCoherent>(reinterpret_cast<const float4*>(u + (int64_t)b * kD) + i * 32 + lane);
  uint64_t u2[8], acc2[8];                              // (x,y) and (z,w) pairs of the lane's four float4
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    u2[2 * i] = pk2(uq[i].x, uq[i].y); u2[2 * i + 1] = pk2(uq[i].z, uq[i].w);
    acc2[2 * i] = acc2[2 * i + 1] = pk2(0.f, 0.f);
  }
  float run_max = -INFINITY, run_sum = 0.f;

  for (int k0 = 0; k0 < L; k0 += KB) {
    float4 x[KB][4];
    float d[KB];
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      if (k0 + kk < L) {
        const float4* row = rowptr(k0 + kk);
#pragma unroll
        for (int i = 0; i < 4; ++i) x[kk][i] = ldg_nc(row + i * 32 + lane);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[kk][i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      uint64_t p2 = pk2(0.f, 0.f);                      // two interleaved partial sums (even / odd channels)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        p2 = ffma2(pk2(x[kk][i].x, x[kk][i].y), u2[2 * i], p2);
        p2 = ffma2(pk2(x[kk][i].z, x[kk][i].w), u2[2 * i + 1], p2);
      }
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
      for (int i = 0; i < 4; ++i) {
        acc2[2 * i] = ffma2(pp, pk2(x[kk][i].x, x[kk][i].y), acc2[2 * i]);
        acc2[2 * i + 1] = ffma2(pp, pk2(x[kk][i].z, x[kk][i].w), acc2[2 * i + 1]);
      }
    }
    run_max = new_max;
  }
  const float inv = 1.f / run_sum;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 o;
    upk2(acc2[2 * i], o.x, o.y); upk2(acc2[2 * i + 1], o.z, o.w);
    o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
    if (half_out) reinterpret_cast<uint2*>(reinterpret_cast<half_t*>(a) + (int64_t)b * kD)[i * 32 + lane] = pack_h4(o);
    else reinterpret_cast<float4*>(reinterpret_cast<float*>(a) + (int64_t)b * kD)[i * 32 + lane] = o;
  }
}


// LayerNorm([1,512]) (biased variance, eps 1e-5) + ReLU (NLB:35-36) of row b by one warp.
template <bool kCoherent>
__device__ __forceinline__ void layernorm_relu_row(const float* __restrict__ v, const float* __restrict__ w,
                                                   const float* __restrict__ bsh, int b, void* __restrict__ y, int half_out) {
  const int lane = threadIdx.x & 31;
  const float4* src = reinterpret_cast<const float4*>(v + (int64_t)b * kD);
  float4 x[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (kCoherent) x[i] = ld_row4<true>(src + i * 32 + lane); else x[i] = src[i * 32 + lane];
    s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
  }
  const float mean = warp_sum(s) * (1.f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a0 = x[i].x - mean, a1 = x[i].y - mean, a2 = x[i].z - mean, a3 = x[i].w - mean;
    q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / kD) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + i * 32 + lane);
    const float4 be = __ldg(reinterpret_cast<const float4*>(bsh) + i * 32 + lane);
    float4 o;
    o.x = fmaxf((x[i].x - mean) * rstd * g.x + be.x, 0.f);
    o.y = fmaxf((x[i].y - mean) * rstd * g.y + be.y, 0.f);
    o.z = fmaxf((x[i].z - mean) * rstd * g.z + be.z, 0.f);
    o.w = fmaxf((x[i].w - mean) * rstd * g.w + be.w, 0.f);
    if (half_out) reinterpret_cast<uint2*>(reinterpret_cast<half_t*>(y) + (int64_t)b * kD)[i * 32 + lane] = pack_h4(o);
    else reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + (int64_t)b * kD)[i * 32 + lane] = o;
  }
}

// fc_c (512 -> C) + Softmax + torch.max (TRAIN:252, EVAL:491-493) of clip b by one warp.
// pred = first index of the maximum logit (ties -> lowest index); score = 1 / sum exp(l - lmax).
template <bool kCoherent>
__device__ __forceinline__ void fc_argmax_row(const float* __restrict__ z, const float* __restrict__ wc,
                                              const float* __restrict__ bc, int b, int C, float* __restrict__ logits,
                                              int64_t* __restrict__ pred, float* __restrict__ score) {
  const int lane = threadIdx.x & 31;
  float4 x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = ld_row4<kCoherent>(reinterpret_cast<const float4*>(z + (int64_t)b * kD) + i * 32 + lane);
  float mine = -INFINITY;                    // lane c keeps logit c (C <= 32)
  for (int c = 0; c < C; ++c) {
    const float4* wr = reinterpret_cast<const float4*>(wc + (int64_t)c * kD);
    float p = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 w = __ldg(wr + i * 32 + lane);
      p = fmaf(x[i].x, w.x, p); p = fmaf(x[i].y, w.y, p); p = fmaf(x[i].z, w.z, p); p = fmaf(x[i].w, w.w, p);
    }
    p = warp_sum(p) + __ldg(bc + c);
    if (lane == c) mine = p;
  }
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

}  // namespace tmr
