// TMR_MATH_F16 TimeConv (NLB:43-79): the three zero-padded temporal convolutions (k=3/5/7,
// 512->512) as ONE implicit GEMM on tcgen05 with a fused 5-way-max epilogue.
//
// Tile: up to 128 window rows (nb = 128/L whole windows when L <= 128, else a 128-slot chunk of one
// window) x 128 output channels.  Three fp32 accumulators (conv3/conv5/conv7, 128 columns each) live
// in TMEM.  The K loop walks (input-channel chunk of 64 fp16) x (time shift d = -3..3): for each pair one
// TMA load brings the window rows shifted by d — a rank-3 tensor map (D, L, B) whose out-of-bounds
// fill supplies the zero "same" padding at both window edges — plus the weight tap tiles of every
// conv that has that shift (3, 2 or 1 tiles).  One elected thread issues the tcgen05.mma.kind::f16
// instructions; four epilogue warps then read the accumulators with tcgen05.ld, add the biases and
// take max(conv3, conv5, conv7, x[k], k>0 ? x[k-1] : 0).
#include "tmr_internal.h"
#include "umma_common.cuh"

namespace tmr {
namespace umma {

constexpr int TC_BM = 128;          // window rows per tile (TMEM lanes)
constexpr int TC_BN = 128;          // output channels per tile
constexpr int TC_BK = 64;           // fp16 input channels per k-block (128-byte swizzle row)
constexpr int TC_STAGES = 3;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;            // 16 KB
constexpr int TC_W_BYTES = TC_BN * TC_BK * 2;            // 16 KB per tap tile
constexpr int TC_STAGE_BYTES = TC_A_BYTES + 3 * TC_W_BYTES;   // 64 KB
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 + 256;
constexpr int TC_THREADS = 192;
constexpr int TC_TMEM_COLS = 512;   // 3 x 128 used

struct TimeConvParams {
  const float* x; float* out; const float* bias3; const float* bias5; const float* bias7;
  int B; int L; int box_l; int nb; int l_chunks;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
umma_timeconv_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w3,
                     const __grid_constant__ CUtensorMap tma_w5, const __grid_constant__ CUtensorMap tma_w7,
                     const TimeConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the __shared__ array (no integer round trip) keeps the shared address space, so the
  // epilogue staging compiles to STS/LDS instead of generic ST.E/LD.E
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + TC_STAGES;
  uint64_t* acc_full = bars + 2 * TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int N_TILES = kD / TC_BN;
  const int n0 = (blockIdx.x % N_TILES) * TC_BN;             // output-channel tile (fastest: A tile shared in L2)
  const int row_tile = blockIdx.x / N_TILES;
  const int grp = row_tile / p.l_chunks;                     // clip group
  const int lc = row_tile % p.l_chunks;                      // 128-slot chunk of the window (L > 128)
  const int b0 = grp * p.nb;
  const int k_lo = lc * TC_BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_x); tma_prefetch_desc(&tma_w3); tma_prefetch_desc(&tma_w5); tma_prefetch_desc(&tma_w7);
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t x_bytes = (uint32_t)(p.nb * p.box_l * TC_BK * 2);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int chunk = 0; chunk < kD / TC_BK; ++chunk) {
        const int c0 = chunk * TC_BK;
        for (int d = -3; d <= 3; ++d) {
          const int ad = d < 0 ? -d : d;
          const int n_w = (ad <= 1) ? 3 : (ad == 2 ? 2 : 1);       // convs that own this shift
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * TC_STAGE_BYTES;
          uint8_t* sw = sa + TC_A_BYTES;
          mbar_expect_tx(&full_bar[stage], x_bytes + n_w * TC_W_BYTES);
          // window rows k_lo+d .. : slots outside [0,L) are zero-filled by TMA (conv zero padding)
          tma_load_3d(sa, &tma_x, &full_bar[stage], c0, k_lo + d, b0);
          // packed weights: row = out channel, col = tap*512 + in channel; tap index = d + half
          tma_load_2d(sw + 0 * TC_W_BYTES, &tma_w7, &full_bar[stage], (d + 3) * kD + c0, n0);
          if (n_w >= 2) tma_load_2d(sw + 1 * TC_W_BYTES, &tma_w5, &full_bar[stage], (d + 2) * kD + c0, n0);
          if (n_w >= 3) tma_load_2d(sw + 2 * TC_W_BYTES, &tma_w3, &full_bar[stage], (d + 1) * kD + c0, n0);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(TC_BM, TC_BN);
      int stage = 0; uint32_t phase = 0;
      for (int chunk = 0; chunk < kD / TC_BK; ++chunk) {
        for (int d = -3; d <= 3; ++d) {
          const int ad = d < 0 ? -d : d;
          const int n_w = (ad <= 1) ? 3 : (ad == 2 ? 2 : 1);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * TC_STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(sa);
#pragma unroll
          for (int w = 0; w < 3; ++w) {
            if (w < n_w) {
              // w = 0: conv7 (TMEM cols 256..383), 1: conv5 (128..255), 2: conv3 (0..127)
              const int half = 3 - w;
              const uint32_t d_tmem = tmem_base + (uint32_t)((2 - w) * TC_BN);
              const uint64_t db = make_smem_desc_sw128(sa + TC_A_BYTES + w * TC_W_BYTES);
              const bool first = (chunk == 0) && (d == -half);       // first contribution to this accumulator
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k)
                mma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, !(first && k == 0));
            }
          }
          mma_commit(&empty_bar[stage]);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      mma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;                                   // tile row = TMEM lane
    const int bb = r / p.box_l;
    const int kk = r - bb * p.box_l;
    const int b = b0 + bb;
    const int k = k_lo + kk;
    const bool valid = (bb < p.nb) && (b < p.B) && (k < p.L);
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const int64_t m = (int64_t)b * p.L + k;
#pragma unroll 1
    for (int cc = 0; cc < TC_BN; cc += 32) {
      uint32_t r3[32], r5[32], r7[32];
      tmem_ld32(t_row + 0 * TC_BN + cc, r3);
      tmem_ld32(t_row + 1 * TC_BN + cc, r5);
      tmem_ld32(t_row + 2 * TC_BN + cc, r7);
      tmem_ld_wait();
      if (valid) {
        const int n = n0 + cc;
        const float* xc = p.x + m * kD + n;
        float* dst = p.out + m * kD + n;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b3 = __ldg(reinterpret_cast<const float4*>(p.bias3 + n + j));
          const float4 b5 = __ldg(reinterpret_cast<const float4*>(p.bias5 + n + j));
          const float4 b7 = __ldg(reinterpret_cast<const float4*>(p.bias7 + n + j));
          const float4 x0 = __ldg(reinterpret_cast<const float4*>(xc + j));
          // F.pad(x,(1,0)) + MaxPool1d(2,1): the zero pad takes part in the max at k == 0
          const float4 x1 = (k > 0) ? __ldg(reinterpret_cast<const float4*>(xc - kD + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 o;
          o.x = fmaxf(fmaxf(fmaxf(__uint_as_float(r3[j + 0]) + b3.x, __uint_as_float(r5[j + 0]) + b5.x), __uint_as_float(r7[j + 0]) + b7.x), fmaxf(x0.x, x1.x));
          o.y = fmaxf(fmaxf(fmaxf(__uint_as_float(r3[j + 1]) + b3.y, __uint_as_float(r5[j + 1]) + b5.y), __uint_as_float(r7[j + 1]) + b7.y), fmaxf(x0.y, x1.y));
          o.z = fmaxf(fmaxf(fmaxf(__uint_as_float(r3[j + 2]) + b3.z, __uint_as_float(r5[j + 2]) + b5.z), __uint_as_float(r7[j + 2]) + b7.z), fmaxf(x0.z, x1.z));
          o.w = fmaxf(fmaxf(fmaxf(__uint_as_float(r3[j + 3]) + b3.w, __uint_as_float(r5[j + 3]) + b5.w), __uint_as_float(r7[j + 3]) + b7.w), fmaxf(x0.w, x1.w));
          *reinterpret_cast<float4*>(dst + j) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TC_TMEM_COLS); }
}

}  // namespace umma

int umma_timeconv(const float* packed, const float* x, const half_t* x16, int B, int L, float* out, cudaStream_t st) {
  using namespace umma;
  if (B == 0) return TMR_OK;
  TimeConvParams p{};
  p.x = x; p.out = out;
  p.bias3 = packed + TimeConvPacked::b3_off; p.bias5 = packed + TimeConvPacked::b5_off; p.bias7 = packed + TimeConvPacked::b7_off;
  p.B = B; p.L = L;
  p.box_l = L < TC_BM ? L : TC_BM;
  p.nb = TC_BM / p.box_l;
  p.l_chunks = (L + TC_BM - 1) / TC_BM;
  const int groups = (B + p.nb - 1) / p.nb;

  CUtensorMap tx, tw3, tw5, tw7;
  {
    uint64_t dims[3] = {(uint64_t)kD, (uint64_t)L, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)kD * 2, (uint64_t)L * kD * 2};
    uint32_t box[3] = {TC_BK, (uint32_t)p.box_l, (uint32_t)p.nb};
    TMR_TRY(make_tmap(&tx, x16, 3, dims, str, box, 2));
    const half_t* pr = mirror16<TimeConvPacked>(packed);        // fp16 weight mirror
    const half_t* w[3] = {pr + TimeConvPacked::w3_off, pr + TimeConvPacked::w5_off, pr + TimeConvPacked::w7_off};
    CUtensorMap* tw[3] = {&tw3, &tw5, &tw7};
    for (int i = 0; i < 3; ++i) {
      const int taps = 3 + 2 * i;
      uint64_t dw[2] = {(uint64_t)taps * kD, (uint64_t)kD};
      uint64_t sw[1] = {(uint64_t)taps * kD * 2};
      uint32_t bw[2] = {TC_BK, TC_BN};
      TMR_TRY(make_tmap(tw[i], w[i], 2, dw, sw, bw, 2));
    }
  }
  TMR_CUDA(cudaFuncSetAttribute(umma_timeconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  const int64_t tiles = (int64_t)groups * p.l_chunks * (kD / TC_BN);
  TMR_CHECK_ARG(tiles < (int64_t)INT32_MAX, "timeconv: batch too large");
  umma_timeconv_kernel<<<(unsigned)tiles, TC_THREADS, TC_SMEM_BYTES, st>>>(tx, tw3, tw5, tw7, p);
  TMR_LAUNCH_CHECK("umma_timeconv_kernel");
  return TMR_OK;
}

}  // namespace tmr
