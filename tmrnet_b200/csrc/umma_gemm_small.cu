// out[M <= 128, N] = A[M,K] . W[N,K]^T (+ bias) (+ residual) (relu): the GEMMs of a training step (40 clips) and of
// every other caller with at most one 128-row tile.  The persistent 256 x 256 engine of umma_gemm.cu gives such a problem
// ONE or two busy CTA pairs that walk all of K alone: 12-17 us of pipeline set-up and serial k-steps for 20-170 MFLOP.
// Here the problem is cut by OUTPUT COLUMN: N / 16 CTAs, each multiplies the whole (zero-filled) 128-row A tile by its
// 16 weight rows - 4 tcgen05.mma (M = 128, N = 16, K = 16) per 64-wide k-block, k-blocks through an 8-stage TMA ring
// (18 KB per stage, all of K = 512 in flight at once) - and writes its 16 columns: thread = row, bias / residual /
// relu in the order of the large engine.  Same fp16 operands, same K order, fp32 accumulate.
#include "tmr_internal.h"
#include "umma_common.cuh"

namespace tmr {
namespace umma {

constexpr int GS_BM = 128, GS_BN = 16, GS_BK = 64;
constexpr int GS_A_BYTES = GS_BM * GS_BK * 2;          // 16 KB
constexpr int GS_W_BYTES = GS_BN * GS_BK * 2;          //  2 KB
constexpr int GS_STAGE = GS_A_BYTES + GS_W_BYTES;      // 18 KB (a multiple of 1024: every tile stays swizzle-aligned)
constexpr int GS_STAGES = 8;
constexpr int GS_SMEM = GS_STAGES * GS_STAGE + 1024 + 256;
constexpr int GS_THREADS = 64 + 128;

struct GemmSmallParams {
  int M, N, K;
  const float* bias; const float* residual; int64_t ldr; float* out; int64_t ldo; int relu;
};

__global__ void __launch_bounds__(GS_THREADS, 1)
umma_gemm_small_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w,
                       const GemmSmallParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GS_STAGES * GS_STAGE);
  uint64_t* full = bars;                      // [GS_STAGES] TMA -> MMA
  uint64_t* empty = full + GS_STAGES;         // [GS_STAGES] MMA -> TMA
  uint64_t* acc_full = empty + GS_STAGES;     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * GS_BN;
  const int kb_total = p.K / GS_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_w);
    for (int s = 0; s < GS_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * GS_STAGE;
        mbar_expect_tx(&full[stage], GS_STAGE);
        tma_load_2d(sa, &tma_a, &full[stage], kb * GS_BK, 0);                   // rows past M: zero fill
        tma_load_2d(sa + GS_A_BYTES, &tma_w, &full[stage], kb * GS_BK, n0);
        if (++stage == GS_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(GS_BM, GS_BN);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t da = make_smem_desc_sw128(smem_u32(smem + stage * GS_STAGE));
        const uint64_t db = make_smem_desc_sw128(smem_u32(smem + stage * GS_STAGE + GS_A_BYTES));
#pragma unroll
        for (int k = 0; k < GS_BK / 16; ++k)
          mma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        mma_commit(&empty[stage]);
        if (++stage == GS_STAGES) { stage = 0; phase ^= 1; }
      }
      mma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;                             // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16), r);
    tmem_ld_wait_dep16(r);
    if (row < p.M) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + 4 * j;
        if (n < p.N) {                                  // N % 4 == 0
          float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                 __uint_as_float(r[4 * j + 3]));
          if (p.bias) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          }
          if (p.residual) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(p.residual + (int64_t)row * p.ldr + n));
            v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
          }
          if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          *reinterpret_cast<float4*>(p.out + (int64_t)row * p.ldo + n) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 32); }
}

}  // namespace umma

// M <= 128.  Operand contracts of umma_linear (K % 64 == 0, N % 4 == 0, lda / ldw % 8 == 0, ldo % 4 == 0).
int umma_linear_small(const LinearArgs& g, cudaStream_t st) {
  using namespace umma;
  GemmSmallParams p{};
  p.M = (int)g.M; p.N = g.N; p.K = g.K; p.bias = g.bias; p.residual = g.residual; p.ldr = g.ldr; p.out = g.out; p.ldo = g.ldo;
  p.relu = g.relu;
  CUtensorMap ta, tw;
  {
    uint64_t da[2] = {(uint64_t)g.K, (uint64_t)g.M};
    uint64_t sa[1] = {(uint64_t)g.lda * 2};
    uint32_t ba[2] = {GS_BK, GS_BM};
    TMR_TRY(make_tmap(&ta, g.a16, 2, da, sa, ba, 2));
    uint64_t dw[2] = {(uint64_t)g.K, (uint64_t)g.N};
    uint64_t sw[1] = {(uint64_t)g.ldw * 2};
    uint32_t bw[2] = {GS_BK, GS_BN};
    TMR_TRY(make_tmap(&tw, g.w16, 2, dw, sw, bw, 2));
  }
  // (per call, not once per process: the attribute is per device, and one process may drive several)
  TMR_CUDA(cudaFuncSetAttribute(umma_gemm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM));
  const int grid = (g.N + GS_BN - 1) / GS_BN;
  umma_gemm_small_kernel<<<grid, GS_THREADS, GS_SMEM, st>>>(ta, tw, p);
  TMR_LAUNCH_CHECK("umma_gemm_small_kernel");
  return TMR_OK;
}

}  // namespace tmr
