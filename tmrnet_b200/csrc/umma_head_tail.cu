// The non-local relation block and the classifier (NLB:25-40, TRAIN:245-252) of up to 512 clips in ONE launch - the
// "fused attention-style kernel" for the reference's own batch sizes (EVAL:470-495 calls the head on 120 clips).
// As separate launches the tail of a 120-clip call is four GEMMs of 63 MFLOP, each a 12-15 us round of pipeline set-up
// for 2 busy CTA pairs, plus five row kernels of 4-8 us: 77 us in which the GPU does ~10 us of work.  Here a
// 128-clip tile is owned by 32 CTAs for the whole chain
//     u = W21 St + bu -> a = softmax_k(scale u.Lt_k) Lt -> v = W3 a + b3 -> r = relu(LN(v)) -> y = St + W4 r + b4
//       -> z = relu(Wh [St || y] + bh) -> logits = Wc z + bc, softmax score, argmax
// with the GEMMs cut by OUTPUT COLUMN and the row operations cut by CLIP:
//   * CTA j keeps rows 16j .. 16j+15 of W21, W3, W4 and Wh resident in shared memory (fp16, 80 KB, one TMA burst at
//     entry) and has room for a whole 128-clip x 512 fp16 operand tile (128 KB, eight TMA boxes on eight mbarriers);
//     a GEMM is 32 tcgen05.mma (M = 128, N = 16, K = 16) into 16 TMEM columns, its epilogue 16 values per thread;
//   * attention / LayerNorm / FC+argmax run one warp per clip, four clips per CTA, with the SAME row code as the
//     stand-alone kernels (row_ops.cuh);
//   * between stages the tile's 32 CTAs exchange their [128][16] column blocks / 4 rows through L2: stores, a named
//     barrier over the CTA's four worker warps, one device-scope fence + arrival on the tile's counter; consumers spin
//     on the counter (the TMA producer before an operand load, lane 0 of each worker warp before a row stage) and
//     read what other CTAs wrote through L2 only (TMA after a proxy fence, ld.global.cg in the row stages).
// The classifier's first half (St . Wh[:, :512]) is issued while the tile waits for y.  Arithmetic - fp16 operands
// rounded once by their producers, fp32 accumulate, bias / residual / relu order - is that of the separate launches.
// Grid = 32 x ceil(B / 128) CTAs, all of which must be resident (cooperative launch outside stream capture).
#include "tmr_internal.h"
#include "umma_common.cuh"
#include "row_ops.cuh"

namespace tmr {
namespace umma {

constexpr int T_BM = 128;                     // clips per tile
constexpr int T_BN = 16;                      // output columns per CTA
constexpr int T_BK = 64;
constexpr int T_KB = kD / T_BK;               // 8 k-blocks per 512-wide operand
constexpr int T_A_BYTES = T_BM * T_BK * 2;    // 16 KB
constexpr int T_W_BYTES = T_BN * T_BK * 2;    //  2 KB
constexpr int T_SLICES = kD / T_BN;           // 32 CTAs per tile
constexpr int T_ROWS = T_BM / T_SLICES;       // 4 clips per CTA in the row stages = one per worker warp
constexpr int T_WKB = 5 * T_KB;               // weight k-blocks: W21, W3, W4 (8 each), Wh (16)
constexpr int T_SMEM = T_WKB * T_W_BYTES + T_KB * T_A_BYTES + 1024 + 256;
constexpr int T_THREADS = 64 + 128;
constexpr int T_MAX_TILES = 4;

struct HeadTailParams {
  int M; int L; int C; int cls;    // cls = 0: relation block only (y1 = St + W4 r + b4 in fp32 -> y1_out)
  const float* St;                 // [M][512] fp32 (residual)
  const float* Lt;                 // [M][L][512] fp32 memory slots
  float* u;                        // [M][512] fp32 scratch: u, then v
  half_t* a16;                     // [M][512] fp16 scratch: a, then r   (tma_a)
  half_t* y16;                     // [M][512] fp16 scratch: fp16(y)     (tma_y)
  float* z;                        // [M][512] fp32 scratch
  float* y1_out;
  const float* bu; const float* b3; const float* b4; const float* lnw; const float* lnb;
  const float* bh; const float* wc; const float* bc;
  float* logits; int64_t* pred; float* score;
  int32_t* flags;                  // [tiles], zero on entry
  float scale;
};

__device__ __forceinline__ void spin_until(const int32_t* f, int target) {
  int v;
  const long long t0 = clock64();
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (v < target && clock64() - t0 > 4000000000LL) __trap();
  } while (v < target);
}

__global__ void __launch_bounds__(T_THREADS, 1)
umma_head_tail_kernel(const __grid_constant__ CUtensorMap tma_st, const __grid_constant__ CUtensorMap tma_a,
                      const __grid_constant__ CUtensorMap tma_y, const __grid_constant__ CUtensorMap tma_w21,
                      const __grid_constant__ CUtensorMap tma_w3, const __grid_constant__ CUtensorMap tma_w4,
                      const __grid_constant__ CUtensorMap tma_wh, const HeadTailParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                                   // [40 k-blocks][16 weight rows][64 fp16]: W21 | W3 | W4 | Wh
  uint8_t* sA = sW + T_WKB * T_W_BYTES;                 // [8][128 clips][64 fp16]: the current operand tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + T_KB * T_A_BYTES);
  uint64_t* a_full = bars;                    // [8]  TMA -> MMA, one completion per operand load
  uint64_t* a_empty = a_full + T_KB;          // [1]  MMA -> TMA
  uint64_t* w_full = a_empty + 1;             // [1]
  uint64_t* acc_full = w_full + 1;            // [1]
  uint64_t* acc_empty = acc_full + 1;         // [1]  four worker warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x / T_SLICES;
  const int slice = blockIdx.x % T_SLICES;
  const int n0 = slice * T_BN;
  const int m0 = tile * T_BM;
  int32_t* flag = p.flags + tile;
  const int n_loads = p.cls ? 5 : 3;                    // operand tiles: St, a, r (, St, y)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_st); tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_y); tma_prefetch_desc(&tma_w21);
    tma_prefetch_desc(&tma_w3); tma_prefetch_desc(&tma_w4); tma_prefetch_desc(&tma_wh);
    for (int k = 0; k < T_KB; ++k) mbar_init(&a_full[k], 1);
    mbar_init(a_empty, 1); mbar_init(w_full, 1); mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, (p.cls ? T_WKB : 3 * T_KB) * T_W_BYTES);
      for (int kb = 0; kb < T_KB; ++kb) {
        tma_load_2d(sW + kb * T_W_BYTES, &tma_w21, w_full, kb * T_BK, n0);
        tma_load_2d(sW + (T_KB + kb) * T_W_BYTES, &tma_w3, w_full, kb * T_BK, n0);
        tma_load_2d(sW + (2 * T_KB + kb) * T_W_BYTES, &tma_w4, w_full, kb * T_BK, n0);
      }
      if (p.cls)
        for (int kb = 0; kb < 2 * T_KB; ++kb) tma_load_2d(sW + (3 * T_KB + kb) * T_W_BYTES, &tma_wh, w_full, kb * T_BK, n0);
      for (int i = 0; i < n_loads; ++i) {
        // what the tile must have published before this operand exists: a after 2 arrivals per CTA (u, a), r after 4
        // (+ v, r), y after 5; St is an input
        const int need = (i == 1) ? 2 : (i == 2) ? 4 : (i == 4) ? 5 : 0;
        if (need) {
          spin_until(flag, T_SLICES * need);
          asm volatile("fence.proxy.async.global;" ::: "memory");     // those generic-proxy writes before my TMA reads
        }
        if (i > 0) mbar_wait(a_empty, (uint32_t)(i - 1) & 1u);         // the previous operand's MMAs have read the tile
        const CUtensorMap* tm = (i == 0 || i == 3) ? &tma_st : (i == 4) ? &tma_y : &tma_a;
        for (int kb = 0; kb < T_KB; ++kb) {
          mbar_expect_tx(&a_full[kb], T_A_BYTES);
          tma_load_2d(sA + kb * T_A_BYTES, tm, &a_full[kb], kb * T_BK, m0);   // rows past M: zero fill
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(T_BM, T_BN);
      mbar_wait(w_full, 0);
      tc_fence_after();
      for (int i = 0; i < n_loads; ++i) {
        // operands 1, 2, 3 start a new accumulation: the previous GEMM's epilogue must have drained the accumulator
        if (i >= 1 && i <= 3) { mbar_wait(acc_empty, (uint32_t)(i - 1) & 1u); tc_fence_after(); }
        const int wkb0 = (i == 4) ? 4 * T_KB : i * T_KB;               // Wh[:, :512] for operand 3, Wh[:, 512:] for 4
        for (int kb = 0; kb < T_KB; ++kb) {
          mbar_wait(&a_full[kb], (uint32_t)i & 1u);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(sA + kb * T_A_BYTES));
          const uint64_t db = make_smem_desc_sw128(smem_u32(sW + (wkb0 + kb) * T_W_BYTES));
#pragma unroll
          for (int k = 0; k < T_BK / 16; ++k)
            mma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, i == 4 || (kb | k) != 0);
        }
        mma_commit(a_empty);
        if (i != 3) mma_commit(acc_full);
      }
    }
  } else {
    // ===================== worker warps =====================
    const int q = warp & 3;                             // TMEM lane quarter this warp may read
    const int clip = m0 + q * 32 + lane;                // my clip in the GEMM epilogues (thread = TMEM lane)
    const bool valid = clip < p.M;
    const int rclip = m0 + slice * T_ROWS + (warp - 2); // my warp's clip in the row stages
    const bool rvalid = rclip < p.M;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_par = 0;
    // acc[16] = this GEMM's accumulator row + bias
    auto gemm_out = [&](const float* bias, float (&acc)[16]) {
      mbar_wait(acc_full, acc_par);
      acc_par ^= 1u;
      tc_fence_after();
      uint32_t r[16];
      tmem_ld16(t_row, r);
      tmem_ld_wait_dep16(r);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0) + j);
        acc[4 * j] = __uint_as_float(r[4 * j]) + b4.x; acc[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
        acc[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z; acc[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
      }
    };
    auto store16 = [&](float* dst, const float (&v)[16]) {
      float4* d = reinterpret_cast<float4*>(dst + (int64_t)clip * kD + n0);
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    };
    // my CTA's part of a stage is stored -> one thread publishes it device-wide
    auto publish = [&]() {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        __threadfence();
        asm volatile("fence.proxy.async.global;" ::: "memory");
        atomicAdd(flag, 1);
      }
    };
    auto await = [&](int arrivals) {
      if (lane == 0) spin_until(flag, T_SLICES * arrivals);
      __syncwarp();
    };
    float acc[16];

    gemm_out(p.bu, acc);                                // u = W21 St + bu
    if (valid) store16(p.u, acc);
    publish();                                          // 1
    await(1);
    if (rvalid) {                                       // a = sum_k softmax(scale u.Lt_k) Lt_k
      const float4* base = reinterpret_cast<const float4*>(p.Lt + (int64_t)rclip * p.L * kD);
      attention_body<true>(p.u, rclip, p.L, p.scale, p.a16, 1, [&](int k) { return base + (int64_t)k * (kD / 4); });
    }
    publish();                                          // 2
    gemm_out(p.b3, acc);                                // v = W3 a + b3
    if (valid) store16(p.u, acc);
    publish();                                          // 3
    await(3);
    if (rvalid) layernorm_relu_row<true>(p.u, p.lnw, p.lnb, rclip, p.a16, 1);    // r = relu(LN(v))
    publish();                                          // 4
    gemm_out(p.b4, acc);                                // y = St + (W4 r + b4)
    if (valid) {
      const float4* sp = reinterpret_cast<const float4*>(p.St + (int64_t)clip * kD + n0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 e = __ldg(sp + j);
        acc[4 * j] += e.x; acc[4 * j + 1] += e.y; acc[4 * j + 2] += e.z; acc[4 * j + 3] += e.w;
      }
      if (p.cls) {
        uint4* d = reinterpret_cast<uint4*>(p.y16 + (int64_t)clip * kD + n0);
        d[0] = make_uint4(pack_h2(acc[0], acc[1]), pack_h2(acc[2], acc[3]), pack_h2(acc[4], acc[5]), pack_h2(acc[6], acc[7]));
        d[1] = make_uint4(pack_h2(acc[8], acc[9]), pack_h2(acc[10], acc[11]), pack_h2(acc[12], acc[13]), pack_h2(acc[14], acc[15]));
      } else {
        store16(p.y1_out, acc);
      }
    }
    if (p.cls) {
      publish();                                        // 5
      gemm_out(p.bh, acc);                              // z = relu(Wh [St || y] + bh)
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fmaxf(acc[j], 0.f);
        store16(p.z, acc);
      }
      publish();                                        // 6
      await(6);
      if (rvalid) fc_argmax_row<true>(p.z, p.wc, p.bc, rclip, p.C, p.logits, p.pred, p.score);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 32); }
}

}  // namespace umma

// Largest clip count the device can run through the fused tail (0: not at all).
int umma_head_tail_max_clips() {
  static int max_clips = -1;
  if (max_clips >= 0) return max_clips;
  using namespace umma;
  max_clips = 0;
  if (cudaFuncSetAttribute(umma_head_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM) != cudaSuccess) {
    cudaGetLastError();
    return max_clips;
  }
  int sms = 0, dev = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, umma_head_tail_kernel, T_THREADS, T_SMEM) != cudaSuccess) {
    cudaGetLastError();
    per_sm = 0;
  }
  int tiles = (sms * per_sm) / T_SLICES;                // every CTA of the grid must be resident
  if (tiles > T_MAX_TILES) tiles = T_MAX_TILES;
  max_clips = tiles * T_BM;
  return max_clips;
}

// St16 = fp16(St) (the LSTM kernel's copy, or a conversion pass).  u, z: [B][512] fp32 scratch; a16, y16: [B][512] fp16
// scratch; flags: >= ceil(B / 128) int32.  cls_packed == nullptr: relation block only, y1 (fp32, residual added) -> y1_out.
int umma_head_tail(const float* nl_packed, const float* cls_packed, const float* St, const half_t* St16, const float* Lt,
                   int B, int L, int C, float* u, half_t* a16, half_t* y16, float* z, float* y1_out, float* logits,
                   int64_t* pred, float* score, int32_t* flags, cudaStream_t st) {
  using namespace umma;
  if (B == 0) return TMR_OK;
  if (B > umma_head_tail_max_clips())
    return set_error(TMR_ERR_UNSUPPORTED, "fused relation + classifier kernel: batch exceeds the co-resident grid");
  TMR_CUDA(cudaFuncSetAttribute(umma_head_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM));   // per device
  HeadTailParams p{};
  p.M = B; p.L = L; p.C = C; p.cls = cls_packed ? 1 : 0;
  p.St = St; p.Lt = Lt; p.u = u; p.a16 = a16; p.y16 = y16; p.z = z; p.y1_out = y1_out;
  p.bu = nl_packed + NLBlockPacked::bu_off; p.b3 = nl_packed + NLBlockPacked::b3_off; p.b4 = nl_packed + NLBlockPacked::b4_off;
  p.lnw = nl_packed + NLBlockPacked::lnw_off; p.lnb = nl_packed + NLBlockPacked::lnb_off;
  if (cls_packed) {
    p.bh = cls_packed + ClassifierPacked::bh_off; p.wc = cls_packed + ClassifierPacked::wc_off; p.bc = cls_packed + ClassifierPacked::bc_off;
  } else {
    p.bh = p.b4; p.wc = nullptr; p.bc = nullptr;
  }
  p.logits = logits; p.pred = pred; p.score = score; p.flags = flags;
  p.scale = (float)0.044194173824159216;   // (1/512)**0.5 as python computes it (NLB:31)
  const int tiles = (B + T_BM - 1) / T_BM;
  CUtensorMap tst, ta, ty, tw21, tw3, tw4, twh;
  {
    uint64_t da[2] = {(uint64_t)kD, (uint64_t)B};
    uint64_t sa[1] = {(uint64_t)kD * 2};
    uint32_t ba[2] = {T_BK, T_BM};
    TMR_TRY(make_tmap(&tst, St16, 2, da, sa, ba, 2));
    TMR_TRY(make_tmap(&ta, a16, 2, da, sa, ba, 2));
    TMR_TRY(make_tmap(&ty, y16, 2, da, sa, ba, 2));
    const half_t* n16 = mirror16<NLBlockPacked>(nl_packed);
    uint64_t dw[2] = {(uint64_t)kD, (uint64_t)kD};
    uint32_t bw[2] = {T_BK, T_BN};
    TMR_TRY(make_tmap(&tw21, n16 + NLBlockPacked::w21_off, 2, dw, sa, bw, 2));
    TMR_TRY(make_tmap(&tw3, n16 + NLBlockPacked::w3_off, 2, dw, sa, bw, 2));
    TMR_TRY(make_tmap(&tw4, n16 + NLBlockPacked::w4_off, 2, dw, sa, bw, 2));
    twh = tw4;
    if (cls_packed) {
      uint64_t dh[2] = {(uint64_t)2 * kD, (uint64_t)kD};
      uint64_t sh[1] = {(uint64_t)2 * kD * 2};
      TMR_TRY(make_tmap(&twh, mirror16<ClassifierPacked>(cls_packed) + ClassifierPacked::wh_off, 2, dh, sh, bw, 2));
    }
  }
  TMR_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t) * tiles, st));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(T_SLICES * tiles)); cfg.blockDim = dim3(T_THREADS); cfg.dynamicSmemBytes = T_SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  static int coop_ok = 1;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
  if (coop_ok && cap == cudaStreamCaptureStatusNone) {   // (a refused launch would invalidate a stream capture)
    cudaError_t e = cudaLaunchKernelEx(&cfg, umma_head_tail_kernel, tst, ta, ty, tw21, tw3, tw4, twh, p);
    if (e == cudaSuccess) return TMR_OK;
    cudaGetLastError();
    coop_ok = 0;
  }
  cfg.numAttrs = 0;
  TMR_CUDA(cudaLaunchKernelEx(&cfg, umma_head_tail_kernel, tst, ta, ty, tw21, tw3, tw4, twh, p));
  return TMR_OK;
}

}  // namespace tmr
