// The LSTM recurrence (TRAIN:224, 241-244) for the reference's OWN batch sizes (EVAL:470-495 / TRAIN:836-880 call the
// head on 120 clips at a time): steps 1 .. seq-1 of up to 512 clips in one launch whose step time is a LATENCY chain,
// not a throughput problem.  The large-batch persistent kernel (umma_lstm_persist.cu) gives a CTA pair a 256-clip x
// 256-column item - at 120 clips that is 16 SMs running 2.3 us of MMAs, a 4 us epilogue and a 2.5 us h exchange back
// to back, 14 us per step.  Here the SAME step is cut the other way:
//   * a tile is 128 clips (one TMEM lane each); its 2048 gate columns are spread over 32 CTAs of 64 columns
//     (= the i,f,g,o gates of 16 hidden units), so a step's MMAs are 32 instructions (M = 128, N = 64, K = 16) and its
//     epilogue is 16 cells per thread;
//   * the CTA keeps its 64 rows of Whh' (64 KB) resident and has room for the WHOLE h tile (128 clips x 512 fp16 =
//     128 KB, eight TMA boxes on eight mbarriers: the MMAs of k-block kb start when ITS box has landed);
//   * c stays in registers for all steps (a thread owns one clip x 16 units), the thread's 64 projected values
//     of the step are requested from L2 before it waits for the accumulator;
//   * h_t goes out as ONE 32-byte store per thread; the four epilogue warps meet at a named barrier and one thread
//     pays the device-scope fence and bumps the tile's arrival counter; the TMA producers of the tile's 32 CTAs spin
//     on that counter (no sleep: there are at most 128 pollers on the device).
// Same operands, same K order and the same cell arithmetic as the large-batch kernels: h_T is checked bit for bit
// against them in the tests.  Grid = 32 x ceil(B / 128) CTAs, all of which must be resident (cooperative launch).
#include "tmr_internal.h"
#include "umma_common.cuh"

namespace tmr {
namespace umma {

constexpr int S_BM = 128;                     // clips per tile
constexpr int S_BN = 64;                      // gate columns per CTA
constexpr int S_BK = 64;                      // fp16 per 128-byte swizzle row
constexpr int S_KB = kD / S_BK;               // 8 k-blocks
constexpr int S_A_BYTES = S_BM * S_BK * 2;    // 16 KB
constexpr int S_B_BYTES = S_BN * S_BK * 2;    //  8 KB
constexpr int S_SLICES = 4 * kD / S_BN;       // 32 CTAs per tile
constexpr int S_UNITS = S_BN / 4;             // 16 hidden units per CTA
constexpr int S_SMEM = S_KB * (S_A_BYTES + S_B_BYTES) + 1024 + 256;
constexpr int S_THREADS = 64 + 128;           // TMA producer, MMA issuer, four epilogue warps
constexpr int S_MAX_TILES = 4;

struct LstmSmallParams {
  int64_t M;                       // clips
  const float* xp; const int64_t* starts; int seq;
  const float* c0;                 // c after step 0 [M][512]
  half_t* h16a; half_t* h16b;      // h exchange buffers [M][512] fp16; h16a holds h after step 0 on entry
  float* h_out;                    // h after the last step (fp32, the clip's St)
  int32_t* flags;                  // [tiles] arrivals; zero on entry
};

__device__ __forceinline__ int ld_acquire_gpu_s(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(S_THREADS, 1)
umma_lstm_small_kernel(const __grid_constant__ CUtensorMap tma_h0, const __grid_constant__ CUtensorMap tma_h1,
                       const __grid_constant__ CUtensorMap tma_b, const LstmSmallParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                                   // [S_KB][64 weight rows][64 fp16], resident
  uint8_t* sA = sB + S_KB * S_B_BYTES;                  // [S_KB][128 clips][64 fp16]: h_{t-1} of the whole tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + S_KB * S_A_BYTES);
  uint64_t* a_full = bars;                    // [S_KB]  TMA -> MMA, one completion per step
  uint64_t* a_empty = a_full + S_KB;          // [1]     MMA -> TMA: the step's MMAs have read the h tile
  uint64_t* b_full = a_empty + 1;             // [1]
  uint64_t* acc_full = b_full + 1;            // [1]
  uint64_t* acc_empty = acc_full + 1;         // [1]     four epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x / S_SLICES;
  const int slice = blockIdx.x % S_SLICES;
  const int n0 = slice * S_BN;                          // my gate columns
  const int m0 = tile * S_BM;                           // my tile's first clip

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_h0); tma_prefetch_desc(&tma_h1); tma_prefetch_desc(&tma_b);
    for (int k = 0; k < S_KB; ++k) mbar_init(&a_full[k], 1);
    mbar_init(a_empty, 1); mbar_init(b_full, 1); mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, S_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(b_full, S_KB * S_B_BYTES);
      for (int kb = 0; kb < S_KB; ++kb) tma_load_2d(sB + kb * S_B_BYTES, &tma_b, b_full, kb * S_BK, n0);
      for (int t = 1; t < p.seq; ++t) {
        if (t > 1) {
          // h_{t-1} of the tile is complete when all 32 CTAs have arrived t-1 times
          const int32_t* f = p.flags + tile;
          const int target = S_SLICES * (t - 1);
          const long long t0 = clock64();
          while (ld_acquire_gpu_s(f) < target) {
            if (clock64() - t0 > 4000000000LL) __trap();
          }
          asm volatile("fence.proxy.async.global;" ::: "memory");   // those generic-proxy writes before my TMA reads
          mbar_wait(a_empty, (uint32_t)(t - 2) & 1u);                // step t-1's MMAs have read the tile
        }
        const CUtensorMap* th = ((t - 1) & 1) ? &tma_h1 : &tma_h0;
        for (int kb = 0; kb < S_KB; ++kb) {
          mbar_expect_tx(&a_full[kb], S_A_BYTES);
          tma_load_2d(sA + kb * S_A_BYTES, th, &a_full[kb], kb * S_BK, m0);     // rows past M: zero fill
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(S_BM, S_BN);
      mbar_wait(b_full, 0);
      tc_fence_after();
      for (int t = 1; t < p.seq; ++t) {
        if (t > 1) { mbar_wait(acc_empty, (uint32_t)(t - 2) & 1u); tc_fence_after(); }
        for (int kb = 0; kb < S_KB; ++kb) {
          mbar_wait(&a_full[kb], (uint32_t)(t - 1) & 1u);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(sA + kb * S_A_BYTES));
          const uint64_t db = make_smem_desc_sw128(smem_u32(sB + kb * S_B_BYTES));
#pragma unroll
          for (int k = 0; k < S_BK / 16; ++k)
            mma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        mma_commit(a_empty);
        mma_commit(acc_full);
      }
    }
  } else {
    // ===================== epilogue warps: thread = clip, 16 hidden units =====================
    const int q = warp & 3;                             // TMEM lane quarter this warp may read
    const int64_t clip = (int64_t)m0 + q * 32 + lane;
    const bool valid = clip < p.M;
    const int u0 = slice * S_UNITS;                     // my first hidden unit
    const int64_t xbase = valid ? (p.starts ? p.starts[clip] : clip * p.seq) : 0;
    float cst[S_UNITS];
#pragma unroll
    for (int k = 0; k < S_UNITS; ++k) cst[k] = 0.f;
    if (valid) {
      float v[8];
      ldg256(p.c0 + clip * kD + u0, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) cst[k] = v[k];
      ldg256(p.c0 + clip * kD + u0 + 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) cst[8 + k] = v[k];
    }
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int t = 1; t < p.seq; ++t) {
      const bool last_step = (t == p.seq - 1);
      // my 64 projected values of this step, requested before the accumulator is awaited
      float4 xv[16];
      if (valid) {
        const float4* xr = reinterpret_cast<const float4*>(p.xp + (xbase + t) * (4 * kD) + n0);
#pragma unroll
        for (int j = 0; j < 16; ++j) xv[j] = __ldg(xr + j);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(acc_full, (uint32_t)(t - 1) & 1u);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld32(t_row, r0);
      tmem_ld32(t_row + 32, r1);
      tmem_ld_wait();
      tc_fence_before();                                // the accumulator may be overwritten by the next step
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      float hn[S_UNITS];
#pragma unroll
      for (int j = 0; j < S_UNITS; ++j) {
        const uint32_t* r = j < 8 ? r0 : r1;
        const int b = 4 * (j & 7);
        float cn;
        lstm_cell_fast(__uint_as_float(r[b]) + xv[j].x, __uint_as_float(r[b + 1]) + xv[j].y, __uint_as_float(r[b + 2]) + xv[j].z,
                       __uint_as_float(r[b + 3]) + xv[j].w, cst[j], cn, hn[j]);
        cst[j] = cn;
      }
      if (valid) {                                      // fp16 h: the next step's MMA operand (the last step's: the
        half_t* hdst = (t & 1) ? p.h16b : p.h16a;       // relation block's query operand), one 32-byte sector
        const uint2 a = pack_h4(hn[0], hn[1], hn[2], hn[3]), b = pack_h4(hn[4], hn[5], hn[6], hn[7]);
        const uint2 c = pack_h4(hn[8], hn[9], hn[10], hn[11]), d = pack_h4(hn[12], hn[13], hn[14], hn[15]);
        const uint32_t w8[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
        stg256u(hdst + clip * kD + u0, w8);
      }
      if (last_step) {
        if (valid) {                                    // the last step's h is the clip's St: fp32
          const float lo[8] = {hn[0], hn[1], hn[2], hn[3], hn[4], hn[5], hn[6], hn[7]};
          const float hi[8] = {hn[8], hn[9], hn[10], hn[11], hn[12], hn[13], hn[14], hn[15]};
          stg256(p.h_out + clip * kD + u0, lo);
          stg256(p.h_out + clip * kD + u0 + 8, hi);
        }
      } else {
        // my CTA's 128 clips x 16 units of h_t are stored -> one thread publishes them device-wide
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane == 0) {
          __threadfence();
          asm volatile("fence.proxy.async.global;" ::: "memory");
          atomicAdd(p.flags + tile, 1);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, S_BN); }
}


// ---- training forward (csrc/train.cu, TMR_MATH_F16): the same one-launch recurrence on the RAW parameter layout --------
// Whh (4D, D) in torch's gate-block row order (i | f | g | o blocks of D rows), projected rows xp time-major
// ([t * B + b][4D], same column order), ALL steps including t = 0 (zero state, no MMA), exact expf / tanhf gates, and
// every step's activated gates, c_t and h_t saved for BPTT (gates [T][4D], c / h [T][D], T = seq * B rows, time-major).
// CTA j owns hidden units 16j .. 16j+15: its B tile is four 16-row TMA boxes (one per gate block), so TMEM column
// g * 16 + u is gate g of unit 16j + u.  h_{t-1} reaches the MMA as fp16 exactly as the per-step path's conversion pass
// rounds it, products accumulate in the same K order: bit-identical to that path.
struct LstmTrainParams {
  int B; int seq;
  const float* xp;                 // [seq * B][4D]
  float* gates; float* c; float* h;
  half_t* h16a; half_t* h16b;      // exchange buffers [B][512] fp16
  int32_t* flags;                  // [tiles], zero on entry
};

__device__ __forceinline__ float sigm_exact(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void __launch_bounds__(S_THREADS, 1)
umma_lstm_train_fwd_kernel(const __grid_constant__ CUtensorMap tma_h0, const __grid_constant__ CUtensorMap tma_h1,
                           const __grid_constant__ CUtensorMap tma_b, const LstmTrainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                                   // [S_KB][4 gates x 16 weight rows][64 fp16], resident
  uint8_t* sA = sB + S_KB * S_B_BYTES;                  // [S_KB][128 clips][64 fp16]: h_{t-1} of the whole tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + S_KB * S_A_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + S_KB;
  uint64_t* b_full = a_empty + 1;
  uint64_t* acc_full = b_full + 1;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x / S_SLICES;
  const int slice = blockIdx.x % S_SLICES;
  const int u0 = slice * S_UNITS;                       // my first hidden unit
  const int m0 = tile * S_BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_h0); tma_prefetch_desc(&tma_h1); tma_prefetch_desc(&tma_b);
    for (int k = 0; k < S_KB; ++k) mbar_init(&a_full[k], 1);
    mbar_init(a_empty, 1); mbar_init(b_full, 1); mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, S_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(b_full, S_KB * S_B_BYTES);
      for (int kb = 0; kb < S_KB; ++kb)
        for (int g = 0; g < 4; ++g)                     // 16 rows of gate block g: rows g * D + u0 ..
          tma_load_2d(sB + kb * S_B_BYTES + g * (S_UNITS * S_BK * 2), &tma_b, b_full, kb * S_BK, g * kD + u0);
      for (int t = 1; t < p.seq; ++t) {
        // h_{t-1} of the tile is complete when all 32 CTAs have arrived t times (steps 0 .. t-1)
        const int32_t* f = p.flags + tile;
        const int target = S_SLICES * t;
        const long long t0 = clock64();
        while (ld_acquire_gpu_s(f) < target) {
          if (clock64() - t0 > 4000000000LL) __trap();
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");
        if (t > 1) mbar_wait(a_empty, (uint32_t)(t - 2) & 1u);
        const CUtensorMap* th = ((t - 1) & 1) ? &tma_h1 : &tma_h0;
        for (int kb = 0; kb < S_KB; ++kb) {
          mbar_expect_tx(&a_full[kb], S_A_BYTES);
          tma_load_2d(sA + kb * S_A_BYTES, th, &a_full[kb], kb * S_BK, m0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(S_BM, S_BN);
      mbar_wait(b_full, 0);
      tc_fence_after();
      for (int t = 1; t < p.seq; ++t) {
        if (t > 1) { mbar_wait(acc_empty, (uint32_t)(t - 2) & 1u); tc_fence_after(); }
        for (int kb = 0; kb < S_KB; ++kb) {
          mbar_wait(&a_full[kb], (uint32_t)(t - 1) & 1u);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(sA + kb * S_A_BYTES));
          const uint64_t db = make_smem_desc_sw128(smem_u32(sB + kb * S_B_BYTES));
#pragma unroll
          for (int k = 0; k < S_BK / 16; ++k)
            mma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        mma_commit(a_empty);
        mma_commit(acc_full);
      }
    }
  } else {
    const int q = warp & 3;
    const int clip = m0 + q * 32 + lane;
    const bool valid = clip < p.B;
    float cst[S_UNITS];
#pragma unroll
    for (int k = 0; k < S_UNITS; ++k) cst[k] = 0.f;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int t = 0; t < p.seq; ++t) {
      const int64_t row = (int64_t)t * p.B + clip;      // time-major row of (t, clip)
      float4 xv[16];                                    // [gate][4 float4]: 16 projected values per gate
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4* xr = reinterpret_cast<const float4*>(p.xp + row * (4 * kD) + g * kD + u0);
#pragma unroll
          for (int j = 0; j < 4; ++j) xv[4 * g + j] = __ldg(xr + j);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      uint32_t r0[32], r1[32];                          // columns 0..31 = gates i, f; 32..63 = gates g, o
      if (t > 0) {
        mbar_wait(acc_full, (uint32_t)(t - 1) & 1u);
        tc_fence_after();
        tmem_ld32(t_row, r0);
        tmem_ld32(t_row + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) { r0[j] = 0u; r1[j] = 0u; }
      }
      float gi[S_UNITS], gf[S_UNITS], gg[S_UNITS], go[S_UNITS], hn[S_UNITS];
#pragma unroll
      for (int u = 0; u < S_UNITS; ++u) {
        const float* xi = reinterpret_cast<const float*>(&xv[0]);
        // pre = xp + h.Whh^T (the per-step path adds in this order); at t = 0 the recurrent term is absent
        const float pi = (t > 0) ? xi[u] + __uint_as_float(r0[u]) : xi[u];
        const float pf = (t > 0) ? xi[16 + u] + __uint_as_float(r0[16 + u]) : xi[16 + u];
        const float pg = (t > 0) ? xi[32 + u] + __uint_as_float(r1[u]) : xi[32 + u];
        const float po = (t > 0) ? xi[48 + u] + __uint_as_float(r1[16 + u]) : xi[48 + u];
        gi[u] = sigm_exact(pi); gf[u] = sigm_exact(pf); gg[u] = tanhf(pg); go[u] = sigm_exact(po);
        const float cn = gf[u] * cst[u] + gi[u] * gg[u];
        cst[u] = cn;
        hn[u] = go[u] * tanhf(cn);
      }
      if (valid) {
        auto st16 = [&](float* dst, const float (&v)[S_UNITS]) {
          float4* d = reinterpret_cast<float4*>(dst);
#pragma unroll
          for (int j = 0; j < 4; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        };
        float* gr = p.gates + row * (4 * kD) + u0;
        st16(gr, gi); st16(gr + kD, gf); st16(gr + 2 * kD, gg); st16(gr + 3 * kD, go);
        st16(p.c + row * kD + u0, cst);
        st16(p.h + row * kD + u0, hn);
      }
      if (t + 1 < p.seq) {
        if (valid) {
          half_t* hdst = (t & 1) ? p.h16b : p.h16a;
          const uint2 a = pack_h4(hn[0], hn[1], hn[2], hn[3]), b = pack_h4(hn[4], hn[5], hn[6], hn[7]);
          const uint2 c = pack_h4(hn[8], hn[9], hn[10], hn[11]), d = pack_h4(hn[12], hn[13], hn[14], hn[15]);
          const uint32_t w8[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
          stg256u(hdst + (int64_t)clip * kD + u0, w8);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane == 0) {
          __threadfence();
          asm volatile("fence.proxy.async.global;" ::: "memory");
          atomicAdd(p.flags + tile, 1);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, S_BN); }
}

}  // namespace umma

// Largest clip count the device can run through the small-batch recurrence (0: not at all).
int umma_lstm_small_max_clips() {
  static int max_clips = -1;
  if (max_clips >= 0) return max_clips;
  using namespace umma;
  max_clips = 0;
  if (cudaFuncSetAttribute(umma_lstm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM) != cudaSuccess) {
    cudaGetLastError();
    return max_clips;
  }
  int sms = 0, dev = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, umma_lstm_small_kernel, S_THREADS, S_SMEM) != cudaSuccess) {
    cudaGetLastError();
    per_sm = 0;
  }
  int tiles = (sms * per_sm) / S_SLICES;                // every CTA of the grid must be resident
  if (tiles > S_MAX_TILES) tiles = S_MAX_TILES;
  max_clips = tiles * S_BM;
  return max_clips;
}

// Recurrent steps 1 .. seq-1 of B <= umma_lstm_small_max_clips() clips in one launch.  c0 / h16a hold the state after
// step 0; h16b is the second exchange buffer; flags: >= ceil(B / 128) int32 of scratch.  On return the buffer of
// parity (seq - 1) & 1 (h16b for even seq) also holds fp16(h_T).
int umma_lstm_small(const half_t* whh16, const float* xp, const int64_t* starts, int seq, half_t* h16a, half_t* h16b,
                    float* h_out, const float* c0, int B, int32_t* flags, cudaStream_t st) {
  using namespace umma;
  if (B == 0 || seq < 2) return TMR_OK;
  if (B > umma_lstm_small_max_clips())
    return set_error(TMR_ERR_UNSUPPORTED, "small-batch LSTM recurrence: batch exceeds the co-resident grid");
  TMR_CUDA(cudaFuncSetAttribute(umma_lstm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM));   // per device
  LstmSmallParams p{};
  p.M = B; p.xp = xp; p.starts = starts; p.seq = seq; p.c0 = c0; p.h16a = h16a; p.h16b = h16b; p.h_out = h_out; p.flags = flags;
  const int tiles = (B + S_BM - 1) / S_BM;
  CUtensorMap th0, th1, tb;
  {
    uint64_t da[2] = {(uint64_t)kD, (uint64_t)B};
    uint64_t sa[1] = {(uint64_t)kD * 2};
    uint32_t ba[2] = {S_BK, S_BM};
    TMR_TRY(make_tmap(&th0, h16a, 2, da, sa, ba, 2));
    TMR_TRY(make_tmap(&th1, h16b, 2, da, sa, ba, 2));
    uint64_t dw[2] = {(uint64_t)kD, (uint64_t)4 * kD};
    uint32_t bw[2] = {S_BK, S_BN};
    TMR_TRY(make_tmap(&tb, whh16, 2, dw, sa, bw, 2));
  }
  TMR_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t) * tiles, st));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(S_SLICES * tiles)); cfg.blockDim = dim3(S_THREADS); cfg.dynamicSmemBytes = S_SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  static int coop_ok = 1;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
  // (a refused launch would invalidate a stream capture: inside one, rely on the occupancy bound)
  if (coop_ok && cap == cudaStreamCaptureStatusNone) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, umma_lstm_small_kernel, th0, th1, tb, p);
    if (e == cudaSuccess) return TMR_OK;
    cudaGetLastError();
    coop_ok = 0;
  }
  cfg.numAttrs = 0;
  TMR_CUDA(cudaLaunchKernelEx(&cfg, umma_lstm_small_kernel, th0, th1, tb, p));
  return TMR_OK;
}

// Training forward of the recurrence (all seq steps, activations saved) for B <= umma_lstm_small_max_clips() clips.
// whh16: fp16 copy of the RAW Whh (4D, D); h16a / h16b: [B][512] fp16 scratch; flags: >= ceil(B / 128) int32.
int umma_lstm_train_fwd(const half_t* whh16, const float* xp, int B, int seq, float* gates, float* c, float* h,
                        half_t* h16a, half_t* h16b, int32_t* flags, cudaStream_t st) {
  using namespace umma;
  if (B == 0 || seq < 1) return TMR_OK;
  if (B > umma_lstm_small_max_clips())
    return set_error(TMR_ERR_UNSUPPORTED, "one-launch training recurrence: batch exceeds the co-resident grid");
  // (per call, not once per process: the attribute is per device, and one process may drive several)
  TMR_CUDA(cudaFuncSetAttribute(umma_lstm_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM));
  LstmTrainParams p{};
  p.B = B; p.seq = seq; p.xp = xp; p.gates = gates; p.c = c; p.h = h; p.h16a = h16a; p.h16b = h16b; p.flags = flags;
  const int tiles = (B + S_BM - 1) / S_BM;
  CUtensorMap th0, th1, tb;
  {
    uint64_t da[2] = {(uint64_t)kD, (uint64_t)B};
    uint64_t sa[1] = {(uint64_t)kD * 2};
    uint32_t ba[2] = {S_BK, S_BM};
    TMR_TRY(make_tmap(&th0, h16a, 2, da, sa, ba, 2));
    TMR_TRY(make_tmap(&th1, h16b, 2, da, sa, ba, 2));
    uint64_t dw[2] = {(uint64_t)kD, (uint64_t)4 * kD};
    uint32_t bw[2] = {S_BK, S_UNITS};                   // 16 rows of one gate block
    TMR_TRY(make_tmap(&tb, whh16, 2, dw, sa, bw, 2));
  }
  TMR_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t) * tiles, st));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(S_SLICES * tiles)); cfg.blockDim = dim3(S_THREADS); cfg.dynamicSmemBytes = S_SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  static int coop_ok = 1;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
  if (coop_ok && cap == cudaStreamCaptureStatusNone) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, umma_lstm_train_fwd_kernel, th0, th1, tb, p);
    if (e == cudaSuccess) return TMR_OK;
    cudaGetLastError();
    coop_ok = 0;
  }
  cfg.numAttrs = 0;
  TMR_CUDA(cudaLaunchKernelEx(&cfg, umma_lstm_train_fwd_kernel, th0, th1, tb, p));
  return TMR_OK;
}

}  // namespace tmr
