// The LSTM recurrence (TRAIN:224, 241-244) as ONE persistent launch over all seq-1 recurrent steps:
//     gates_t = h_{t-1} . Whh'^T + xp[frame + t],   (c_t, h_t) = cell(gates_t, c_{t-1}),   t = 1 .. seq-1
// (step 0 from zero state rides in the input projection's epilogue).  The per-step kernel (umma_lstm_ws.cu) moved
// 14 KB per clip and step through HBM: c in and out, h in and out, the projected row - nine times over.  Here
//   * c never leaves the SM: an epilogue thread owns one clip (TMEM lane) and 32 hidden units for the whole
//     recurrence and keeps their cell state in registers;
//   * h_t is exchanged between the SMs through L2: the eight CTA pairs that own the eight 256-column gate slices
//     of a 256-clip tile each write their 64 hidden units of h_t (fp16) and bump a per-tile counter (a dedicated
//     publisher warp pays the device-scope fence); the TMA producers of the same eight pairs wait for that
//     counter before loading the tile's h rows for step t+1;
//   * projected rows of consecutive steps overlap (clip b at step t reads the row clip b+1 read at step t-1), and
//     the tiles in flight (two per pair) cover a few thousand clips, so those reads hit L2.
// As in the per-step kernel a CTA pair (cta_group::2, M = 256) keeps its slice of Whh' resident in shared memory
// (128 KB per CTA).  Two clip tiles are in flight per pair, one per TMEM accumulator: while the epilogue warps run
// the cells of tile A's step t, the tensor cores run tile B's step t, and A's step t+1 only waits for the other
// seven slices of A.  Grid = 8 slices x G groups of pairs (G = 9 on 148 SMs); group g owns tiles g, g+G, ...
//
// The pairs of a group wait on one another, so every CTA of the grid must be resident: the grid never exceeds the
// device's co-resident cluster capacity (checked with cudaOccupancyMaxActiveClusters; cooperative launch).  Every
// wait is bounded (__trap after ~2 s) so a scheduling surprise is a launch error, not a hung GPU.
#include "tmr_internal.h"
#include "umma_common.cuh"

namespace tmr {
namespace umma {

constexpr int P_BM = 128;                     // clips per CTA (256 per pair)
constexpr int P_BN = 256;                     // gate columns per pair; each CTA holds 128 of the weight rows
constexpr int P_BK = 64;                      // fp16 per 128-byte swizzle row
constexpr int P_KB = kD / P_BK;               // 8 k-blocks
constexpr int P_A_BYTES = P_BM * P_BK * 2;    // 16 KB
constexpr int P_B_BYTES = (P_BN / 2) * P_BK * 2;   // 16 KB per k-block per CTA
constexpr int P_NA = 4;                       // h-row stages
constexpr int P_EW = 8;                       // epilogue warps: 4 TMEM lane quarters x 2 column halves
constexpr int P_WCOLS = P_BN * 4 / P_EW;      // 128 gate columns = 32 hidden units per epilogue warp
constexpr int P_NCH = P_WCOLS / 32;           // 4 chunks of 32 columns (8 units)
constexpr int P_SLICES = 4 * kD / P_BN;       // 8
constexpr int P_SMEM = P_KB * P_B_BYTES + P_NA * P_A_BYTES + P_EW * 4096 + 1024 + 512;
constexpr int P_THREADS = 64 + 32 * P_EW + 32;         // + the publisher warp
constexpr int P_ARRIVALS = P_SLICES;          // per (tile half, step): one arrival per slice's CTA

struct LstmPersistParams {
  int64_t M;                       // clips
  const float* xp; const int64_t* starts; int seq;
  const float* c0;                 // c after step 0 [M][512]
  half_t* h16a; half_t* h16b;      // h exchange buffers [M][512] fp16; h16a holds h after step 0 on entry
                                   // (two fields, not an array: a dynamically indexed parameter array would move
                                   // the whole parameter block to local memory)
  float* h_out;                    // h after the last step (fp32, the clip's St)
  int x_tma; int64_t x_row0;       // tma_x covers the projected rows; its row 0 is projected row x_row0
  int groups;                      // G: groups of 8 pairs
  int64_t m_pairs;                 // 256-clip tiles
  int32_t* flags;                  // [2 * m_pairs] arrivals per (tile, CTA half); zero on entry
};

__device__ __forceinline__ int ld_acquire_gpu(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Experiment builds (-DTMR_EXPERIMENT): %globaltimer stamps per CTA and item - TMA producer [0] flag wait begins,
// [1] flag seen; MMA thread [2] accumulator free, [3] last commit issued; epilogue warp 2 [4] waits for the
// accumulator, [5] has it, [6] published h.  Read back with tmr_debug_persist_timeline (scripts/persist_timeline.py).
#ifdef TMR_EXPERIMENT
constexpr int PTL_ITEMS = 96;
__device__ unsigned long long g_ptl[148 * PTL_ITEMS * 8];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PTL(item, k) do { if ((item) < PTL_ITEMS && blockIdx.x < 148) g_ptl[((size_t)blockIdx.x * PTL_ITEMS + (item)) * 8 + (k)] = gtimer(); } while (0)
#else
#define PTL(item, k) do { } while (0)
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
umma_lstm_persist_kernel(const __grid_constant__ CUtensorMap tma_h0, const __grid_constant__ CUtensorMap tma_h1,
                         const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_x,
                         const LstmPersistParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                                   // [P_KB][128 weight rows][64 fp16], resident
  uint8_t* sA = sB + P_KB * P_B_BYTES;                  // [P_NA][128 clips][64 fp16]
  float* sX = reinterpret_cast<float*>(sA + P_NA * P_A_BYTES);     // [epilogue warps][32 clips x 32 fp32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sX) + P_EW * 4096);
  uint64_t* a_full = bars;                    // [P_NA]  TMA -> MMA (leader's barrier collects both CTAs' bytes)
  uint64_t* a_empty = bars + P_NA;            // [P_NA]  MMA -> TMA (multicast commit)
  uint64_t* b_full = bars + 2 * P_NA;         // [1]
  uint64_t* acc_full = b_full + 1;            // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint64_t* xfull = acc_empty + 2;            // [P_EW]
  uint64_t* h_done = xfull + P_EW;            // [2]  epilogue warps -> publisher: this CTA's part of h_t is written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int slice = pair % P_SLICES;                    // my 256 gate columns
  const int group = pair / P_SLICES;
  const int n0 = slice * P_BN;
  const int G = p.groups;
  // tiles of my group: group, group + G, ...; two in flight (slots 0 / 1 = TMEM accumulators 0 / 1)
  const int64_t n_local = (p.m_pairs > group) ? (p.m_pairs - group + G - 1) / G : 0;
  constexpr uint16_t kMask = 3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_h0); tma_prefetch_desc(&tma_h1); tma_prefetch_desc(&tma_b); tma_prefetch_desc(&tma_x);
    for (int s = 0; s < P_NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(b_full, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 2 * P_EW); }
    for (int i = 0; i < P_EW; ++i) mbar_init(&xfull[i], 1);
    for (int a = 0; a < 2; ++a) mbar_init(&h_done[a], P_EW);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // peer barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own weight rows once, own clip rows per item) =============
    if (lane == 0) {
      if (crank == 0) mbar_expect_tx(b_full, 2 * P_KB * P_B_BYTES);
      for (int kb = 0; kb < P_KB; ++kb)
        tma_load_2d_2sm(sB + kb * P_B_BYTES, &tma_b, b_full, kb * P_BK, n0 + (int)crank * (P_BN / 2));
      int stage = 0; uint32_t phase = 0;
      int item = 0;
      for (int64_t li = 0; li < n_local; li += 2) {
        const int ns = (li + 1 < n_local) ? 2 : 1;
        for (int t = 1; t < p.seq; ++t) {
          for (int s = 0; s < ns; ++s, ++item) {
            const int64_t mp = group + (li + s) * G;
            const int m0 = (int)(mp * 2 + crank) * P_BM;
            PTL(item, 0);
            if (t > 1) {
              // h_{t-1} of my 128 clips is complete when all 8 slices' CTAs of this half have arrived t-1 times
              const int32_t* f = p.flags + (mp * 2 + crank);
              const int target = P_ARRIVALS * (t - 1);
              if (ld_acquire_gpu(f) < target) {
                const long long t0 = clock64();
                while (ld_acquire_gpu(f) < target) {
                  __nanosleep(64);
                  if (clock64() - t0 > 4000000000LL) __trap();
                }
              }
              fence_proxy_async_global();     // those generic-proxy writes before my async-proxy (TMA) reads
            }
            PTL(item, 1);
            const CUtensorMap* th = ((t - 1) & 1) ? &tma_h1 : &tma_h0;
            for (int kb = 0; kb < P_KB; ++kb) {
              mbar_wait(&a_empty[stage], phase ^ 1);
              if (crank == 0) mbar_expect_tx(&a_full[stage], 2 * P_A_BYTES);
              tma_load_2d_2sm(sA + stage * P_A_BYTES, th, &a_full[stage], kb * P_BK, m0);
              if (++stage == P_NA) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(2 * P_BM, P_BN);
      mbar_wait(b_full, 0);
      tc_fence_after();
      int stage = 0; uint32_t phase = 0;
      uint32_t use0 = 0u, use1 = 0u;                        // uses of each accumulator (scalars: no local-memory array)
      int item = 0;
      for (int64_t li = 0; li < n_local; li += 2) {
        const int ns = (li + 1 < n_local) ? 2 : 1;
        for (int t = 1; t < p.seq; ++t) {
          for (int s = 0; s < ns; ++s, ++item) {
            const uint32_t u = s ? use1 : use0;
            mbar_wait(&acc_empty[s], (u & 1u) ^ 1u);          // both CTAs' epilogues have drained this accumulator
            if (s) ++use1; else ++use0;
            tc_fence_after();
            PTL(item, 2);
            const uint32_t d_tmem = tmem_base + s * P_BN;
            for (int kb = 0; kb < P_KB; ++kb) {
              mbar_wait(&a_full[stage], phase);
              tc_fence_after();
              const uint64_t da = make_smem_desc_sw128(smem_u32(sA + stage * P_A_BYTES));
              const uint64_t db = make_smem_desc_sw128(smem_u32(sB + kb * P_B_BYTES));
#pragma unroll
              for (int k = 0; k < P_BK / 16; ++k)
                mma_f16_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
              mma_commit_2sm_mcast(&a_empty[stage], kMask);     // frees the h-row stage in both CTAs
              if (++stage == P_NA) { stage = 0; phase ^= 1; }
            }
            mma_commit_2sm_mcast(&acc_full[s], kMask);          // accumulator complete -> both CTAs' epilogues
            PTL(item, 3);
          }
        }
      }
    }
  } else if (warp == 2 + P_EW) {
    // ===================== publisher =====================
    // The epilogue warps hand "my 32 clips x 32 units of h_t are stored" to this thread through a CTA-scope
    // mbarrier (release / acquire), and THIS thread pays the gpu-scope fence before bumping the tile half's
    // counter: the fence waits until the CTA's stores are visible device-wide (hundreds of ns), which would
    // otherwise stall every epilogue warp once per item.  Causality is cumulative across the two hops.
    if (lane == 0) {
      uint32_t use0 = 0u, use1 = 0u;
      for (int64_t li = 0; li < n_local; li += 2) {
        const int ns = (li + 1 < n_local) ? 2 : 1;
        for (int t = 1; t + 1 < p.seq; ++t) {               // the last step's h is the output: nothing to publish
          for (int s = 0; s < ns; ++s) {
            mbar_wait(&h_done[s], (s ? use1 : use0) & 1u);
            if (s) ++use1; else ++use0;
            __threadfence();
            fence_proxy_async_global();
            atomicAdd(p.flags + ((group + (li + s) * G) * 2 + crank), 1);
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                             // TMEM lane quarter this warp may read
    const int colq = (warp - 2) >> 2;                   // which 128-column half of the slice
    float* sb = sX + (warp - 2) * 1024;
    uint64_t* my_xfull = xfull + (warp - 2);
    const int ncol0 = n0 + colq * P_WCOLS;              // first gate column of this warp
    uint32_t x_par = 0;
    bool x_pend = false;                                // a TMA load of projected rows into my tile is in flight
    auto issue_x = [&](int x0, int col) {               // rows x0 .. x0+31, columns col .. col+31 -> my tile
      if (lane == 0) {
        mbar_expect_tx(my_xfull, 4096);
        tma_load_2d(sb, &tma_x, my_xfull, col, (int)(x0 - p.x_row0));
      }
    };
    auto clip_row = [&](int64_t mp) -> int64_t { return (mp * 2 + crank) * P_BM + q * 32 + lane; };
    auto x_row = [&](int64_t mrow, int t) -> int {
      return (mrow < p.M) ? (int)((p.starts ? p.starts[mrow] : mrow * p.seq) + t) : -1;
    };
    float cst[2][32];                                   // cell state of my clip's 32 units, per slot
    uint32_t use[2] = {0u, 0u};
    int item = 0;

    for (int64_t li = 0; li < n_local; li += 2) {
      const int ns = (li + 1 < n_local) ? 2 : 1;
      int64_t mrow_s[2]; int xbase_s[2];
      mrow_s[0] = clip_row(group + li * G);
      mrow_s[1] = (ns == 2) ? clip_row(group + (li + 1) * G) : p.M;
      xbase_s[0] = x_row(mrow_s[0], 0);
      xbase_s[1] = x_row(mrow_s[1], 0);
      // c after step 0 of both tiles
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (s < ns && mrow_s[s] < p.M) {
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float v[8];
            ldg256(p.c0 + mrow_s[s] * kD + (ncol0 >> 2) + 8 * g8, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) cst[s][8 * g8 + k] = v[k];
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) cst[s][k] = 0.f;
        }
      }
      // first projected-row tile of the chunk's first item
      {
        const int xr0 = xbase_s[0] < 0 ? -1 : xbase_s[0] + 1;
        const int x0 = __shfl_sync(0xffffffffu, xr0, 0);
        if (p.x_tma && __all_sync(0xffffffffu, xr0 >= 0 && xr0 == x0 + lane)) { issue_x(x0, ncol0); x_pend = true; }
      }
      for (int t = 1; t < p.seq; ++t) {
        const bool last_step = (t == p.seq - 1);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (s >= ns) continue;
          const int64_t mrow = mrow_s[s];
          const bool rvalid = mrow < p.M;
          const int xrow = xbase_s[s] < 0 ? -1 : xbase_s[s] + t;
          const int x0 = __shfl_sync(0xffffffffu, xrow, 0);
          const bool contig = p.x_tma && __all_sync(0xffffffffu, xrow >= 0 && xrow == x0 + lane);
          // the item after this one (same chunk): other slot at this step, or slot 0 at the next step
          int nxrow = -1;
          bool has_next = true;
          if (s + 1 < ns) nxrow = xbase_s[s + 1] < 0 ? -1 : xbase_s[s + 1] + t;
          else if (!last_step) nxrow = xbase_s[0] < 0 ? -1 : xbase_s[0] + t + 1;
          else has_next = false;                          // the next chunk issues its own first tile
          const int nx0 = __shfl_sync(0xffffffffu, nxrow, 0);
          const bool ncontig = has_next && p.x_tma && __all_sync(0xffffffffu, nxrow >= 0 && nxrow == nx0 + lane);

          const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + s * P_BN + colq * P_WCOLS;
          if (warp == 2 && lane == 0) PTL(item, 4);
          mbar_wait(&acc_full[s], use[s] & 1u);
          ++use[s];
          tc_fence_after();
          if (warp == 2 && lane == 0) PTL(item, 5);
          half_t* hdst = (t & 1) ? p.h16b : p.h16a;       // the last step's fp16 h is the relation block's query operand
          uint4 hpack = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int ch = 0; ch < P_NCH; ++ch) {
            const int cc = 32 * ch;
            const bool lastc = ch + 1 == P_NCH;
            const bool nx_ok = lastc ? ncontig : contig;
            const int nx_x0 = lastc ? nx0 : x0;
            const int nx_col = lastc ? ncol0 : ncol0 + cc + 32;
            uint32_t r[32];
            tmem_ld32(t_row + cc, r);
            tmem_ld_wait_dep(r);
            float* gsum = reinterpret_cast<float*>(r);
            if (x_pend) {
              mbar_wait(my_xfull, x_par);
              x_par ^= 1u;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 v = *reinterpret_cast<const float4*>(sb + lane * 32 + ((j ^ (lane & 7)) << 2));
                gsum[4 * j] += v.x; gsum[4 * j + 1] += v.y; gsum[4 * j + 2] += v.z; gsum[4 * j + 3] += v.w;
              }
              // WAR on the tile (my reads, then the next TMA write): no proxy fence (it lowers to MEMBAR.ALL.CTA, which
              // also waits for this thread's h stores to be acknowledged - measured ~10 % of the epilogue).  The reads
              // have PERFORMED once their values are consumed; the asm below pins the sums (hence the loads' results)
              // before the __syncwarp that precedes the next issue_x, as in any TMA pipeline's consumer release.
              pin_regs32(r);
            } else if (rvalid) {
              const float* xr = p.xp + (int64_t)xrow * (4 * kD) + ncol0 + cc;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(xr) + j);
                gsum[4 * j] += v.x; gsum[4 * j + 1] += v.y; gsum[4 * j + 2] += v.z; gsum[4 * j + 3] += v.w;
              }
            }
            __syncwarp();
            if (nx_ok) issue_x(nx_x0, nx_col);
            x_pend = nx_ok;
            if (lastc) {                                  // last TMEM read of this item: hand the accumulator back early
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_remote(&acc_empty[s], 0);
            }
            if (rvalid) {
              float hn[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float cn;
                lstm_cell_fast(gsum[4 * j], gsum[4 * j + 1], gsum[4 * j + 2], gsum[4 * j + 3], cst[s][8 * ch + j], cn, hn[j]);
                cst[s][8 * ch + j] = cn;
              }
              const int64_t o = mrow * kD + ((ncol0 + cc) >> 2);
              {
                // h feeds the next step's MMA (the last step's: the relation block's first GEMM): fp16; two 8-unit groups
                // leave as one 32-byte sector
                const uint2 lo = pack_h4(hn[0], hn[1], hn[2], hn[3]), hi = pack_h4(hn[4], hn[5], hn[6], hn[7]);
                if ((ch & 1) == 0) {
                  hpack = make_uint4(lo.x, lo.y, hi.x, hi.y);
                } else {
                  const uint32_t w8[8] = {hpack.x, hpack.y, hpack.z, hpack.w, lo.x, lo.y, hi.x, hi.y};
                  stg256u(hdst + o - 8, w8);
                }
              }
              if (last_step) stg256(p.h_out + o, hn);     // the last step's h is the clip's St: fp32
            }
          }
          if (!last_step) {                               // my 32 clips x 32 units of h_t are stored -> publisher
            __syncwarp();
            if (lane == 0) mbar_arrive(&h_done[s]);
          }
          if (warp == 2 && lane == 0) PTL(item, 6);
          ++item;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer may still multicast into this CTA's barriers
  if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm(tmem_base, 512); }
}

}  // namespace umma

#ifdef TMR_EXPERIMENT
}  // namespace tmr
extern "C" int tmr_debug_persist_timeline(unsigned long long* out_host, int n) {
  return cudaMemcpyFromSymbol(out_host, tmr::umma::g_ptl, sizeof(unsigned long long) * n) == cudaSuccess ? 0 : 2;
}
namespace tmr {
#endif

// Largest G (groups of 8 CTA pairs) the device can keep resident, 0 if the kernel cannot run here.
static int persist_groups() {
  static int groups = -1;
  if (groups >= 0) return groups;
  using namespace umma;
  groups = 0;
  if (cudaFuncSetAttribute(umma_lstm_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM) != cudaSuccess) {
    cudaGetLastError();
    return groups;
  }
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * P_SLICES * 9); cfg.blockDim = dim3(P_THREADS); cfg.dynamicSmemBytes = P_SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, umma_lstm_persist_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    clusters = 0;
  }
  if (clusters > sms / 2) clusters = sms / 2;
  groups = clusters / P_SLICES;
  return groups;
}

// Clips one full round of the persistent grid processes (2 tiles in flight x G groups x 256 clips), 0 if unavailable.
// A launch runs ceil(B / that) rounds and every round costs the same (a tile step is a latency chain), so a caller
// with a few hundred clips beyond a whole number of rounds does better handing them to the small-batch kernel.
int umma_lstm_persist_round_clips() { return 2 * persist_groups() * 2 * umma::P_BM; }

// Recurrent steps 1 .. seq-1 of every clip in one launch.  c0 / h16a hold the state after step 0; h16b is the
// second exchange buffer (on return the buffer of parity (seq - 1) & 1 also holds fp16(h_T)); flags: >= 2 * ceil(B / 256)
// int32 of scratch.  Returns TMR_ERR_UNSUPPORTED when the
// device cannot keep one group of 8 CTA pairs resident (the caller falls back to the per-step kernels).
int umma_lstm_persist(const half_t* whh16, const float* xp, const int64_t* starts, int seq, half_t* h16a, half_t* h16b,
                      float* h_out, const float* c0, int B, int32_t* flags, cudaStream_t st, const float* xp_base,
                      int64_t xp_rows, int64_t xp_row0) {
  using namespace umma;
  if (B == 0 || seq < 2) return TMR_OK;
  const int gmax = persist_groups();
  if (gmax < 1) {
    static bool warned = false;
    if (!warned) {                 // loud, once: the per-step kernels are correct but ~40 % slower
      warned = true;
      fprintf(stderr, "libtmr_b200: persistent LSTM recurrence unavailable on this device (fewer than 8 co-resident CTA "
                      "pairs); using the per-step kernels\n");
    }
    return set_error(TMR_ERR_UNSUPPORTED, "persistent LSTM recurrence: fewer than 8 co-resident CTA pairs");
  }
  LstmPersistParams p{};
  p.M = B; p.xp = xp; p.starts = starts; p.seq = seq; p.c0 = c0; p.h16a = h16a; p.h16b = h16b; p.h_out = h_out;
  p.x_row0 = xp_row0; p.flags = flags;
  p.m_pairs = ((int64_t)B + 2 * P_BM - 1) / (2 * P_BM);
  // two tiles in flight per pair: no more groups than pairs of tiles
  int64_t G = (p.m_pairs + 1) / 2;
  if (G > gmax) G = gmax;
  if (G < 1) G = 1;
  p.groups = (int)G;
  CUtensorMap th0, th1, tb, tx;
  {
    uint64_t da[2] = {(uint64_t)kD, (uint64_t)B};
    uint64_t sa[1] = {(uint64_t)kD * 2};
    uint32_t ba[2] = {P_BK, P_BM};
    TMR_TRY(make_tmap(&th0, h16a, 2, da, sa, ba, 2));
    TMR_TRY(make_tmap(&th1, h16b, 2, da, sa, ba, 2));
    uint64_t dw[2] = {(uint64_t)kD, (uint64_t)4 * kD};
    uint32_t bw[2] = {P_BK, P_BN / 2};
    TMR_TRY(make_tmap(&tb, whh16, 2, dw, sa, bw, 2));
    tx = th0;
    p.x_tma = 0;
    if (xp_base && xp_rows > 0) {                        // projected rows [xp_rows][4D], 32-clip x 32-column fp32 boxes
      uint64_t dx[2] = {(uint64_t)4 * kD, (uint64_t)xp_rows};
      uint64_t sx[1] = {(uint64_t)4 * kD * 4};
      uint32_t bx[2] = {32, 32};
      TMR_TRY(make_tmap(&tx, xp_base, 2, dx, sx, bx, 4, 128));
      p.x_tma = 1;
    }
  }
  TMR_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t) * 2 * p.m_pairs, st));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * P_SLICES * G)); cfg.blockDim = dim3(P_THREADS); cfg.dynamicSmemBytes = P_SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  static int coop_ok = 1;           // some drivers refuse cooperative + cluster launches: then the occupancy bound alone holds
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
  // (a refused launch would invalidate a stream capture: inside one, rely on the occupancy bound)
  if (coop_ok && cap == cudaStreamCaptureStatusNone) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, umma_lstm_persist_kernel, th0, th1, tb, tx, p);
    if (e == cudaSuccess) return TMR_OK;
    cudaGetLastError();
    coop_ok = 0;
  }
  cfg.numAttrs = 1;
  TMR_CUDA(cudaLaunchKernelEx(&cfg, umma_lstm_persist_kernel, th0, th1, tb, tx, p));
  return TMR_OK;
}

}  // namespace tmr
