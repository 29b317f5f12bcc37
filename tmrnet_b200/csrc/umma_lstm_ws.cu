// Recurrent LSTM step (TRAIN:224, 241-244), weights-stationary:  gates = h_{t-1} . Whh'^T + xp[frame],
// cell update in the epilogue.  One launch = one time step of every clip of the batch.
//
// What bounds this step is neither the MMAs (2.1 MFLOP per clip) nor HBM alone (14 KB per clip) but the bytes
// that cross L2 -> SM: with a streamed 256 x 256 x 512 tile per CTA pair, every tile pulls 256 KB of h rows AND
// 256 KB of Whh through L2 next to its 256 KB of projected rows (ncu: 1.13 GB of L2 reads per 41 600-clip launch
// against ~10 TB/s of aggregate L2 throughput; the main loop alone ran at that cap).  Whh' is only 2 MB in
// fp16, so here it never moves after kernel start: a CTA pair (cta_group::2, M = 256) owns ONE 256-column
// slice of the gates for the whole launch, each CTA keeping its 128 weight rows x K = 512 resident in shared
// memory (128 KB), and only h rows (four 16 KB stages), projected rows and c stream through.  Pairs are
// grouped by slice (pair p -> slice p mod 8) and walk the clip tiles round-robin inside their group, so the
// eight pairs that need the same h rows ask for them at about the same time (L2 hits).
//
// Warps: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2.. = epilogue (four per TMEM lane quarter, 64 gate
// columns each, in two 32-column chunks; or two per quarter with 128 columns).  Epilogue as in umma_gemm.cu's EPI_LSTM: TMEM lane = clip, a chunk
// = (i,f,g,o) of 8 hidden units, thread-private cell update; projected rows by the warp's own TMA load one
// chunk ahead; c by one 256-bit load one chunk ahead; c / h by 256 / 128-bit stores.
#include <stdlib.h>

#include "tmr_internal.h"
#include "umma_common.cuh"

namespace tmr {
namespace umma {

constexpr int WS_BM = 128;                    // rows per CTA (256 per pair)
constexpr int WS_BN = 256;                    // gate columns per pair; each CTA holds 128 of the weight rows
constexpr int WS_BK = 64;                     // fp16 per 128-byte swizzle row
constexpr int WS_KB = kD / WS_BK;             // 8 k-blocks
constexpr int WS_A_BYTES = WS_BM * WS_BK * 2; // 16 KB
constexpr int WS_B_BYTES = (WS_BN / 2) * WS_BK * 2;   // 16 KB per k-block per CTA
// Shared memory holds the resident weights (128 KB) plus 96 KB of landing zones to split between stages of h
// rows (16 KB each) and the epilogue warps' projected-row tiles (4 KB each): EW epilogue warps, NA stages.
// A tile holds 32 clips x CW gate columns (CW = 32: 4 KB, CW = 16: 2 KB).
template <int EW, int NA, int CW> constexpr int ws_smem_bytes() { return WS_KB * WS_B_BYTES + NA * WS_A_BYTES + EW * 128 * CW + 1024 + 512; }
constexpr int WS_N_SLICES = 4 * kD / WS_BN;   // 8

struct LstmWsParams {
  int64_t M;                       // clips
  const float* xp; const int64_t* starts; int seq; int t;
  float* h_out; half_t* h_out16; float* c;
  int x_tma; int64_t x_row0;       // tma_x covers the projected rows; its row 0 is projected row x_row0
  int pairs_per_slice;             // CTA pairs working on each 256-column slice
  int64_t m_pairs;                 // 256-row tiles
};

__device__ __forceinline__ int ws_xrow(const LstmWsParams& p, int64_t mr) {
  return (mr < p.M) ? (int)((p.starts ? p.starts[mr] : mr * p.seq) + p.t) : -1;
}

template <int WS_EPI_WARPS, int WS_NA, int CW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * WS_EPI_WARPS, 1)
umma_lstm_ws_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_x, const LstmWsParams p) {
  constexpr int WS_WCOLS = WS_BN * 4 / WS_EPI_WARPS;    // gate columns per epilogue warp (128 or 64)
  constexpr int WS_NCH = WS_WCOLS / CW;                 // chunks per warp and tile
  constexpr int TILE_FLOATS = 32 * CW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                                   // [WS_KB][128 weight rows][64 fp16], resident
  uint8_t* sA = sB + WS_KB * WS_B_BYTES;                // [WS_NA][128 clips][64 fp16]
  float* sX = reinterpret_cast<float*>(sA + WS_NA * WS_A_BYTES);   // [epilogue warps][32 clips x CW fp32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sX) + WS_EPI_WARPS * TILE_FLOATS * 4);
  uint64_t* a_full = bars;                    // [WS_NA]  TMA -> MMA (leader's barrier collects both CTAs' bytes)
  uint64_t* a_empty = bars + WS_NA;           // [WS_NA]  MMA -> TMA (multicast commit)
  uint64_t* b_full = bars + 2 * WS_NA;        // [1]
  uint64_t* acc_full = b_full + 1;            // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint64_t* xfull = acc_empty + 2;            // [WS_EPI_WARPS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull + WS_EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int slice = pair % WS_N_SLICES;                 // my 256 gate columns
  const int64_t mp0 = pair / WS_N_SLICES;               // first 256-row tile; stride = pairs_per_slice
  const int n0 = slice * WS_BN;
  constexpr uint16_t kMask = 3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_b); tma_prefetch_desc(&tma_x);
    for (int s = 0; s < WS_NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(b_full, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 2 * WS_EPI_WARPS); }
    for (int i = 0; i < WS_EPI_WARPS; ++i) mbar_init(&xfull[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // peer barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own weight rows once, own clip rows per tile) =============
    if (lane == 0) {
      if (crank == 0) mbar_expect_tx(b_full, 2 * WS_KB * WS_B_BYTES);
      for (int kb = 0; kb < WS_KB; ++kb)
        tma_load_2d_2sm(sB + kb * WS_B_BYTES, &tma_b, b_full, kb * WS_BK, n0 + (int)crank * (WS_BN / 2));
      int stage = 0; uint32_t phase = 0;
      for (int64_t mp = mp0; mp < p.m_pairs; mp += p.pairs_per_slice) {
        const int m0 = (int)(mp * 2 + crank) * WS_BM;
        for (int kb = 0; kb < WS_KB; ++kb) {
          mbar_wait(&a_empty[stage], phase ^ 1);
          if (crank == 0) mbar_expect_tx(&a_full[stage], 2 * WS_A_BYTES);
          tma_load_2d_2sm(sA + stage * WS_A_BYTES, &tma_a, &a_full[stage], kb * WS_BK, m0);
          if (++stage == WS_NA) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(2 * WS_BM, WS_BN);
      mbar_wait(b_full, 0);
      tc_fence_after();
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int64_t mp = mp0; mp < p.m_pairs; mp += p.pairs_per_slice, ++it) {
        const int acc = it & 1;
        mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);      // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * WS_BN;
        for (int kb = 0; kb < WS_KB; ++kb) {
          mbar_wait(&a_full[stage], phase);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(sA + stage * WS_A_BYTES));
          const uint64_t db = make_smem_desc_sw128(smem_u32(sB + kb * WS_B_BYTES));
#pragma unroll
          for (int k = 0; k < WS_BK / 16; ++k)
            mma_f16_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          mma_commit_2sm_mcast(&a_empty[stage], kMask);       // frees the h-row stage in both CTAs
          if (++stage == WS_NA) { stage = 0; phase ^= 1; }
        }
        mma_commit_2sm_mcast(&acc_full[acc], kMask);          // accumulator complete -> both CTAs' epilogues
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                             // TMEM lane quarter this warp may read
    const int colq = (warp - 2) >> 2;                   // which WS_WCOLS-wide part of the slice
    float* sb = sX + (warp - 2) * TILE_FLOATS;
    uint64_t* my_xfull = xfull + (warp - 2);
    bool x_pend = false;
    uint32_t x_par = 0;
    auto issue_x = [&](int x0, int col) {               // rows x0 .. x0+31, columns col .. col+CW-1 -> my tile
      if (lane == 0) {
        mbar_expect_tx(my_xfull, TILE_FLOATS * 4);
        tma_load_2d(sb, &tma_x, my_xfull, col, (int)(x0 - p.x_row0));
      }
    };
    auto tile_rows = [&](int64_t mp) -> int64_t { return (mp * 2 + crank) * WS_BM + q * 32; };
    const int ncol0 = n0 + colq * WS_WCOLS;             // first gate column of this warp
    int xrow_next = -1;
    float cpre[8];                                      // c of the next chunk (prefetched)
#pragma unroll
    for (int k = 0; k < 8; ++k) cpre[k] = 0.f;
    if (mp0 < p.m_pairs) {
      const int64_t mb = tile_rows(mp0);
      xrow_next = ws_xrow(p, mb + lane);
      const int x0 = __shfl_sync(0xffffffffu, xrow_next, 0);
      if (p.x_tma && __all_sync(0xffffffffu, xrow_next == x0 + lane)) { issue_x(x0, ncol0); x_pend = true; }
      if (mb + lane < p.M) ldg256(p.c + (mb + lane) * kD + (ncol0 >> 2), cpre);
    }
    int it = 0;
    for (int64_t mp = mp0; mp < p.m_pairs; mp += p.pairs_per_slice, ++it) {
      const int acc = it & 1;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * WS_BN + colq * WS_WCOLS;
      const int64_t mrow = tile_rows(mp) + lane;
      const bool rvalid = mrow < p.M;
      const int xrow = xrow_next;
      const int x0 = __shfl_sync(0xffffffffu, xrow, 0);
      const bool contig = p.x_tma && __all_sync(0xffffffffu, xrow == x0 + lane);
      const int64_t nmp = mp + p.pairs_per_slice;       // the next tile of this pair
      const bool has_next = nmp < p.m_pairs;
      int64_t nmrow = 0; int nx0 = 0; bool ncontig = false;
      if (has_next) {
        nmrow = tile_rows(nmp) + lane;
        xrow_next = ws_xrow(p, nmrow);
        nx0 = __shfl_sync(0xffffffffu, xrow_next, 0);
        ncontig = p.x_tma && __all_sync(0xffffffffu, xrow_next == nx0 + lane);
      }
      mbar_wait(&acc_full[acc], (it >> 1) & 1);
      tc_fence_after();
      float cin[8], cn[8], hn[8];
      uint4 hpack = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int ch = 0; ch < WS_NCH; ++ch) {
        constexpr int U = CW / 4;                       // hidden units per chunk (8 or 4)
        constexpr int CPG = 32 / CW;                    // chunks per 8-unit group: c / h move once per group
        const int cc = CW * ch;
        const int sub = ch % CPG;                       // position inside the group
        const bool last = ch + 1 == WS_NCH;
        uint32_t r[CW];
        if constexpr (CW == 32) tmem_ld32(t_row + cc, r); else tmem_ld16(t_row + cc, r);
        if (sub == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) cin[k] = cpre[k];
          // c of the group after this one
          if (ch + CPG < WS_NCH) { if (rvalid) ldg256(p.c + mrow * kD + ((ncol0 + cc + 32) >> 2), cpre); }
          else if (has_next && nmrow < p.M) ldg256(p.c + nmrow * kD + (ncol0 >> 2), cpre);
        }
        const bool nx_ok = last ? (has_next && ncontig) : contig;
        const int nx_x0 = last ? nx0 : x0;
        const int nx_col = last ? ncol0 : ncol0 + cc + CW;
        // gates = accumulator + projected row, summed in place in the accumulator registers
        if constexpr (CW == 32) tmem_ld_wait_dep(r); else tmem_ld_wait_dep16(r);
        float* gsum = reinterpret_cast<float*>(r);
        if (x_pend) {
          mbar_wait(my_xfull, x_par);
          x_par ^= 1u;
#pragma unroll
          for (int j = 0; j < U; ++j) {
            // 32-column tiles: 128-byte rows, SWIZZLE_128B (16-byte chunk ^= row & 7); 16-column tiles: 64-byte
            // rows, SWIZZLE_64B (chunk ^= (row >> 1) & 3).  Either way the 8 lanes of a quarter warp hit 8
            // different 16-byte bank groups.
            const float4 v = (CW == 32)
                ? *reinterpret_cast<const float4*>(sb + lane * 32 + ((j ^ (lane & 7)) << 2))
                : *reinterpret_cast<const float4*>(sb + lane * 16 + ((j ^ ((lane >> 1) & 3)) << 2));
            gsum[4 * j] += v.x; gsum[4 * j + 1] += v.y; gsum[4 * j + 2] += v.z; gsum[4 * j + 3] += v.w;
          }
          fence_proxy_async_smem();                     // the tile's generic reads before its next TMA write
        } else if (rvalid) {
          const float* xr = p.xp + (int64_t)xrow * (4 * kD) + ncol0 + cc;
#pragma unroll
          for (int j = 0; j < U; ++j) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xr) + j);
            gsum[4 * j] += v.x; gsum[4 * j + 1] += v.y; gsum[4 * j + 2] += v.z; gsum[4 * j + 3] += v.w;
          }
        }
        __syncwarp();
        if (nx_ok) issue_x(nx_x0, nx_col);
        x_pend = nx_ok;
        if (last) {                                     // last TMEM read of this tile: hand the accumulator back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&acc_empty[acc], 0);
        }
        if (rvalid) {
#pragma unroll
          for (int j = 0; j < U; ++j)
            lstm_cell_fast(gsum[4 * j], gsum[4 * j + 1], gsum[4 * j + 2], gsum[4 * j + 3], cin[U * sub + j], cn[U * sub + j], hn[U * sub + j]);
          if (sub == CPG - 1) {
            const int64_t o = mrow * kD + ((ncol0 + cc + CW - 32) >> 2);
            stg256(p.c + o, cn);
            // h only feeds the next step's MMA: fp16; the last step's h is the clip's St: fp32.  The step pays
            // per L2 request, so the fp16 h of TWO 8-unit groups leaves as one 32-byte sector.
            if (p.h_out16) {
              const uint2 lo = pack_h4(hn[0], hn[1], hn[2], hn[3]), hi = pack_h4(hn[4], hn[5], hn[6], hn[7]);
              if (((cc + CW - 32) >> 5) % 2 == 0) {
                hpack = make_uint4(lo.x, lo.y, hi.x, hi.y);
              } else {
                const uint32_t w8[8] = {hpack.x, hpack.y, hpack.z, hpack.w, lo.x, lo.y, hi.x, hi.y};
                stg256u(p.h_out16 + o - 8, w8);
              }
            } else {
              stg256(p.h_out + o, hn);
            }
          }
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

int umma_lstm_step_ws(const half_t* whh16, const float* xp, const int64_t* starts, int seq, int t, const half_t* h_prev,
                      half_t* h_out16, float* h_out, float* c, int B, cudaStream_t st, const float* xp_base,
                      int64_t xp_rows, int64_t xp_row0) {
  using namespace umma;
  if (B == 0) return TMR_OK;
  LstmWsParams p{};
  p.M = B; p.xp = xp; p.starts = starts; p.seq = seq; p.t = t; p.h_out = h_out; p.h_out16 = h_out16; p.c = c;
  p.x_row0 = xp_row0;
  p.m_pairs = ((int64_t)B + 2 * WS_BM - 1) / (2 * WS_BM);
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int pps = (sms / 2) / WS_N_SLICES;                     // 9 on a 148-SM B200
  if (pps < 1) pps = 1;
  if ((int64_t)pps > p.m_pairs) pps = (int)p.m_pairs;
  p.pairs_per_slice = pps;
  static const int cfg = env_int("TMR_LSTM_WS_CFG", 8);
  CUtensorMap ta, tb, tx;
  {
    uint64_t da[2] = {(uint64_t)kD, (uint64_t)B};
    uint64_t sa[1] = {(uint64_t)kD * 2};
    uint32_t ba[2] = {WS_BK, WS_BM};
    TMR_TRY(make_tmap(&ta, h_prev, 2, da, sa, ba, 2));
    uint64_t dw[2] = {(uint64_t)kD, (uint64_t)4 * kD};
    uint32_t bw[2] = {WS_BK, WS_BN / 2};
    TMR_TRY(make_tmap(&tb, whh16, 2, dw, sa, bw, 2));
    tx = ta;
    p.x_tma = 0;
    if (xp_base && xp_rows > 0) {                        // projected rows [xp_rows][4D], 32-clip x CW-column fp32 boxes per epilogue warp
      uint64_t dx[2] = {(uint64_t)4 * kD, (uint64_t)xp_rows};
      uint64_t sx[1] = {(uint64_t)4 * kD * 4};
      uint32_t bx[2] = {(uint32_t)(cfg == 8 ? 32 : 16), 32};
      TMR_TRY(make_tmap(&tx, xp_base, 2, dx, sx, bx, 4, cfg == 8 ? 128 : 64));
      p.x_tma = 1;
    }
  }
  if (cfg == 8) {          // 8 epilogue warps x 128 columns, 32-column chunks
    TMR_CUDA(cudaFuncSetAttribute(umma_lstm_ws_kernel<8, 4, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ws_smem_bytes<8, 4, 32>()));
    umma_lstm_ws_kernel<8, 4, 32><<<2 * WS_N_SLICES * pps, 64 + 32 * 8, ws_smem_bytes<8, 4, 32>(), st>>>(ta, tb, tx, p);
  } else {                 // 16 epilogue warps x 64 columns, 16-column chunks (2 KB tiles leave room for 4 stages)
    TMR_CUDA(cudaFuncSetAttribute(umma_lstm_ws_kernel<16, 4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ws_smem_bytes<16, 4, 16>()));
    umma_lstm_ws_kernel<16, 4, 16><<<2 * WS_N_SLICES * pps, 64 + 32 * 16, ws_smem_bytes<16, 4, 16>(), st>>>(ta, tb, tx, p);
  }
  TMR_LAUNCH_CHECK("umma_lstm_ws_kernel");
  return TMR_OK;
}

}  // namespace tmr
