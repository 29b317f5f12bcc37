// TMR_MATH_F16 GEMM engine: persistent, warp-specialised tcgen05 kernel.
//   C[M,N] = A[M,K] . W[N,K]^T   (both K-major fp16 in HBM, fp32 accumulate in TMEM)
// Tile 128 x 256 x 64: TMA (SWIZZLE_128B) stages A (16 KB) + W (32 KB) per k-block through a 4-deep
// mbarrier ring; one elected thread issues four tcgen05.mma.kind::f16 (K=16) per stage; accumulators
// are double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the main loop of
// tile i+1.  Warps: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..9 = epilogue (tcgen05.ld; two
// warps per TMEM lane quarter, one per 128-column half, so every SM sub-partition has two warps to
// hide the epilogue's load / MUFU latencies behind each other).
// Epilogues: EPI_LINEAR (bias / residual / relu -> out) and EPI_LSTM (gate-interleaved LSTM cell).
#define TMR_HAVE_UMMA 1
#include <stdlib.h>

#include "tmr_internal.h"
#include <type_traits>
#include "umma_common.cuh"

namespace tmr {
namespace umma {

int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int elem_bytes, int swizzle_bytes) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
      return set_error(TMR_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
    encode = (EncodeFn)fn;
  }
  cuuint64_t gdims[3] = {1, 1, 1};
  cuuint64_t gstr[2] = {0, 0};
  cuuint32_t gbox[3] = {1, 1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; gbox[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const CUtensorMapDataType dt = (elem_bytes == 2) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = encode(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE,
                      swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(TMR_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu,%llu,%llu box %u,%u,%u)",
                     (int)r, rank, (unsigned long long)gdims[0], (unsigned long long)gdims[1],
                     (unsigned long long)gdims[2], gbox[0], gbox[1], gbox[2]);
  return TMR_OK;
}

constexpr int BM = 128, BN = 256, BK = 64;           // BK fp16 = 128 bytes = one swizzle row
enum { EPI_LINEAR = 0, EPI_LSTM = 1 };
// Per-epilogue configuration.  The LSTM-cell epilogue, not the MMAs, bounds the recurrent step, so it gets
// 16 epilogue warps (4 per SM sub-partition), each with XBUF 32x32 fp32 smem tiles: the landing zones of the
// warp's own TMA loads of projected rows.  XBUF = 2 (a second tile for two of the five stages) measured the same
// 143 us per 41 600 clips as XBUF = 1 and is not instantiated.  The plain epilogue keeps 8 warps (one 4 KB
// transpose tile each) and 4 / 6 stages.  Batches of >= 256 clips take umma_lstm_ws.cu instead of EPI_LSTM here.
template <int EPI, int XBUF = 1> struct Cfg {
  static constexpr int STAGES = (EPI == EPI_LSTM) ? 3 : 4;
  static constexpr int STAGES_2SM = (EPI == EPI_LSTM) ? (XBUF == 2 ? 3 : 5) : 6;      // 32 KB stages in 2-SM mode
  static constexpr int EPI_WARPS = (EPI == EPI_LSTM) ? 16 : 8;
  static constexpr int NTHREADS = 64 + 32 * EPI_WARPS;
  static constexpr int EPI_TILES = (EPI == EPI_LSTM) ? XBUF : 1;   // 32x32 fp32 smem tiles per epilogue warp
  static constexpr int EPI_STAGE_BYTES = EPI_WARPS * EPI_TILES * 4096;
  static constexpr int COLS_PER_WARP = 1024 / EPI_WARPS;         // 4 warps per TMEM lane quarter share 256 columns
};
constexpr int A_BYTES = BM * BK * 2;                 // 16 KB
constexpr int B_BYTES = BN * BK * 2;                 // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
template <int EPI, bool TWOSM = false, int XBUF = 1> constexpr int smem_bytes() {
  return (TWOSM ? Cfg<EPI, XBUF>::STAGES_2SM * (A_BYTES + B_BYTES / 2) : Cfg<EPI, XBUF>::STAGES * STAGE_BYTES) +
         Cfg<EPI, XBUF>::EPI_STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
}
constexpr int TMEM_COLS = 512;

// Debug timeline (TMR_TIMELINE=1): per CTA and tile, clock64 stamps of the MMA thread (wait for a free
// accumulator, main loop start/end) and of epilogue warp 2 (accumulator ready, epilogue done).
__device__ long long g_timeline[148 * 16 * 6];
__device__ long long g_timeline_cta[148 * 4];          // per CTA: kernel entry, setup done, roles done, exit (clock64) 
__device__ long long g_timeline_warps[16 * 16 * 4];   // CTA 0: [tile][epilogue warp][wait start, acc ready, chunk 0 done, done]

struct GemmParams {
  int64_t M; int N; int K; int k_split;
  // EPI_LINEAR
  const float* bias; const float* residual; int64_t ldr; float* out; int64_t ldo; int relu;
  // EPI_LINEAR as the LSTM input projection (gate-interleaved columns): rows that start a clip also get step 0 of
  // that clip's recurrence from zero state, c0 = sig(i) tanh(g), h0 = sig(o) tanh(c0), written as c0[clip] / h0[clip]
  const int32_t* row2clip; float* c0; half_t* h0_16;
  // EPI_LSTM (N = 4*512 gate-interleaved columns)
  const float* xp; const int64_t* starts; int seq; int t; float* h_out; half_t* h_out16; float* c;
  int x_tma; int64_t x_row0;          // tma_x covers the projected rows; its row 0 is projected row x_row0
  const float* xp_base; int64_t xp_rows;   // host side only: what tma_x is built over
  int timeline;
  int stages;      // > 0: use only this many pipeline stages (experiments)
};

__device__ __forceinline__ int lstm_xrow(const GemmParams& p, int64_t mr) {
  return (mr < p.M) ? (int)((p.starts ? p.starts[mr] : mr * p.seq) + p.t) : -1;
}
// CL = 1: independent CTAs.  CL = 2: clusters of two CTAs on adjacent M tiles of the same N tile; each
// CTA loads its own A tile and HALF of the shared W tile and multicasts that half into both CTAs, so
// weight traffic from L2 per CTA halves (48 -> 32 KB per k-block).  A stage may only be refilled when
// BOTH consumers have released it: every tcgen05.commit on a stage arrives on both CTAs' empty barrier.
// CL = 3: CTA PAIRS with 2-SM MMA (tcgen05 cta_group::2): the pair computes a 256 x 256 tile, each CTA
// stages only its own 128 rows of A and its own 128 rows (N half) of W — 32 KB instead of 48 KB per
// k-block through the SM's ~55 B/cycle ingest.  The
// leader CTA's thread issues the MMAs for both; both CTAs' TMA loads complete on the leader's full
// barrier; tcgen05.commit multicasts stage releases and accumulator-ready signals to both CTAs; both
// CTAs' epilogues report to the leader's accumulator-empty barrier.
template <int EPI, int CL, int XBUF = 1>
__global__ void __launch_bounds__(Cfg<EPI>::NTHREADS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_a2,
                 const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_x,
                 const GemmParams p) {
  constexpr bool TWOSM = (CL == 3);
  constexpr int CSIZE = (CL == 1) ? 1 : 2;                       // CTAs per cluster
  constexpr int STAGES = TWOSM ? Cfg<EPI, XBUF>::STAGES_2SM : Cfg<EPI, XBUF>::STAGES;
  constexpr int STAGE_BYTES = TWOSM ? (A_BYTES + B_BYTES / 2) : (A_BYTES + B_BYTES);
  constexpr int EPI_WARPS = Cfg<EPI>::EPI_WARPS;
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the __shared__ array (no integer round trip) keeps the shared address space, so the
  // epilogue staging compiles to STS/LDS instead of generic ST.E/LD.E
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + Cfg<EPI, XBUF>::EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;    // [2]       MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]    epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* xfull = bars + 2 * STAGES + 6;   // [EPI_WARPS][XBUF]  EPI_LSTM: a warp's own projected-row tile has landed

  const int NST = (p.stages > 0 && p.stages < STAGES) ? p.stages : STAGES;
  if (p.timeline && threadIdx.x == 0 && blockIdx.x < 148) g_timeline_cta[blockIdx.x * 4 + 0] = clock64();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int64_t m_tiles = (p.M + BM - 1) / BM;
  const int k_blocks = p.K / BK;
  // work items: CL consecutive M tiles x one N tile; CTA `crank` of the cluster takes M tile CL*mp + crank
  const uint32_t crank = (CSIZE > 1) ? cluster_ctarank() : 0;
  const int64_t num_items = ((m_tiles + CSIZE - 1) / CSIZE) * n_tiles;
  const int64_t item0 = blockIdx.x / CSIZE;
  const int64_t item_stride = gridDim.x / CSIZE;
  constexpr uint16_t kMask = (uint16_t)((1u << CSIZE) - 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_a2); tma_prefetch_desc(&tma_b);
    // empty: CL=2 both consumers release a stage (2 arrivals); 2-SM: one multicast commit per CTA
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CL == 2 ? 2 : 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], TWOSM ? 2 * EPI_WARPS : EPI_WARPS); }
    if (EPI == EPI_LSTM) { tma_prefetch_desc(&tma_x); for (int i = 0; i < EPI_WARPS * XBUF; ++i) mbar_init(&xfull[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) { if (TWOSM) tmem_alloc_2sm(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS); }
  tc_fence_before();
  __syncthreads();
  if (CSIZE > 1) cluster_sync_all();       // peer barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.timeline && threadIdx.x == 0 && blockIdx.x < 148) g_timeline_cta[blockIdx.x * 4 + 1] = clock64();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t item = item0; item < num_items; item += item_stride) {
        const int m0 = (int)((item / n_tiles) * CSIZE + crank) * BM;
        const int n0 = (int)(item % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);       // CL = 2: released by BOTH CTAs' consumers
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = kb * BK;
          if (TWOSM) {
            // both CTAs' bytes (2 x 32 KB) land on the LEADER's barrier, which only the leader arms
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
            if (k0 < p.k_split) tma_load_2d_2sm(sa, &tma_a, &full_bar[stage], k0, m0);
            else tma_load_2d_2sm(sa, &tma_a2, &full_bar[stage], k0 - p.k_split, m0);
            tma_load_2d_2sm(sb, &tma_b, &full_bar[stage], k0, n0 + (int)crank * (BN / 2));   // my half of W's N rows
            if (++stage == NST) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (k0 < p.k_split) tma_load_2d(sa, &tma_a, &full_bar[stage], k0, m0);
          else tma_load_2d(sa, &tma_a2, &full_bar[stage], k0 - p.k_split, m0);
          if (CL == 1) {
            tma_load_2d(sb, &tma_b, &full_bar[stage], k0, n0);
            tma_load_2d(sb + B_BYTES / 2, &tma_b, &full_bar[stage], k0, n0 + BN / 2);
          } else {   // my half of the W tile, written into both CTAs
            tma_load_2d_mcast(sb + crank * (B_BYTES / 2), &tma_b, &full_bar[stage], k0, n0 + (int)crank * (BN / 2), kMask);
          }
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && (!TWOSM || crank == 0)) {                 // 2-SM: only the leader CTA issues
      constexpr uint32_t idesc = make_idesc_f16(TWOSM ? 2 * BM : BM, BN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int64_t item = item0; item < num_items; item += item_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        long long tl0 = 0;
        if (p.timeline && it < 16) tl0 = clock64();
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        if (p.timeline && it < 16) { long long* tl = g_timeline + ((int64_t)blockIdx.x * 16 + it) * 6; tl[0] = tl0; tl[1] = clock64(); }
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {           // UMMA_K = 16 fp16 = 32 bytes inside the swizzle row
            if (TWOSM) mma_f16_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            else mma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          if (CL == 1) mma_commit(&empty_bar[stage]);   // frees the smem stage when these MMAs retire
          else if (TWOSM) mma_commit_2sm_mcast(&empty_bar[stage], kMask);
          else mma_commit_mcast(&empty_bar[stage], kMask);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        if (TWOSM) mma_commit_2sm_mcast(&acc_full[acc], kMask);   // accumulator complete -> both CTAs' epilogues
        else mma_commit(&acc_full[acc]);
        if (p.timeline && it < 16) g_timeline[((int64_t)blockIdx.x * 16 + it) * 6 + 2] = clock64();
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                             // TMEM lane quarter this warp may read
    constexpr int WCOLS = Cfg<EPI>::COLS_PER_WARP;      // accumulator columns owned by this warp (128 or 64)
    const int colq = (warp - 2) >> 2;                   // which WCOLS-wide slice of the 256 columns
    if constexpr (EPI == EPI_LSTM) {
      // LSTM cell on gate-interleaved columns.  TMEM hands every thread one accumulator ROW (lane = clip) and a
      // 32-column chunk = the (i,f,g,o) gates of 8 hidden units, so the whole cell update of those units is
      // thread-private: no transpose through shared memory.  What the thread needs beside the accumulator:
      //  * its clip's projected row (bias folded), 128 contiguous bytes per chunk: when the warp's 32 clips read
      //    32 CONSECUTIVE projected rows (every warp except those straddling a video boundary or the end of the
      //    batch) the 32x32 tile arrives by the warp's OWN TMA load, XBUF chunks ahead, in a SWIZZLE_128B tile
      //    the thread reads back with eight conflict-free 128-bit loads; otherwise eight LDG.128 of its own row;
      //  * c of the 8 units: ONE 256-bit load (a whole 32-byte sector per thread), fetched a chunk ahead;
      //  * c out: one 256-bit store; h out: 8 fp16 = one 128-bit store (fp32 / 256-bit for the last step).
      // The epilogue, not the MMAs, bounds this kernel; its cost is latency (HBM round trips), so everything it
      // reads is requested one chunk (c) or XBUF chunks (projected rows) before it is used.
      constexpr int NCH = WCOLS / 32;                   // chunks per tile and warp (2)
      float* sbuf0 = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES) + (warp - 2) * (XBUF * 1024);
      uint64_t* my_xfull = xfull + (warp - 2) * XBUF;
      uint32_t x_pend = 0, x_par = 0;                   // bit b: buffer b has a TMA load in flight / its barrier parity
      auto issue_x = [&](int b, int x0, int col) {      // rows x0 .. x0+31, columns col .. col+31 -> tile b
        if (lane == 0) {
          mbar_expect_tx(my_xfull + b, 4096);
          tma_load_2d(sbuf0 + b * 1024, &tma_x, my_xfull + b, col, (int)(x0 - p.x_row0));
        }
      };
      auto tile_rows = [&](int64_t it_item) -> int64_t { return ((it_item / n_tiles) * CSIZE + crank) * BM + q * 32; };
      // per-tile state of the NEXT tile, computed a tile ahead: projected-row index of this lane's clip
      int xrow_next = -1;
      float cpre[8];                                     // c of the next chunk (prefetched)
#pragma unroll
      for (int k = 0; k < 8; ++k) cpre[k] = 0.f;
      if (item0 < num_items) {
        const int64_t mb = tile_rows(item0);
        xrow_next = lstm_xrow(p, mb + lane);
        const int x0 = __shfl_sync(0xffffffffu, xrow_next, 0);
        const int ncol = (int)(item0 % n_tiles) * BN + colq * WCOLS;
        if (p.x_tma && __all_sync(0xffffffffu, xrow_next == x0 + lane)) {
#pragma unroll
          for (int b = 0; b < XBUF; ++b) { issue_x(b, x0, ncol + 32 * b); x_pend |= 1u << b; }
        }
        if (mb + lane < p.M) ldg256(p.c + (mb + lane) * kD + (ncol >> 2), cpre);
      }
      int it = 0;
      for (int64_t item = item0; item < num_items; item += item_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int64_t m_base = tile_rows(item);
        const int n0 = (int)(item % n_tiles) * BN + colq * WCOLS;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + colq * WCOLS;
        const int64_t mrow = m_base + lane;
        const bool rvalid = mrow < p.M;
        const int xrow = xrow_next;
        const int x0 = __shfl_sync(0xffffffffu, xrow, 0);
        const bool contig = p.x_tma && __all_sync(0xffffffffu, xrow == x0 + lane);
        // the next tile of this warp
        const int64_t nitem = item + item_stride;
        const bool has_next = nitem < num_items;
        int64_t nm_base = 0; int nn0 = 0; int nx0 = 0; bool ncontig = false;
        if (has_next) {
          nm_base = tile_rows(nitem);
          nn0 = (int)(nitem % n_tiles) * BN + colq * WCOLS;
          xrow_next = lstm_xrow(p, nm_base + lane);
          nx0 = __shfl_sync(0xffffffffu, xrow_next, 0);
          ncontig = p.x_tma && __all_sync(0xffffffffu, xrow_next == nx0 + lane);
        }
        if (p.timeline && it < 16 && warp == 2 && lane == 0) g_timeline[((int64_t)blockIdx.x * 16 + it) * 6 + 3] = clock64();
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
        if (p.timeline && it < 16 && warp == 2 && lane == 0) g_timeline[((int64_t)blockIdx.x * 16 + it) * 6 + 4] = clock64();
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int cc = 32 * ch;
          const int b = (XBUF == 2) ? ch : 0;
          float* sb = sbuf0 + b * 1024;
          uint32_t r[32];
          tmem_ld32(t_row + cc, r);
          float cin[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) cin[k] = cpre[k];
          // c of the chunk after this one
          if (ch + 1 < NCH) { if (rvalid) ldg256(p.c + mrow * kD + ((n0 + cc + 32) >> 2), cpre); }
          else if (has_next && nm_base + lane < p.M) ldg256(p.c + (nm_base + lane) * kD + (nn0 >> 2), cpre);
          // the chunk whose projected rows go into tile b once this chunk has read it: XBUF chunks ahead
          bool nx_ok; int nx_x0, nx_col;
          if (XBUF == 1 && ch + 1 < NCH) { nx_ok = contig; nx_x0 = x0; nx_col = n0 + cc + 32; }
          else { nx_ok = has_next && ncontig; nx_x0 = nx0; nx_col = nn0 + ((XBUF == 2) ? cc : 0); }
          float4 g[8];                                    // (i,f,g,o) of the chunk's 8 units, projected row first
          if (x_pend & (1u << b)) {
            mbar_wait(my_xfull + b, (x_par >> b) & 1u);
            x_par ^= 1u << b;
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = *reinterpret_cast<const float4*>(sb + lane * 32 + ((j ^ (lane & 7)) << 2));
            fence_proxy_async_smem();                     // the tile's generic reads before its next TMA write
          } else {
            const float* xr = p.xp + (int64_t)(rvalid ? xrow : 0) * (4 * kD) + n0 + cc;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              g[j] = rvalid ? __ldg(reinterpret_cast<const float4*>(xr) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          __syncwarp();
          if (nx_ok) { issue_x(b, nx_x0, nx_col); x_pend |= 1u << b; } else { x_pend &= ~(1u << b); }
          tmem_ld_wait_dep(r);
          if (ch + 1 == NCH) {                            // last TMEM read of this tile: hand the accumulator back early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (TWOSM) mbar_arrive_remote(&acc_empty[acc], 0); else mbar_arrive(&acc_empty[acc]); }
          }
          if (rvalid) {
            float cn[8], hn[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              lstm_cell_fast(g[j].x + __uint_as_float(r[4 * j]), g[j].y + __uint_as_float(r[4 * j + 1]),
                             g[j].z + __uint_as_float(r[4 * j + 2]), g[j].w + __uint_as_float(r[4 * j + 3]), cin[j], cn[j], hn[j]);
            const int64_t o = mrow * kD + ((n0 + cc) >> 2);
            stg256(p.c + o, cn);
            // h only feeds the next step's MMA: fp16; the last step's h is the clip's St: fp32
            if (p.h_out16) {
              const uint2 lo = pack_h4(hn[0], hn[1], hn[2], hn[3]), hi = pack_h4(hn[4], hn[5], hn[6], hn[7]);
              *reinterpret_cast<uint4*>(p.h_out16 + o) = make_uint4(lo.x, lo.y, hi.x, hi.y);
            } else {
              stg256(p.h_out + o, hn);
            }
          }
        }
        if (p.timeline && it < 16 && warp == 2 && lane == 0) g_timeline[((int64_t)blockIdx.x * 16 + it) * 6 + 5] = clock64();
      }
    } else {
      // Plain epilogue.  A row-per-thread global access pattern touches 32 different cache lines per warp
      // instruction, so each 32x32 chunk goes TMEM -> registers -> a 4 KB XOR-swizzled smem tile (phase A) and
      // the bias / residual / relu + global I/O happen in a COALESCED layout (phase B): per instruction 8 lanes
      // cover the 128 contiguous bytes of one row, 4 rows per warp instruction.
      float* sbuf = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES) + (warp - 2) * 1024;   // 32 rows x 32 floats
      const int prow = lane >> 3;                         // phase B: row within a group of 4
      const int pch = lane & 7;                           // phase B: 16-byte chunk (4 columns) of the row
      int it = 0;
      for (int64_t item = item0; item < num_items; item += item_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int64_t m_base = ((item / n_tiles) * CSIZE + crank) * BM + q * 32;  // first row of this warp
        const int n0 = (int)(item % n_tiles) * BN + colq * WCOLS;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + colq * WCOLS;
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < WCOLS; cc += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + cc, r);
          tmem_ld_wait();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)                     // phase A: this thread's row -> swizzled smem (conflict-free)
            *reinterpret_cast<uint4*>(sbuf + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          __syncwarp();
          const int n = n0 + cc + 4 * pch;
          if (n < p.N) {                                  // N % 4 == 0
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = 4 * i + prow;
              const int64_t mr = m_base + row;
              if (mr < p.M) {
                float4 v = *reinterpret_cast<const float4*>(sbuf + row * 32 + ((pch ^ (row & 7)) << 2));
                v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
                if (p.residual) {
                  const float4 e = __ldg(reinterpret_cast<const float4*>(p.residual + mr * p.ldr + n));
                  v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
                }
                if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                *reinterpret_cast<float4*>(p.out + mr * p.ldo + n) = v;
                if (p.row2clip) {                       // v = (i, f, g, o) of hidden unit n / 4; f * c_{-1} = 0 drops out
                  const int clip = __ldg(p.row2clip + mr);
                  if (clip >= 0) {
                    const float a = ex2_approx(fminf(-1.4426950408889634f * v.x, 40.f));
                    const float d = ex2_approx(fminf(-2.8853900817779268f * v.z, 40.f));
                    const float e = ex2_approx(fminf(-1.4426950408889634f * v.w, 40.f));
                    const float cn = (1.f - d) * rcp_approx((1.f + a) * (1.f + d));          // sigmoid(i) tanh(g)
                    const float f2 = ex2_approx(fminf(-2.8853900817779268f * cn, 40.f));
                    const float hn = (1.f - f2) * rcp_approx((1.f + e) * (1.f + f2));        // sigmoid(o) tanh(c)
                    const int64_t o = (int64_t)clip * kD + (n >> 2);
                    p.c0[o] = cn;
                    reinterpret_cast<uint16_t*>(p.h0_16)[o] = (uint16_t)(pack_h2(hn, 0.f) & 0xffffu);
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (TWOSM) mbar_arrive_remote(&acc_empty[acc], 0); else mbar_arrive(&acc_empty[acc]); }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.timeline && threadIdx.x == 0 && blockIdx.x < 148) g_timeline_cta[blockIdx.x * 4 + 2] = clock64();
  if (CSIZE > 1) cluster_sync_all();       // the peer may still multicast into this CTA's smem / barriers
  if (warp == 1) { tc_fence_after(); if (TWOSM) tmem_dealloc_2sm(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
  if (p.timeline && threadIdx.x == 0 && blockIdx.x < 148) g_timeline_cta[blockIdx.x * 4 + 3] = clock64();
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int EPI>
static int launch_gemm(const half_t* a, int64_t lda, const half_t* a2, int64_t lda2, int k_split, const half_t* w,
                       int64_t ldw, const GemmParams& p, cudaStream_t st) {
  CUtensorMap ta, ta2, tb, tx;
  {
    const int ka = a2 ? k_split : p.K;
    uint64_t dims[2] = {(uint64_t)ka, (uint64_t)p.M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {BK, BM};
    TMR_TRY(make_tmap(&ta, a, 2, dims, str, box, 2));
    if (a2) {
      uint64_t d2[2] = {(uint64_t)(p.K - k_split), (uint64_t)p.M};
      uint64_t s2[1] = {(uint64_t)lda2 * 2};
      TMR_TRY(make_tmap(&ta2, a2, 2, d2, s2, box, 2));
    } else {
      ta2 = ta;
    }
    uint64_t dw[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t sw[1] = {(uint64_t)ldw * 2};
    uint32_t bw[2] = {BK, BN / 2};                 // W tiles are fetched as two 128-row halves
    TMR_TRY(make_tmap(&tb, w, 2, dw, sw, bw, 2));
    tx = ta;
    const_cast<GemmParams&>(p).x_tma = 0;
    if (EPI == EPI_LSTM && p.xp_base && p.xp_rows > 0) {   // projected rows [xp_rows][N], 32 x 32 boxes per epilogue warp
      uint64_t dx[2] = {(uint64_t)p.N, (uint64_t)p.xp_rows};
      uint64_t sx[1] = {(uint64_t)p.N * 4};
      uint32_t bx[2] = {32, 32};
      TMR_TRY(make_tmap(&tx, p.xp_base, 2, dx, sx, bx, 4));
      const_cast<GemmParams&>(p).x_tma = 1;
    }
  }
  static const int tlflag = env_int("TMR_TIMELINE", 0);
  const_cast<GemmParams&>(p).timeline = tlflag;
  static const int lstm_stages = env_int("TMR_LSTM_STAGES", 0);
  const_cast<GemmParams&>(p).stages = (EPI == EPI_LSTM) ? lstm_stages : 0;
  static const int cluster = env_int("TMR_GEMM_CLUSTER", 3);
  const int64_t m_tiles = (p.M + BM - 1) / BM;
  const int64_t n_tiles = (p.N + BN - 1) / BN;
  if ((cluster == 2 || cluster == 3) && m_tiles >= 2) {
    const int64_t items = ((m_tiles + 1) / 2) * n_tiles;
    const int max_clusters = num_sms() / 2;
    const int clusters = (int)(items < max_clusters ? items : max_clusters);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(Cfg<EPI>::NTHREADS); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cluster == 3) {
      cfg.dynamicSmemBytes = smem_bytes<EPI, true>();
      TMR_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<EPI, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<EPI, true>()));
      TMR_CUDA(cudaLaunchKernelEx(&cfg, umma_gemm_kernel<EPI, 3>, ta, ta2, tb, tx, p));
    } else {
      cfg.dynamicSmemBytes = smem_bytes<EPI>();
      TMR_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<EPI, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<EPI>()));
      TMR_CUDA(cudaLaunchKernelEx(&cfg, umma_gemm_kernel<EPI, 2>, ta, ta2, tb, tx, p));
    }
    return TMR_OK;
  }
  TMR_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<EPI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<EPI>()));
  const int64_t tiles = m_tiles * n_tiles;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  umma_gemm_kernel<EPI, 1><<<grid, Cfg<EPI>::NTHREADS, smem_bytes<EPI>(), st>>>(ta, ta2, tb, tx, p);
  TMR_LAUNCH_CHECK("umma_gemm_kernel");
  return TMR_OK;
}

}  // namespace umma

}  // namespace tmr
extern "C" int tmr_debug_timeline_cta(long long* out_host, int n) {
  return cudaMemcpyFromSymbol(out_host, tmr::umma::g_timeline_cta, sizeof(long long) * n) == cudaSuccess ? 0 : 2;
}
extern "C" int tmr_debug_timeline_warps(long long* out_host, int n) {
  return cudaMemcpyFromSymbol(out_host, tmr::umma::g_timeline_warps, sizeof(long long) * n) == cudaSuccess ? 0 : 2;
}
extern "C" int tmr_debug_timeline(long long* out_host, int n) {
  return cudaMemcpyFromSymbol(out_host, tmr::umma::g_timeline, sizeof(long long) * n) == cudaSuccess ? 0 : 2;
}
namespace tmr {

bool umma_available() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    ok = (major == 10) ? 1 : 0;
  }
  return ok == 1;
}

int umma_linear(const LinearArgs& g, cudaStream_t st) {
  TMR_CHECK_ARG(g.K % umma::BK == 0 && g.K > 0, "f16 linear: K=%d must be a multiple of %d", g.K, umma::BK);
  TMR_CHECK_ARG(g.N % 4 == 0, "f16 linear: N=%d must be a multiple of 4", g.N);
  TMR_CHECK_ARG(g.a16 && g.w16, "f16 linear: fp16 operands missing");
  TMR_CHECK_ARG(g.lda % 8 == 0 && g.ldw % 8 == 0 && g.ldo % 4 == 0, "f16 linear: leading dims must be multiples of 8 (A, W) / 4 (out)");
  TMR_CHECK_ARG(g.M < (int64_t)INT32_MAX, "f16 linear: M too large");
  if (g.M == 0) return TMR_OK;
  // at most one 128-row tile (a training batch, a module-level call on a few clips): N / 16 CTAs instead of one or two
  // busy CTA pairs of the persistent engine
  if (g.M <= 128 && !g.row2clip && env_int("TMR_GEMM_SMALL", 1)) return umma_linear_small(g, st);
  umma::GemmParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K; p.k_split = g.K;
  p.bias = g.bias; p.residual = g.residual; p.ldr = g.ldr; p.out = g.out; p.ldo = g.ldo; p.relu = g.relu;
  p.row2clip = g.row2clip; p.c0 = g.c0; p.h0_16 = g.h0_16;
  TMR_CHECK_ARG(!g.row2clip || (g.c0 && g.h0_16 && g.N == 4 * kD), "f16 linear: fused LSTM step 0 needs c0 / h0 and N = 4 x 512");
  return umma::launch_gemm<umma::EPI_LINEAR>(g.a16, g.lda, nullptr, 0, 0, g.w16, g.ldw, p, st);
}

int umma_lstm_step(const half_t* whh16, const float* xp, const int64_t* starts, int seq, int t, const half_t* h_prev,
                   half_t* h_out16, float* h_out, float* c, int B, cudaStream_t st, const float* xp_base,
                   int64_t xp_rows, int64_t xp_row0) {
  if (B == 0) return TMR_OK;
  // batches of at least one 256-clip tile: the weights-stationary kernel (umma_lstm_ws.cu); TMR_LSTM_WS=0 keeps
  // the streamed GEMM engine below, which also serves small batches
  static const int ws = env_int("TMR_LSTM_WS", 1);
  if (ws && B >= 256)
    return umma_lstm_step_ws(whh16, xp, starts, seq, t, h_prev, h_out16, h_out, c, B, st, xp_base, xp_rows, xp_row0);
  umma::GemmParams p{};
  p.xp_base = xp_base; p.xp_rows = xp_rows; p.x_row0 = xp_row0;
  p.M = B; p.N = 4 * kD; p.K = kD; p.k_split = kD;
  p.xp = xp; p.starts = starts; p.seq = seq; p.t = t; p.h_out = h_out; p.h_out16 = h_out16; p.c = c;
  return umma::launch_gemm<umma::EPI_LSTM>(h_prev, kD, nullptr, 0, 0, whh16, kD, p, st);
}

}  // namespace tmr
