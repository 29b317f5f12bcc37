// Bank-level TimeConv (TMR_MATH_F16): the multi-scale temporal convolutions computed ONCE PER BANK
// ROW instead of once per (clip, slot).
//
// For a clip whose window is a contiguous run of bank rows (every clip at least L clips into its
// video — 98.5 % of a Cholec80-shaped bank) slot k is bank row rho = r0 - k and
//     conv_K[k] = b_K + sum_t W_K[:,:,t+h] . bank[rho - t],   t in [-h,h], 0 <= k+t <= L-1,
// so the per-tap products P_{K,t}[rho] = W_K[:,:,t+h] . bank[rho - t] depend only on the row.  Away
// from the window edges (3 <= k <= L-4) all taps are present and the TimeConv output is a pure
// function of rho; at the three slots next to either edge the taps that fall outside the window
// are dropped (zero "same" padding, NLB:55-65).  This kernel accumulates the 15 tap products of a
// 128-row x 16-channel tile in TMEM (7 shift groups: 16/32/48/48/48/32/16 columns = 240, two buffers) with
// tcgen05.mma.kind::f16 (fp16 operands, fp32 accumulate), and its epilogue assembles, per row, the SEVEN variants a window can ask
// of that row:
//     v0: interior      v1..v3: slot k = 0,1,2 (left-clipped)      v4..v6: slot k = L-1,L-2,L-3
// each = max(bank[rho], pool, conv3, conv5, conv7) with pool = bank[rho+1] (slot k-1) or 0 for v1
// (F.pad + MaxPool1d(2,1), NLB:67-68).  Output PB[row][7][512]; attention_pb_kernel consumes it.
// 236 MFLOP/clip become 7.9 MFLOP/row; only summation order changes.
//
// The SM's ingest from L2 (about 55 B/cycle) is dear, so the seven time
// shifts are NOT taken as seven shifted activation loads: the MMAs multiply the tile's rows unshifted
// and the epilogue applies the shift by exchanging accumulator rows through shared memory (3-row halo,
// 122 of 128 rows emitted per tile).  (Taking the shifts as row-offset descriptor views of one smem
// tile — matrix base offset — was tried first and produced wrong products on B200.)
#include <stdlib.h>
#include "tmr_internal.h"
#include "umma_common.cuh"

namespace tmr {
namespace umma {

constexpr int BC_BM = 128;                 // bank rows whose tap products one tile computes
constexpr int BC_OUT = BC_BM - 6;          // rows it emits: the 3-row halo on either side feeds the shifts
constexpr int BC_NCH = 16;                 // output channels per tile: 15 taps x 16 = 240 TMEM columns, double-buffered
constexpr int BC_BK = 64;                  // fp16 input channels per k-step = one 128-byte swizzle row
constexpr int BC_STAGES = 3;
constexpr int BC_A_BYTES = BC_BM * BC_BK * 2;                  // 16 KB: the tile's rows for one channel chunk
// 2-SM MMA (cta_group::2): a CTA PAIR multiplies 2 x 128 rows by the 240 stacked weight rows of the tile's 16
// channels; each CTA stages its own 128 bank rows and HALF of the weight rows - those of 8 channels:
constexpr int BC_HCH = BC_NCH / 2;                             // channels whose weight rows one CTA stages
constexpr int BC_W7_BYTES = 7 * BC_HCH * BC_BK * 2;            //  7 KB: rows ordered [channel][tap]
constexpr int BC_W5_BYTES = 5 * BC_HCH * BC_BK * 2;            //  5 KB
constexpr int BC_W3_BYTES = 3 * BC_HCH * BC_BK * 2;            //  3 KB
constexpr int BC_W_BYTES = BC_W7_BYTES + BC_W5_BYTES + BC_W3_BYTES;   // 15 KB = 120 rows per CTA: ONE MMA of N = 240 per k-step
constexpr int BC_STAGE_BYTES = BC_A_BYTES + BC_W_BYTES;        // 31 KB
constexpr int BC_EX_BYTES = 15 * BC_BM * 8 * 4;        // epilogue exchange: 15 taps x 128 rows x 8 channels
constexpr int BC_SMEM_BYTES = BC_STAGES * BC_STAGE_BYTES + 2 * BC_EX_BYTES + 1024 + 512;   // exchange buffer x 2
constexpr int BC_EPI_WARPS = 8;            // two per TMEM lane quarter
constexpr int BC_THREADS = 64 + 32 * BC_EPI_WARPS;
constexpr int BC_TMEM_COLS = 512;          // 2 accumulator buffers of 256 columns (240 used)
constexpr int BC_N = 15 * BC_NCH;          // 240

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps

// TMEM column of tap t of conv K for channel ch of the tile.  Columns 0..119 are the weight rows CTA 0 staged
// (channels 0..7: conv7 | conv5 | conv3, each [channel][tap]), columns 120..239 those of CTA 1 (channels 8..15),
// so a channel's taps are consecutive columns and an 8-channel chunk is three runs of 56 / 40 / 24 columns.
__host__ __device__ constexpr int col7(int ch) { return (ch / BC_HCH) * (15 * BC_HCH) + (ch % BC_HCH) * 7; }
__host__ __device__ constexpr int col5(int ch) { return (ch / BC_HCH) * (15 * BC_HCH) + 7 * BC_HCH + (ch % BC_HCH) * 5; }
__host__ __device__ constexpr int col3(int ch) { return (ch / BC_HCH) * (15 * BC_HCH) + 12 * BC_HCH + (ch % BC_HCH) * 3; }

struct BankConvParams {
  const float* bank; float* pb; const float* bias3; const float* bias5; const float* bias7;
  int64_t n_rows; int64_t row_base; int64_t pb_rows; int64_t r_lo;   // bank_r holds rows r_lo .. (TMA row = row - r_lo)
  int64_t num_tiles;
  float* q_out; int raw;   // raw: emit the UNSHIFTED tap products Q[row][15][512] instead of the 7 variants
  int ablate;      // timing experiment (WRONG results), env TMR_BC_ABL: 2 = no global stores
};

// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..9 =
// epilogue (two per TMEM lane quarter: one moves the conv7 taps to the exchange buffer, the other conv5 + conv3;
// then each assembles the variants of 16 of the quarter's 32 rows).  The MMAs compute UNSHIFTED products Q_{K,t}[r] = W_K[:,:,t+h] . bank[r] for the tile's 128
// rows: one activation tile per channel chunk feeds all 15 taps, whose weight rows (rank-3 TMA boxes
// over [in-channel][tap][out-channel]) are stacked into ONE 240-row B operand — a single N = 240
// tcgen05.mma per k-step instead of seven narrow ones (a narrow MMA costs ~100 cycles whatever its N).
// The time shift P_{K,t}[rho] = Q_{K,t}[rho - t] is applied in the epilogue by exchanging rows through
// shared memory (hence the 3-row halo).  Accumulators are double-buffered in TMEM so the epilogue of
// tile i overlaps the main loop of tile i+1.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BC_THREADS, 1)
umma_bankconv_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w3,
                     const __grid_constant__ CUtensorMap tma_w5, const __grid_constant__ CUtensorMap tma_w7,
                     const BankConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the __shared__ array (no integer round trip) keeps the shared address space, so the
  // epilogue staging compiles to STS/LDS instead of generic ST.E/LD.E
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* ex = reinterpret_cast<float*>(smem + BC_STAGES * BC_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ex) + 2 * BC_EX_BYTES);
  uint64_t* full_bar = bars;                        // [BC_STAGES]
  uint64_t* empty_bar = bars + BC_STAGES;           // [BC_STAGES]
  uint64_t* acc_full = bars + 2 * BC_STAGES;        // [2]
  uint64_t* acc_empty = acc_full + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int N_TILES = kD / BC_NCH;                                 // 32 channel tiles, fastest index
  const uint32_t crank = cluster_ctarank();
  const int64_t tile0 = blockIdx.x >> 1, tile_stride = gridDim.x >> 1;  // a tile = 2 x 122 rows x 16 channels per CTA pair
  constexpr uint16_t kMask = 3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_x); tma_prefetch_desc(&tma_w3); tma_prefetch_desc(&tma_w5); tma_prefetch_desc(&tma_w7);
    for (int s = 0; s < BC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 2 * BC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, BC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // peer barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t tile = tile0; tile < p.num_tiles; tile += tile_stride) {
        const int n0 = (int)(tile % N_TILES) * BC_NCH + (int)crank * BC_HCH;     // the 8 channels whose weights I stage
        const int64_t q0 = p.row_base + ((tile / N_TILES) * 2 + crank) * BC_OUT - 3;   // first row this CTA multiplies
        for (int chunk = 0; chunk < kD / BC_BK; ++chunk) {
          const int c0 = chunk * BC_BK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * BC_STAGE_BYTES;
          uint8_t* sw = sa + BC_A_BYTES;
          // both CTAs' bytes land on the LEADER's barrier, which only the leader arms
          if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * BC_STAGE_BYTES);
          tma_load_2d_2sm(sa, &tma_x, &full_bar[stage], c0, (int)(q0 - p.r_lo));                 // OOB rows -> 0
          tma_load_3d_2sm(sw, &tma_w7, &full_bar[stage], c0, 0, n0);                             // [8 ch][7 taps] rows
          tma_load_3d_2sm(sw + BC_W7_BYTES, &tma_w5, &full_bar[stage], c0, 0, n0);               // [8 ch][5 taps]
          tma_load_3d_2sm(sw + BC_W7_BYTES + BC_W5_BYTES, &tma_w3, &full_bar[stage], c0, 0, n0); // [8 ch][3 taps]
          if (++stage == BC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {                                               // the leader CTA issues for the pair
      constexpr uint32_t idesc = make_idesc_f16(2 * BC_BM, BC_N);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int64_t tile = tile0; tile < p.num_tiles; tile += tile_stride, ++it) {
        const int acc = it & 1;
        mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);                         // epilogue drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        for (int chunk = 0; chunk < kD / BC_BK; ++chunk) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * BC_STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + BC_A_BYTES);             // my 120 of the 240 stacked weight rows
#pragma unroll
          for (int k = 0; k < BC_BK / 16; ++k)
            mma_f16_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (chunk | k) != 0);
          mma_commit_2sm_mcast(&empty_bar[stage], kMask);
          if (++stage == BC_STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit_2sm_mcast(&acc_full[acc], kMask);
      }
    }
  } else {
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;                               // which of the quarter's two warps
    const int r = q * 32 + lane;                                    // phase 1: row inside the tile = TMEM lane
    // exchange layout: ex[buffer][tap][row][8 channels]; taps 0..6 = conv7 t=-3..3, 7..11 = conv5 t=-2..2,
    // 12..14 = conv3 t=-1..1.  The two 16-byte halves of a row swap places in rows with bit 2 set, so 128-bit
    // accesses by eight consecutive rows (phase 1) or by eight (row, half) pairs (phase 2) never share a bank group.
    auto exh = [&](int buf, int tap, int row, int h) -> float4* {
      return reinterpret_cast<float4*>(ex + (size_t)buf * (BC_EX_BYTES / 4) + ((size_t)tap * BC_BM + row) * 8 + ((h ^ ((row >> 2) & 1)) << 2));
    };
    // phase 2: a lane owns FOUR channels (half = lane & 1 of the 8-channel chunk) of row 32q + 16 part + lane/2,
    // so every global load / store instruction covers whole 32-byte sectors (16 rows x 32 B) instead of 32
    // half-filled ones (row per thread).
    const int half = lane & 1;
    const int r2 = q * 32 + 16 * part + (lane >> 1);
    struct TileInfo { int n0; int acc; int64_t prow; bool valid; float4 x0v[2], x1v[2]; };
    // tile bookkeeping + the exact bank values of the lane's row and of the next one (identity / pool branches),
    // requested before the accumulator is awaited so their latency hides behind the main loop
    auto load_tile = [&](int64_t tile, int it, TileInfo& T) {
      T.acc = it & 1;
      T.n0 = (int)(tile % N_TILES) * BC_NCH;
      const int64_t q0 = p.row_base + ((tile / N_TILES) * 2 + crank) * BC_OUT - 3;
      const int64_t rho = q0 + r2;                                  // bank row
      T.prow = rho - p.row_base;
      T.valid = r2 >= 3 && r2 < 3 + BC_OUT && T.prow >= 0 && T.prow < p.pb_rows && rho < p.n_rows;
      const bool has_next = T.valid && (rho + 1 < p.n_rows);
#pragma unroll
      for (int c8 = 0; c8 < 2; ++c8) {
        const float* src = p.bank + rho * kD + T.n0 + 8 * c8 + 4 * half;
        T.x0v[c8] = (T.valid && !p.raw) ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
        T.x1v[c8] = (has_next && !p.raw) ? __ldg(reinterpret_cast<const float4*>(src + kD)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(&acc_full[T.acc], (it >> 1) & 1);
      tc_fence_after();
    };
    // phase 1 of an 8-channel chunk: accumulator -> exchange buffer.  A channel's taps are consecutive TMEM
    // columns, so the chunk's 8 channels x K taps are ONE run of 56 / 40 / 24 columns: a few wide tcgen05.ld,
    // and the row goes out as two 128-bit stores per tap (8 channels).  The quarter's two warps split the taps.
    auto phase1 = [&](const TileInfo& T, int cc, int buf) {
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(T.acc * 256);
      uint32_t v[32], w[32];
      if (part == 0) {
        tmem_ld32(t_row + col7(cc), v);
        tmem_ld32(t_row + col7(cc) + 32, w);             // 56 used; the rest belongs to conv5
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 7; ++t)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t e[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { const int idx = (4 * h + c) * 7 + t; e[c] = idx < 32 ? v[idx] : w[idx - 32]; }
            *reinterpret_cast<uint4*>(exh(buf, t, r, h)) = make_uint4(e[0], e[1], e[2], e[3]);
          }
      } else {
        float f8[8];
        tmem_ld32(t_row + col5(cc), v);
        tmem_ld8(t_row + col5(cc) + 32, f8);
        tmem_ld32(t_row + col3(cc), w);                  // 24 used (columns up to 247 of the 256-column buffer)
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 5; ++t)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t e[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { const int idx = (4 * h + c) * 5 + t; e[c] = idx < 32 ? v[idx] : __float_as_uint(f8[idx - 32]); }
            *reinterpret_cast<uint4*>(exh(buf, 7 + t, r, h)) = make_uint4(e[0], e[1], e[2], e[3]);
          }
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            *reinterpret_cast<uint4*>(exh(buf, 12 + t, r, h)) =
                make_uint4(w[(4 * h) * 3 + t], w[(4 * h + 1) * 3 + t], w[(4 * h + 2) * 3 + t], w[(4 * h + 3) * 3 + t]);
      }
      if (cc + 8 >= BC_NCH) {                            // last TMEM read of this tile: hand the buffer back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&acc_empty[T.acc], 0);
      }
    };
    // phase 2: exchange buffer -> the lane's row: time shifts, the 7 edge variants, stores
    auto phase2 = [&](const TileInfo& T, int cc, int buf) {
      if (!T.valid) return;
      if (p.raw) {            // Q_{K,t}[rho] = W_K[:,:,t+h] . x[rho] as computed: any window can be assembled from these
        float* dst = p.q_out + T.prow * (15 * kD) + T.n0 + cc + 4 * half;
#pragma unroll
        for (int tap = 0; tap < 15; ++tap) *reinterpret_cast<float4*>(dst + tap * kD) = *exh(buf, tap, r2, half);
        return;
      }
      const float4 bb3 = __ldg(reinterpret_cast<const float4*>(p.bias3 + T.n0 + cc + 4 * half));
      const float4 bb5 = __ldg(reinterpret_cast<const float4*>(p.bias5 + T.n0 + cc + 4 * half));
      const float4 bb7 = __ldg(reinterpret_cast<const float4*>(p.bias7 + T.n0 + cc + 4 * half));
      // P_{K,t}[rho] = Q_{K,t}[rho - t]: row r2 - t of the exchange buffer
      float P7[7][4], P5[5][4], P3[3][4];
#pragma unroll
      for (int t = -3; t <= 3; ++t) {
        const float4 a = *exh(buf, t + 3, r2 - t, half);
        P7[t + 3][0] = a.x; P7[t + 3][1] = a.y; P7[t + 3][2] = a.z; P7[t + 3][3] = a.w;
        if (t >= -2 && t <= 2) {
          const float4 c = *exh(buf, 7 + t + 2, r2 - t, half);
          P5[t + 2][0] = c.x; P5[t + 2][1] = c.y; P5[t + 2][2] = c.z; P5[t + 2][3] = c.w;
        }
        if (t >= -1 && t <= 1) {
          const float4 c = *exh(buf, 12 + t + 1, r2 - t, half);
          P3[t + 1][0] = c.x; P3[t + 1][1] = c.y; P3[t + 1][2] = c.z; P3[t + 1][3] = c.w;
        }
      }
      const float b3[4] = {bb3.x, bb3.y, bb3.z, bb3.w}, b5[4] = {bb5.x, bb5.y, bb5.z, bb5.w}, b7[4] = {bb7.x, bb7.y, bb7.z, bb7.w};
      const float4 xa = T.x0v[cc / 8], xb = T.x1v[cc / 8];
      const float x0[4] = {xa.x, xa.y, xa.z, xa.w}, x1[4] = {xb.x, xb.y, xb.z, xb.w};
      float out[7][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // R_b = bias + sum_{t=0..b} P_t ; Lf_a = sum_{t=-a..-1} P_t ; conv(a,b) = Lf_a + R_b
        const float r7_0 = b7[j] + P7[3][j], r7_1 = r7_0 + P7[4][j], r7_2 = r7_1 + P7[5][j], r7_3 = r7_2 + P7[6][j];
        const float l7_1 = P7[2][j], l7_2 = l7_1 + P7[1][j], l7_3 = l7_2 + P7[0][j];
        const float r5_0 = b5[j] + P5[2][j], r5_1 = r5_0 + P5[3][j], r5_2 = r5_1 + P5[4][j];
        const float l5_1 = P5[1][j], l5_2 = l5_1 + P5[0][j];
        const float r3_0 = b3[j] + P3[1][j], r3_1 = r3_0 + P3[2][j];
        const float l3_1 = P3[0][j];
        const float idp = fmaxf(x0[j], x1[j]);          // identity + pool branches, slots k >= 1
        const float full3 = l3_1 + r3_1, full5 = l5_2 + r5_2;
        out[0][j] = fmaxf(fmaxf(fmaxf(l7_3 + r7_3, full5), full3), idp);
        out[1][j] = fmaxf(fmaxf(fmaxf(r7_3, r5_2), r3_1), fmaxf(x0[j], 0.f));          // k = 0: pool sees the zero pad
        out[2][j] = fmaxf(fmaxf(fmaxf(l7_1 + r7_3, l5_1 + r5_2), full3), idp);         // k = 1
        out[3][j] = fmaxf(fmaxf(fmaxf(l7_2 + r7_3, full5), full3), idp);               // k = 2
        out[4][j] = fmaxf(fmaxf(fmaxf(l7_3 + r7_0, l5_2 + r5_0), l3_1 + r3_0), idp);   // k = L-1
        out[5][j] = fmaxf(fmaxf(fmaxf(l7_3 + r7_1, l5_2 + r5_1), full3), idp);         // k = L-2
        out[6][j] = fmaxf(fmaxf(fmaxf(l7_3 + r7_2, full5), full3), idp);               // k = L-3
      }
      if (!(p.ablate & 2) || out[0][0] == 123.456f) {
        float* dst = p.pb + T.prow * (7 * kD) + T.n0 + cc + 4 * half;
#pragma unroll
        for (int v = 0; v < 7; ++v)
          *reinterpret_cast<float4*>(dst + v * kD) = make_float4(out[v][0], out[v][1], out[v][2], out[v][3]);
      }
    };
    // Software pipeline over the 8-channel chunks of this CTA's tiles with a DOUBLE-BUFFERED exchange buffer:
    // between two barriers every warp runs phase 1 of chunk j+1 (TMEM latency) and phase 2 of chunk j (shared
    // memory, math, stores), so the warps of an SM sub-partition overlap the one with the other and a tile
    // costs two barriers instead of four.  Buffer (j+1)&1 was last read in phase 2 of chunk j-1, which every
    // warp finished before the barrier in between.
    int64_t tile = tile0;
    if (tile < p.num_tiles) {
      TileInfo cur, nxt;
      int it = 0;
      load_tile(tile, it, cur);
      phase1(cur, 0, 0);
      for (;;) {
        epi_barrier();
        phase1(cur, 8, 1);
        phase2(cur, 0, 0);
        epi_barrier();
        const int64_t ntile = tile + tile_stride;
        if (ntile < p.num_tiles) { load_tile(ntile, it + 1, nxt); phase1(nxt, 0, 0); }
        phase2(cur, 8, 1);
        if (ntile >= p.num_tiles) break;
        cur = nxt; tile = ntile; ++it;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer may still multicast into this CTA's barriers
  if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm(tmem_base, BC_TMEM_COLS); }
}

}  // namespace umma

static int launch_bankconv(const float* packed, const half_t* bank16, int64_t r_cnt, umma::BankConvParams p, cudaStream_t st);

// Unshifted tap products of `n` fp16 rows: q[(row*15 + tap)][512], taps 0..6 = conv7 t=-3..3,
// 7..11 = conv5 t=-2..2, 12..14 = conv3 t=-1..1.  The TimeConv of ANY window over these rows is
// conv_K[k] = b_K + sum_t q[row(slot k+t)][tap(K,t)] (irr_assemble_kernel).
int umma_bankconv_raw(const float* packed, const half_t* rows16, int64_t n, float* q, cudaStream_t st) {
  using namespace umma;
  if (n <= 0) return TMR_OK;
  BankConvParams p{};
  p.bank = nullptr; p.pb = nullptr; p.q_out = q; p.raw = 1;
  p.bias3 = packed + TimeConvPacked::b3_off; p.bias5 = packed + TimeConvPacked::b5_off; p.bias7 = packed + TimeConvPacked::b7_off;
  p.n_rows = n; p.row_base = 0; p.pb_rows = n; p.r_lo = 0;
  return launch_bankconv(packed, rows16, n, p, st);
}

// pb[(row - row_base)*7 + v][512] for bank rows row_base .. row_base + pb_rows - 1.
// bank = exact values (identity / pool branches); bank16 = rows r_lo .. r_lo + r_cnt - 1 of the bank
// in fp16 (MMA operand) — must cover row_base - 3 .. row_base + pb_rows + 2 where they exist.
int umma_bankconv(const float* packed, const float* bank, const half_t* bank16, int64_t n_rows, int64_t r_lo,
                  int64_t r_cnt, int64_t row_base, int64_t pb_rows, float* pb, cudaStream_t st) {
  using namespace umma;
  if (pb_rows <= 0) return TMR_OK;
  TMR_CHECK_ARG(n_rows < (int64_t)INT32_MAX - 256, "bankconv: bank too large");
  BankConvParams p{};
  p.bank = bank; p.pb = pb;
  p.bias3 = packed + TimeConvPacked::b3_off; p.bias5 = packed + TimeConvPacked::b5_off; p.bias7 = packed + TimeConvPacked::b7_off;
  p.n_rows = n_rows; p.row_base = row_base; p.pb_rows = pb_rows; p.r_lo = r_lo;
  return launch_bankconv(packed, bank16, r_cnt, p, st);
}

static int launch_bankconv(const float* packed, const half_t* bank16, int64_t r_cnt, umma::BankConvParams p, cudaStream_t st) {
  using namespace umma;
  p.num_tiles = ((p.pb_rows + 2 * BC_OUT - 1) / (2 * BC_OUT)) * (kD / BC_NCH);      // a CTA pair emits 2 x 122 rows
  static const int abl = env_int("TMR_BC_ABL", 0);
  p.ablate = abl;
  CUtensorMap tx, tw3, tw5, tw7;
  {
    uint64_t dims[2] = {(uint64_t)kD, (uint64_t)r_cnt};
    uint64_t str[1] = {(uint64_t)kD * 2};
    uint32_t box[2] = {BC_BK, BC_BM};
    TMR_TRY(make_tmap(&tx, bank16, 2, dims, str, box, 2));
    const half_t* pr = mirror16<TimeConvPacked>(packed);
    const half_t* w[3] = {pr + TimeConvPacked::w3_off, pr + TimeConvPacked::w5_off, pr + TimeConvPacked::w7_off};
    CUtensorMap* tw[3] = {&tw3, &tw5, &tw7};
    for (int i = 0; i < 3; ++i) {      // packed Wp_K[o][tap][c] viewed as (c, tap, o): a box = 8 channels x K taps x 64 c
      const int taps = 3 + 2 * i;
      uint64_t dw[3] = {(uint64_t)kD, (uint64_t)taps, (uint64_t)kD};
      uint64_t sw[2] = {(uint64_t)kD * 2, (uint64_t)taps * kD * 2};
      uint32_t bw[3] = {BC_BK, (uint32_t)taps, BC_HCH};
      TMR_TRY(make_tmap(tw[i], w[i], 3, dw, sw, bw, 2));
    }
  }
  TMR_CUDA(cudaFuncSetAttribute(umma_bankconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM_BYTES));
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t pairs = p.num_tiles < sms / 2 ? p.num_tiles : sms / 2;
  umma_bankconv_kernel<<<(unsigned)(2 * pairs), BC_THREADS, BC_SMEM_BYTES, st>>>(tx, tw3, tw5, tw7, p);
  TMR_LAUNCH_CHECK("umma_bankconv_kernel");
  return TMR_OK;
}

}  // namespace tmr
