// Bank-level TimeConv (TMR_MATH_TF32): the multi-scale temporal convolutions computed ONCE PER BANK
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
// tcgen05.mma.kind::tf32, and its epilogue assembles, per row, the SEVEN variants a window can ask
// of that row:
//     v0: interior      v1..v3: slot k = 0,1,2 (left-clipped)      v4..v6: slot k = L-1,L-2,L-3
// each = max(bank[rho], pool, conv3, conv5, conv7) with pool = bank[rho+1] (slot k-1) or 0 for v1
// (F.pad + MaxPool1d(2,1), NLB:67-68).  Output PB[row][7][512]; attention_pb_kernel consumes it.
// 236 MFLOP/clip become 7.9 MFLOP/row; only summation order changes.
//
// The SM's ingest from L2 (about 55 B/cycle) bounds tcgen05 kernels with fp32 operands, so the seven time
// shifts are NOT taken as seven shifted activation loads: the MMAs multiply the tile's rows unshifted
// and the epilogue applies the shift by exchanging accumulator rows through shared memory (3-row halo,
// 122 of 128 rows emitted per tile).  (Taking the shifts as row-offset descriptor views of one smem
// tile — matrix base offset — was tried first and produced wrong products on B200.)
#include "tmr_internal.h"
#include "umma_common.cuh"

namespace tmr {
namespace umma {

constexpr int BC_BM = 128;                 // bank rows whose tap products one tile computes
constexpr int BC_OUT = BC_BM - 6;          // rows it emits: the 3-row halo on either side feeds the shifts
constexpr int BC_NCH = 16;                 // output channels per tile: 15 taps x 16 = 240 TMEM columns, double-buffered
constexpr int BC_BK = 32;
constexpr int BC_STAGES = 3;
constexpr int BC_A_BYTES = BC_BM * BC_BK * 4;                  // 16 KB: the tile's rows for one channel chunk
constexpr int BC_W7_BYTES = 7 * BC_NCH * BC_BK * 4;            // 14 KB: rows ordered [channel][tap]
constexpr int BC_W5_BYTES = 5 * BC_NCH * BC_BK * 4;            // 10 KB
constexpr int BC_W3_BYTES = 3 * BC_NCH * BC_BK * 4;            //  6 KB
constexpr int BC_W_BYTES = BC_W7_BYTES + BC_W5_BYTES + BC_W3_BYTES;   // 30 KB = 240 rows: ONE MMA of N = 240 per k-step
constexpr int BC_STAGE_BYTES = BC_A_BYTES + BC_W_BYTES;        // 46 KB
constexpr int BC_EX_BYTES = 15 * BC_BM * 8 * 4;        // epilogue exchange: 15 taps x 128 rows x 8 channels
constexpr int BC_SMEM_BYTES = BC_STAGES * BC_STAGE_BYTES + BC_EX_BYTES + 1024 + 512;
constexpr int BC_THREADS = 192;
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
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps

// TMEM column of tap t of conv K for channel ch of the tile: the weight rows are staged [channel][tap]
// per conv (conv7 | conv5 | conv3), so a channel's taps are consecutive columns.
__host__ __device__ constexpr int col7(int ch) { return ch * 7; }
__host__ __device__ constexpr int col5(int ch) { return 7 * BC_NCH + ch * 5; }
__host__ __device__ constexpr int col3(int ch) { return 12 * BC_NCH + ch * 3; }

struct BankConvParams {
  const float* bank; float* pb; const float* bias3; const float* bias5; const float* bias7;
  int64_t n_rows; int64_t row_base; int64_t pb_rows; int64_t r_lo;   // bank_r holds rows r_lo .. (TMA row = row - r_lo)
  int64_t num_tiles;
};

// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 =
// epilogue.  The MMAs compute UNSHIFTED products Q_{K,t}[r] = W_K[:,:,t+h] . bank[r] for the tile's 128
// rows: one activation tile per channel chunk feeds all 15 taps, whose weight rows (rank-3 TMA boxes
// over [in-channel][tap][out-channel]) are stacked into ONE 240-row B operand — a single N = 240
// tcgen05.mma per k-step instead of seven narrow ones (a narrow MMA costs ~100 cycles whatever its N).
// The time shift P_{K,t}[rho] = Q_{K,t}[rho - t] is applied in the epilogue by exchanging rows through
// shared memory (hence the 3-row halo).  Accumulators are double-buffered in TMEM so the epilogue of
// tile i overlaps the main loop of tile i+1.
__global__ void __launch_bounds__(BC_THREADS, 1)
umma_bankconv_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w3,
                     const __grid_constant__ CUtensorMap tma_w5, const __grid_constant__ CUtensorMap tma_w7,
                     const BankConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the __shared__ array (no integer round trip) keeps the shared address space, so the
  // epilogue staging compiles to STS/LDS instead of generic ST.E/LD.E
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* ex = reinterpret_cast<float*>(smem + BC_STAGES * BC_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ex) + BC_EX_BYTES);
  uint64_t* full_bar = bars;                        // [BC_STAGES]
  uint64_t* empty_bar = bars + BC_STAGES;           // [BC_STAGES]
  uint64_t* acc_full = bars + 2 * BC_STAGES;        // [2]
  uint64_t* acc_empty = acc_full + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int N_TILES = kD / BC_NCH;                                 // 32 channel tiles, fastest index

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_x); tma_prefetch_desc(&tma_w3); tma_prefetch_desc(&tma_w5); tma_prefetch_desc(&tma_w7);
    for (int s = 0; s < BC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n0 = (int)(tile % N_TILES) * BC_NCH;
        const int64_t q0 = p.row_base + (tile / N_TILES) * BC_OUT - 3;           // first row the tile multiplies
        for (int chunk = 0; chunk < kD / BC_BK; ++chunk) {
          const int c0 = chunk * BC_BK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * BC_STAGE_BYTES;
          uint8_t* sw = sa + BC_A_BYTES;
          mbar_expect_tx(&full_bar[stage], BC_STAGE_BYTES);
          tma_load_2d(sa, &tma_x, &full_bar[stage], c0, (int)(q0 - p.r_lo));                 // OOB rows -> 0
          tma_load_3d(sw, &tma_w7, &full_bar[stage], c0, 0, n0);                             // [16 ch][7 taps] rows
          tma_load_3d(sw + BC_W7_BYTES, &tma_w5, &full_bar[stage], c0, 0, n0);               // [16 ch][5 taps]
          tma_load_3d(sw + BC_W7_BYTES + BC_W5_BYTES, &tma_w3, &full_bar[stage], c0, 0, n0); // [16 ch][3 taps]
          if (++stage == BC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BC_BM, BC_N);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);                         // epilogue drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        for (int chunk = 0; chunk < kD / BC_BK; ++chunk) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * BC_STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + BC_A_BYTES);             // 240 stacked weight rows
#pragma unroll
          for (int k = 0; k < BC_BK / 8; ++k)
            mma_tf32(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (chunk | k) != 0);
          mma_commit(&empty_bar[stage]);
          if (++stage == BC_STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit(&acc_full[acc]);
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;                                    // row inside the tile = TMEM lane
    // exchange layout: ex[tap][row][8]; taps 0..6 = conv7 t=-3..3, 7..11 = conv5 t=-2..2, 12..14 = conv3 t=-1..1
    auto exq = [&](int tap, int row) -> float* { return ex + ((size_t)tap * BC_BM + row) * 8; };
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int n0 = (int)(tile % N_TILES) * BC_NCH;
      const int64_t q0 = p.row_base + (tile / N_TILES) * BC_OUT - 3;
      const int64_t rho = q0 + r;                                   // bank row of this thread
      const int64_t prow = rho - p.row_base;
      const bool valid = r >= 3 && r < 3 + BC_OUT && prow >= 0 && prow < p.pb_rows && rho < p.n_rows;
      const bool has_next = valid && (rho + 1 < p.n_rows);
      // exact bank values of this row and of the next one (identity / pool branches): requested before the
      // accumulator is ready so their latency hides behind the main loop
      float4 x0v[BC_NCH / 4], x1v[BC_NCH / 4];
#pragma unroll
      for (int h = 0; h < BC_NCH / 4; ++h) {
        x0v[h] = valid ? __ldg(reinterpret_cast<const float4*>(p.bank + rho * kD + n0) + h) : make_float4(0.f, 0.f, 0.f, 0.f);
        x1v[h] = has_next ? __ldg(reinterpret_cast<const float4*>(p.bank + (rho + 1) * kD + n0) + h) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float* dst = p.pb + (valid ? prow : 0) * (7 * kD) + n0;
      mbar_wait(&acc_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
#pragma unroll
      for (int cc = 0; cc < BC_NCH; cc += 8) {
        {
          // a channel's taps are consecutive TMEM columns: three 8-column loads per channel (7 / 5 / 3 taps used)
          float v[8];
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            tmem_ld8(t_row + col7(cc + ch), v);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 7; ++t) exq(t, r)[ch] = v[t];
            tmem_ld8(t_row + col5(cc + ch), v);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 5; ++t) exq(7 + t, r)[ch] = v[t];
            tmem_ld8(t_row + col3(cc + ch), v);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 3; ++t) exq(12 + t, r)[ch] = v[t];
          }
        }
        if (cc + 8 >= BC_NCH) {                         // last TMEM read of this tile: hand the buffer back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
        epi_barrier();
        if (valid) {
          // P_{K,t}[rho] = Q_{K,t}[rho - t]: row r - t of the exchange buffer
          float P7[7][8], P5[5][8], P3[3][8];
#pragma unroll
          for (int t = -3; t <= 3; ++t) {
            const float4 a = *reinterpret_cast<const float4*>(exq(t + 3, r - t));
            const float4 b = *reinterpret_cast<const float4*>(exq(t + 3, r - t) + 4);
            P7[t + 3][0] = a.x; P7[t + 3][1] = a.y; P7[t + 3][2] = a.z; P7[t + 3][3] = a.w;
            P7[t + 3][4] = b.x; P7[t + 3][5] = b.y; P7[t + 3][6] = b.z; P7[t + 3][7] = b.w;
            if (t >= -2 && t <= 2) {
              const float4 c = *reinterpret_cast<const float4*>(exq(7 + t + 2, r - t));
              const float4 d = *reinterpret_cast<const float4*>(exq(7 + t + 2, r - t) + 4);
              P5[t + 2][0] = c.x; P5[t + 2][1] = c.y; P5[t + 2][2] = c.z; P5[t + 2][3] = c.w;
              P5[t + 2][4] = d.x; P5[t + 2][5] = d.y; P5[t + 2][6] = d.z; P5[t + 2][7] = d.w;
            }
            if (t >= -1 && t <= 1) {
              const float4 c = *reinterpret_cast<const float4*>(exq(12 + t + 1, r - t));
              const float4 d = *reinterpret_cast<const float4*>(exq(12 + t + 1, r - t) + 4);
              P3[t + 1][0] = c.x; P3[t + 1][1] = c.y; P3[t + 1][2] = c.z; P3[t + 1][3] = c.w;
              P3[t + 1][4] = d.x; P3[t + 1][5] = d.y; P3[t + 1][6] = d.z; P3[t + 1][7] = d.w;
            }
          }
          float b3[8], b5[8], b7[8], x0[8], x1[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p.bias3 + n0 + cc) + h);
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias5 + n0 + cc) + h);
            const float4 c = __ldg(reinterpret_cast<const float4*>(p.bias7 + n0 + cc) + h);
            const float4 d = x0v[cc / 4 + h];
            const float4 e = x1v[cc / 4 + h];
            b3[4 * h] = a.x; b3[4 * h + 1] = a.y; b3[4 * h + 2] = a.z; b3[4 * h + 3] = a.w;
            b5[4 * h] = b.x; b5[4 * h + 1] = b.y; b5[4 * h + 2] = b.z; b5[4 * h + 3] = b.w;
            b7[4 * h] = c.x; b7[4 * h + 1] = c.y; b7[4 * h + 2] = c.z; b7[4 * h + 3] = c.w;
            x0[4 * h] = d.x; x0[4 * h + 1] = d.y; x0[4 * h + 2] = d.z; x0[4 * h + 3] = d.w;
            x1[4 * h] = e.x; x1[4 * h + 1] = e.y; x1[4 * h + 2] = e.z; x1[4 * h + 3] = e.w;
          }
          float out[7][8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
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
#pragma unroll
          for (int v = 0; v < 7; ++v) {
            float4* d4 = reinterpret_cast<float4*>(dst + v * kD + cc);
            d4[0] = make_float4(out[v][0], out[v][1], out[v][2], out[v][3]);
            d4[1] = make_float4(out[v][4], out[v][5], out[v][6], out[v][7]);
          }
        }
        epi_barrier();                                   // exchange buffer is reused by the next 8 channels / tile
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, BC_TMEM_COLS); }
}

}  // namespace umma

// pb[(row - row_base)*7 + v][512] for bank rows row_base .. row_base + pb_rows - 1.
// bank = exact values (identity / pool branches); bank_r = rows r_lo .. r_lo + r_cnt - 1 of the bank
// rounded to TF32 (MMA operand) — must cover row_base - 3 .. row_base + pb_rows + 2 where they exist.
int umma_bankconv(const float* packed, const float* bank, const float* bank_r, int64_t n_rows, int64_t r_lo,
                  int64_t r_cnt, int64_t row_base, int64_t pb_rows, float* pb, cudaStream_t st) {
  using namespace umma;
  if (pb_rows <= 0) return TMR_OK;
  TMR_CHECK_ARG(n_rows < (int64_t)INT32_MAX - 256, "bankconv: bank too large");
  BankConvParams p{};
  p.bank = bank; p.pb = pb;
  p.bias3 = packed + TimeConvPacked::b3_off; p.bias5 = packed + TimeConvPacked::b5_off; p.bias7 = packed + TimeConvPacked::b7_off;
  p.n_rows = n_rows; p.row_base = row_base; p.pb_rows = pb_rows; p.r_lo = r_lo;
  p.num_tiles = ((pb_rows + BC_OUT - 1) / BC_OUT) * (kD / BC_NCH);
  CUtensorMap tx, tw3, tw5, tw7;
  {
    uint64_t dims[2] = {(uint64_t)kD, (uint64_t)r_cnt};
    uint64_t str[1] = {(uint64_t)kD * 4};
    uint32_t box[2] = {BC_BK, BC_BM};
    TMR_TRY(make_tmap(&tx, bank_r, 2, dims, str, box));
    const float* pr = packed + TimeConvPacked::fp32_total;
    const float* w[3] = {pr + TimeConvPacked::w3_off, pr + TimeConvPacked::w5_off, pr + TimeConvPacked::w7_off};
    CUtensorMap* tw[3] = {&tw3, &tw5, &tw7};
    for (int i = 0; i < 3; ++i) {      // packed Wp_K[o][tap][c] viewed as (c, tap, o): a box = 16 channels x K taps x 32 c
      const int taps = 3 + 2 * i;
      uint64_t dw[3] = {(uint64_t)kD, (uint64_t)taps, (uint64_t)kD};
      uint64_t sw[2] = {(uint64_t)kD * 4, (uint64_t)taps * kD * 4};
      uint32_t bw[3] = {BC_BK, (uint32_t)taps, BC_NCH};
      TMR_TRY(make_tmap(tw[i], w[i], 3, dw, sw, bw));
    }
  }
  TMR_CUDA(cudaFuncSetAttribute(umma_bankconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM_BYTES));
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t grid = p.num_tiles < sms ? p.num_tiles : sms;
  umma_bankconv_kernel<<<(unsigned)grid, BC_THREADS, BC_SMEM_BYTES, st>>>(tx, tw3, tw5, tw7, p);
  TMR_LAUNCH_CHECK("umma_bankconv_kernel");
  return TMR_OK;
}

}  // namespace tmr
