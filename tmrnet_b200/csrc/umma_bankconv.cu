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
// 128-row x 32-channel tile in TMEM (7 shift groups: 32/64/96/96/96/64/32 columns = 480) with
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
constexpr int BC_NCH = 32;                 // output channels per tile
constexpr int BC_BK = 32;
constexpr int BC_A_STAGES = 3;
constexpr int BC_A_BYTES = BC_BM * BC_BK * 4;          // 16 KB: the tile's rows, loaded ONCE per channel chunk
constexpr int BC_W_BYTES = BC_NCH * BC_BK * 4;         // 4 KB per tap tile
constexpr int BC_W_STAGE_BYTES = 3 * BC_W_BYTES;       // 12 KB: the taps of one shift (conv7 | conv5 | conv3)
constexpr int BC_W_STAGES = 9;
constexpr int BC_EX_BYTES = 15 * BC_BM * 8 * 4;        // epilogue exchange: 15 taps x 128 rows x 8 channels
constexpr int BC_SMEM_BYTES = BC_A_STAGES * BC_A_BYTES + BC_W_STAGES * BC_W_STAGE_BYTES + BC_EX_BYTES + 1024 + 512;
constexpr int BC_THREADS = 192;
constexpr int BC_TMEM_COLS = 512;

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps

// TMEM column of shift group t (t = -3..3); inside a group: conv7 | conv5 | conv3 (32 columns each)
__host__ __device__ constexpr int group_col(int t) {
  return t == -3 ? 0 : t == -2 ? 32 : t == -1 ? 96 : t == 0 ? 192 : t == 1 ? 288 : t == 2 ? 384 : 448;
}

struct BankConvParams {
  const float* bank; float* pb; const float* bias3; const float* bias5; const float* bias7;
  int64_t n_rows; int64_t row_base; int64_t pb_rows; int64_t r_lo;   // bank_r holds rows r_lo .. (TMA row = row - r_lo)
};

// The MMAs compute UNSHIFTED products Q_{K,t}[r] = W_K[:,:,t+h] . bank[r] for the tile's 128 rows, so one
// activation tile per channel chunk feeds all 15 taps; the time shift P_{K,t}[rho] = Q_{K,t}[rho - t] is
// applied in the epilogue by exchanging rows through shared memory (hence the 3-row halo).
__global__ void __launch_bounds__(BC_THREADS, 1)
umma_bankconv_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w3,
                     const __grid_constant__ CUtensorMap tma_w5, const __grid_constant__ CUtensorMap tma_w7,
                     const BankConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem + BC_A_STAGES * BC_A_BYTES;
  float* ex = reinterpret_cast<float*>(smem_w + BC_W_STAGES * BC_W_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ex) + BC_EX_BYTES);
  uint64_t* full_bar = bars;                        // [BC_W_STAGES]
  uint64_t* empty_bar = bars + BC_W_STAGES;         // [BC_W_STAGES]
  uint64_t* a_full = bars + 2 * BC_W_STAGES;        // [BC_A_STAGES]
  uint64_t* a_empty = a_full + BC_A_STAGES;         // [BC_A_STAGES]
  uint64_t* acc_full = a_empty + BC_A_STAGES;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int N_TILES = kD / BC_NCH;                                 // 16
  const int n0 = (blockIdx.x % N_TILES) * BC_NCH;
  const int64_t out0 = p.row_base + (int64_t)(blockIdx.x / N_TILES) * BC_OUT;   // first row this tile emits
  const int64_t q0 = out0 - 3;                                                   // first row it multiplies

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_x); tma_prefetch_desc(&tma_w3); tma_prefetch_desc(&tma_w5); tma_prefetch_desc(&tma_w7);
    for (int s = 0; s < BC_W_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < BC_A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(acc_full, 1);
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
      int a_stage = 0; uint32_t a_phase = 0;
      for (int chunk = 0; chunk < kD / BC_BK; ++chunk) {
        const int c0 = chunk * BC_BK;
        mbar_wait(&a_empty[a_stage], a_phase ^ 1);
        mbar_expect_tx(&a_full[a_stage], BC_A_BYTES);
        tma_load_2d(smem + a_stage * BC_A_BYTES, &tma_x, &a_full[a_stage], c0, (int)(q0 - p.r_lo));   // OOB rows -> 0
        if (++a_stage == BC_A_STAGES) { a_stage = 0; a_phase ^= 1; }
        for (int t = -3; t <= 3; ++t) {
          const int at = t < 0 ? -t : t;
          const int n_w = (at <= 1) ? 3 : (at == 2 ? 2 : 1);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sw = smem_w + stage * BC_W_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], n_w * BC_W_BYTES);
          tma_load_2d(sw + 0 * BC_W_BYTES, &tma_w7, &full_bar[stage], (t + 3) * kD + c0, n0);
          if (n_w >= 2) tma_load_2d(sw + 1 * BC_W_BYTES, &tma_w5, &full_bar[stage], (t + 2) * kD + c0, n0);
          if (n_w >= 3) tma_load_2d(sw + 2 * BC_W_BYTES, &tma_w3, &full_bar[stage], (t + 1) * kD + c0, n0);
          if (++stage == BC_W_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int a_stage = 0; uint32_t a_phase = 0;
      for (int chunk = 0; chunk < kD / BC_BK; ++chunk) {
        mbar_wait(&a_full[a_stage], a_phase);
        tc_fence_after();
        const uint64_t da = make_smem_desc_sw128(smem_u32(smem + a_stage * BC_A_BYTES));
        for (int t = -3; t <= 3; ++t) {
          const int at = t < 0 ? -t : t;
          const int n_w = (at <= 1) ? 3 : (at == 2 ? 2 : 1);
          const uint32_t idesc = make_idesc_tf32(BC_BM, n_w * BC_NCH);           // one MMA covers every conv of the shift
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem_w + stage * BC_W_STAGE_BYTES));   // tap tiles stacked along N
          const uint32_t d_tmem = tmem_base + (uint32_t)group_col(t);
#pragma unroll
          for (int k = 0; k < BC_BK / 8; ++k)
            mma_tf32(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (chunk | k) != 0);
          mma_commit(&empty_bar[stage]);
          if (++stage == BC_W_STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit(&a_empty[a_stage]);                                           // all 15 taps of the chunk issued
        if (++a_stage == BC_A_STAGES) { a_stage = 0; a_phase ^= 1; }
      }
      mma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;                                    // row inside the tile = TMEM lane
    const int64_t rho = q0 + r;                                     // bank row of this thread
    const int64_t prow = rho - p.row_base;
    const bool valid = r >= 3 && r < 3 + BC_OUT && prow >= 0 && prow < p.pb_rows && rho < p.n_rows;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const float* x0p = p.bank + (valid ? rho : 0) * kD + n0;
    const bool has_next = valid && (rho + 1 < p.n_rows);
    const float* x1p = p.bank + (has_next ? rho + 1 : 0) * kD + n0;
    float* dst = p.pb + (valid ? prow : 0) * (7 * kD) + n0;
    // exchange layout: ex[tap][row][8]; taps 0..6 = conv7 t=-3..3, 7..11 = conv5 t=-2..2, 12..14 = conv3 t=-1..1
    auto exq = [&](int tap, int row) -> float* { return ex + ((size_t)tap * BC_BM + row) * 8; };
#pragma unroll 1
    for (int cc = 0; cc < BC_NCH; cc += 8) {
      {
        float v[8];
#pragma unroll
        for (int t = -3; t <= 3; ++t) {
          tmem_ld8(t_row + group_col(t) + cc, v);
          tmem_ld_wait();
          *reinterpret_cast<float4*>(exq(t + 3, r)) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(exq(t + 3, r) + 4) = make_float4(v[4], v[5], v[6], v[7]);
          if (t >= -2 && t <= 2) {
            tmem_ld8(t_row + group_col(t) + BC_NCH + cc, v);
            tmem_ld_wait();
            *reinterpret_cast<float4*>(exq(7 + t + 2, r)) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(exq(7 + t + 2, r) + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          if (t >= -1 && t <= 1) {
            tmem_ld8(t_row + group_col(t) + 2 * BC_NCH + cc, v);
            tmem_ld_wait();
            *reinterpret_cast<float4*>(exq(12 + t + 1, r)) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(exq(12 + t + 1, r) + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
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
          const float4 d = __ldg(reinterpret_cast<const float4*>(x0p + cc) + h);
          const float4 e = has_next ? __ldg(reinterpret_cast<const float4*>(x1p + cc) + h) : make_float4(0.f, 0.f, 0.f, 0.f);
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
      epi_barrier();                                   // exchange buffer is reused by the next 8 channels
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
  CUtensorMap tx, tw3, tw5, tw7;
  {
    uint64_t dims[2] = {(uint64_t)kD, (uint64_t)r_cnt};
    uint64_t str[1] = {(uint64_t)kD * 4};
    uint32_t box[2] = {BC_BK, BC_BM};
    TMR_TRY(make_tmap(&tx, bank_r, 2, dims, str, box));
    const float* pr = packed + TimeConvPacked::fp32_total;
    const float* w[3] = {pr + TimeConvPacked::w3_off, pr + TimeConvPacked::w5_off, pr + TimeConvPacked::w7_off};
    CUtensorMap* tw[3] = {&tw3, &tw5, &tw7};
    for (int i = 0; i < 3; ++i) {
      const int taps = 3 + 2 * i;
      uint64_t dw[2] = {(uint64_t)taps * kD, (uint64_t)kD};
      uint64_t sw[1] = {(uint64_t)taps * kD * 4};
      uint32_t bw[2] = {BC_BK, BC_NCH};
      TMR_TRY(make_tmap(tw[i], w[i], 2, dw, sw, bw));
    }
  }
  TMR_CUDA(cudaFuncSetAttribute(umma_bankconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM_BYTES));
  const int64_t tiles = ((pb_rows + BC_OUT - 1) / BC_OUT) * (kD / BC_NCH);
  umma_bankconv_kernel<<<(unsigned)tiles, BC_THREADS, BC_SMEM_BYTES, st>>>(tx, tw3, tw5, tw7, p);
  TMR_LAUNCH_CHECK("umma_bankconv_kernel");
  return TMR_OK;
}

}  // namespace tmr
