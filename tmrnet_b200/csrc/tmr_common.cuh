// Shared helpers for libtmr_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/tmr_b200.h"

namespace tmr {

std::string& last_error_ref();
int set_error(int code, const char* fmt, ...);

#define TMR_CHECK_ARG(cond, ...)                                         \
  do {                                                                   \
    if (!(cond)) return ::tmr::set_error(TMR_ERR_ARG, __VA_ARGS__);      \
  } while (0)

#define TMR_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return ::tmr::set_error(TMR_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                   \
                              cudaGetErrorString(e__), __FILE__, __LINE__);                   \
  } while (0)

#define TMR_LAUNCH_CHECK(name)                                                                \
  do {                                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess)                                                                   \
      return ::tmr::set_error(TMR_ERR_CUDA, "launch of %s failed: %s", name,                  \
                              cudaGetErrorString(e__));                                       \
  } while (0)

#define TMR_TRY(call)            \
  do {                           \
    int rc__ = (call);           \
    if (rc__ != TMR_OK) return rc__; \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kD = 512;   // bank row width / LSTM hidden (reference hard-codes 512, NLB:26,17)
constexpr int kF = 2048;  // backbone feature width

// Round-to-nearest fp32 -> TF32 (10-bit mantissa).  tcgen05.mma.kind::tf32 ignores the low 13
// mantissa bits of its operands (truncation, a systematic shrink of every product); operands of
// the tensor-core path are therefore rounded once, by their producer, with cvt.rna.
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// ---- packed weight layouts (floats) ---------------------------------------------------------
// Every pack holds the fp32 layout below followed by a mirror of it rounded to TF32 (same offsets
// + fp32_total) that the tcgen05 path reads.
// TimeConv: Wp_K[o][tap][c] = w_K[o][c][tap]  (K-major rows of length K*D), then biases.
struct TimeConvPacked {
  static constexpr size_t w3_off = 0;
  static constexpr size_t w5_off = w3_off + (size_t)kD * 3 * kD;
  static constexpr size_t w7_off = w5_off + (size_t)kD * 5 * kD;
  static constexpr size_t b3_off = w7_off + (size_t)kD * 7 * kD;
  static constexpr size_t b5_off = b3_off + kD;
  static constexpr size_t b7_off = b5_off + kD;
  static constexpr size_t fp32_total = b7_off + kD;
  static constexpr size_t total = 2 * fp32_total;   // + TF32-rounded (RN) mirror for the tensor-core path
};
// NLBlock: W1[n][k], W2T[n][k] = W2[k][n], W3[n][k], W4[n][k], b1, b3, b4, ln_w, ln_b, then the query fold
// W21 = W2^T W1 ([n][k]) and bu = W2^T b1, so the tensor-core path gets u = W2^T (W1 St + b1) from ONE GEMM.
// (b2 cancels inside the softmax over L: q.(W2 l_k + b2) = (W2^T q).l_k + const.)
struct NLBlockPacked {
  static constexpr size_t w1_off = 0;
  static constexpr size_t w2t_off = w1_off + (size_t)kD * kD;
  static constexpr size_t w3_off = w2t_off + (size_t)kD * kD;
  static constexpr size_t w4_off = w3_off + (size_t)kD * kD;
  static constexpr size_t b1_off = w4_off + (size_t)kD * kD;
  static constexpr size_t b3_off = b1_off + kD;
  static constexpr size_t b4_off = b3_off + kD;
  static constexpr size_t lnw_off = b4_off + kD;
  static constexpr size_t lnb_off = lnw_off + kD;
  static constexpr size_t w21_off = lnb_off + kD;
  static constexpr size_t bu_off = w21_off + (size_t)kD * kD;
  static constexpr size_t fp32_total = bu_off + kD;
  static constexpr size_t total = 2 * fp32_total;
};
// LSTM: gate-interleaved rows r' = unit*4 + gate (gate order i,f,g,o) so one float4 of the
// projected row holds the four gates of a hidden unit.  Wih'[4D][F], Whh'[4D][D], bias'[4D] = bih+bhh.
struct LstmPacked {
  static constexpr size_t wih_off = 0;
  static constexpr size_t whh_off = wih_off + (size_t)4 * kD * kF;
  static constexpr size_t bias_off = whh_off + (size_t)4 * kD * kD;
  static constexpr size_t fp32_total = bias_off + 4 * kD;
  static constexpr size_t total = 2 * fp32_total;
};
// Classifier: Wh[D][2D], bh[D], Wc[C][D], bc[C] (C padded to 32 rows for alignment).
struct ClassifierPacked {
  static constexpr int kMaxC = 32;
  static constexpr size_t wh_off = 0;
  static constexpr size_t bh_off = wh_off + (size_t)kD * 2 * kD;
  static constexpr size_t wc_off = bh_off + kD;
  static constexpr size_t bc_off = wc_off + (size_t)kMaxC * kD;
  static constexpr size_t fp32_total = bc_off + kMaxC;
  static constexpr size_t total = 2 * fp32_total;
};

}  // namespace tmr
