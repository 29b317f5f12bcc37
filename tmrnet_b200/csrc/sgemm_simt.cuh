// fp32 CUDA-core GEMM main loop (TMR_MATH_FP32): C[128x128] += A[128xK] . W[128xK]^T, both operands
// K-major in global memory.  256 threads, 8x8 register tile per thread, 16-wide k slices staged in
// shared memory with register prefetch.  Row providers return a pointer to the 16 consecutive
// floats of (row, k-slice) or nullptr for an all-zero slice (out-of-range rows, conv zero padding).
#pragma once
#include <cuda_runtime.h>

namespace tmr {
namespace simt {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256, PAD = 4;

struct Smem {
  float a[2][BK][BM + PAD];
  float b[2][BK][BN + PAD];
};

__device__ __forceinline__ float4 ld4(const float* p) {
  return p ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// Thread (ty, tx) = (tid / 16, tid % 16) owns rows {ty*4+i, 64+ty*4+i} and cols {tx*4+j, 64+tx*4+j}.
__device__ __forceinline__ int tile_row(int ty, int i) { return (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4); }
__device__ __forceinline__ int tile_col(int tx, int j) { return (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

template <class AFn, class BFn>
__device__ __forceinline__ void mainloop(float (&acc)[8][8], AFn afn, BFn bfn, int k_tiles, Smem& s) {
  const int tid = threadIdx.x;
  const int lr = tid >> 2;          // 0..63: tile row (and +64) this thread stages
  const int kq = (tid & 3) * 4;     // k offset of its float4 inside the 16-wide slice
  const int ty = tid >> 4, tx = tid & 15;

  float4 ra0, ra1, rb0, rb1;
  auto fetch = [&](int kt) {
    const float* p;
    p = afn(lr, kt);       ra0 = ld4(p ? p + kq : nullptr);
    p = afn(lr + 64, kt);  ra1 = ld4(p ? p + kq : nullptr);
    p = bfn(lr, kt);       rb0 = ld4(p ? p + kq : nullptr);
    p = bfn(lr + 64, kt);  rb1 = ld4(p ? p + kq : nullptr);
  };
  auto stage = [&](int buf) {
    s.a[buf][kq + 0][lr] = ra0.x; s.a[buf][kq + 1][lr] = ra0.y; s.a[buf][kq + 2][lr] = ra0.z; s.a[buf][kq + 3][lr] = ra0.w;
    s.a[buf][kq + 0][lr + 64] = ra1.x; s.a[buf][kq + 1][lr + 64] = ra1.y; s.a[buf][kq + 2][lr + 64] = ra1.z; s.a[buf][kq + 3][lr + 64] = ra1.w;
    s.b[buf][kq + 0][lr] = rb0.x; s.b[buf][kq + 1][lr] = rb0.y; s.b[buf][kq + 2][lr] = rb0.z; s.b[buf][kq + 3][lr] = rb0.w;
    s.b[buf][kq + 0][lr + 64] = rb1.x; s.b[buf][kq + 1][lr + 64] = rb1.y; s.b[buf][kq + 2][lr + 64] = rb1.z; s.b[buf][kq + 3][lr + 64] = rb1.w;
  };

  __syncthreads();                  // previous users of the shared tiles are done
  fetch(0);
  stage(0);
  __syncthreads();
  for (int kt = 0; kt < k_tiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < k_tiles) fetch(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&s.a[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&s.a[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&s.b[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&s.b[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < k_tiles) {
      stage(buf ^ 1);               // the other buffer was last read in iteration kt-1 (barrier below)
      __syncthreads();
    }
  }
}

}  // namespace simt
}  // namespace tmr
