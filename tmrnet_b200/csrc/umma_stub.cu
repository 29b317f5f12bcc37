// Placeholder for the tcgen05 path until umma_gemm.cu lands: TMR_MATH_TF32 reports "unsupported"
// loudly instead of silently running something else.
#include "tmr_internal.h"
namespace tmr {
#ifndef TMR_HAVE_UMMA
bool umma_available() { return false; }
int umma_linear(const LinearArgs&, cudaStream_t) { return set_error(TMR_ERR_UNSUPPORTED, "tcgen05 path not built"); }
int umma_timeconv(const float*, const float*, int, int, float*, cudaStream_t) { return set_error(TMR_ERR_UNSUPPORTED, "tcgen05 path not built"); }
int umma_lstm_step(const float*, const float*, const int64_t*, int, int, const float*, float*, float*, int, cudaStream_t) { return set_error(TMR_ERR_UNSUPPORTED, "tcgen05 path not built"); }
#endif
}  // namespace tmr
