/* TEST INFRASTRUCTURE (like everything under oracle/): array front-ends of the shared reproducible float32 functions of
 * include/mb200_exact_math.h as gcc compiles them — tests/test_exact_math.py checks them against float64 libm, and the GPU tests
 * compare the kernels' results (same header, nvcc) bit for bit.  The PosMLP / CDF / SH oracle units are numpy (aux_oracle.py). */
#include <stddef.h>
#include "../include/mb200_exact_math.h"

void mbo_exact_sincospi(const float* x, size_t n, float* s, float* c) { for (size_t i = 0; i < n; ++i) mbx_sincospi(x[i], s + i, c + i); }
void mbo_exact_atan2(const float* y, const float* x, size_t n, float* o) { for (size_t i = 0; i < n; ++i) o[i] = mbx_atan2(y[i], x[i]); }
void mbo_exact_acos(const float* x, size_t n, float* o) { for (size_t i = 0; i < n; ++i) o[i] = mbx_acos(x[i]); }
void mbo_exact_asin01(const float* x, size_t n, float* o) { for (size_t i = 0; i < n; ++i) o[i] = mbx_asin01(x[i]); }
void mbo_exact_rsqrt(const float* x, size_t n, float* o) { for (size_t i = 0; i < n; ++i) o[i] = mbx_rsqrt(x[i]); }
