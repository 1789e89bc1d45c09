// cuda_shim.h — host stand-ins for the CUDA device intrinsics used by materialist_b200/csrc/*.cuh, so that the kernels' own
// per-sample device functions can be compiled with g++ and run on the CPU (TEST INFRASTRUCTURE: tests/test_host_emulation.py).
// Compile with -ffp-contract=off: the __f*_rn stand-ins are then single IEEE operations, like the intrinsics they replace.
// The approximate device functions (rsqrtf, __expf, sincospif, fast division) become libm calls here: those only feed values
// that are smooth in their inputs, which is the point of the float discipline being tested.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

using std::min;
using std::max;

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __saturatef(float x) { return x != x ? 0.f : fminf(fmaxf(x, 0.f), 1.f); }
static inline float emul_rsqrtf(float x) { return 1.f / sqrtf(x); }
static inline void emul_sincospif(float x, float* s, float* c) { *s = (float)sin(M_PI * (double)x); *c = (float)cos(M_PI * (double)x); }
#define rsqrtf emul_rsqrtf
#define sincospif emul_sincospif
static inline float emul_expf(float x) { return expf(x); }
#define __expf emul_expf
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31u; return s ? (lo >> s) | (hi << (32u - s)) : lo; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p += v; return o; }
static inline float4 atomicAdd(float4* p, float4 v) { float4 o = *p; p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w; return o; }
// one "thread" per block on the host: the CTA-cooperative staging loops degenerate to serial copies
static const uint3 threadIdx = {0, 0, 0};
static const dim3 blockDim = {1, 1, 1};
static inline void __syncthreads() {}
// warp primitives for a "warp" of one lane (the aggregated scatter code must compile; the emulation never calls it)
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __any_sync(unsigned, int p) { return p != 0; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline unsigned __match_any_sync(unsigned, unsigned) { return 1u; }
static inline float __shfl_sync(unsigned, float v, int) { return v; }
