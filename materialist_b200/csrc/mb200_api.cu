// mb200_api.cu — error handling, device queries and the small host-only helpers of the C-ABI.
#include <stdio.h>
#include "mb200_host.h"

namespace {
thread_local char g_err[256] = "";
int g_sm_count[64] = {0};
}

int mb200_check(cudaError_t e) {
    if (e == cudaSuccess) return MB200_OK;
    snprintf(g_err, sizeof(g_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    return MB200_ELAUNCH;
}
int mb200_check_launch() { return mb200_check(cudaGetLastError()); }

int mb200_sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (g_sm_count[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_sm_count[dev] = n;
    }
    return g_sm_count[dev];
}

__global__ void __launch_bounds__(256) probe_ffma_kernel(float* out, int iters) {
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = 1.0f + 1e-3f * (float)(threadIdx.x + k);
    const float b = 0.999f, c = 1e-4f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fmaf(a[k], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keeps the chains alive, (almost) never stores
}

extern "C" {

int mb200_probe_ffma(float* out, int iters, double* flops_host, void* stream) {
    if (!out || iters <= 0) return MB200_EINVAL;
    const int grid = mb200_sm_count() * 8;
    probe_ffma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, iters);
    if (flops_host) *flops_host = 2.0 * 16.0 * (double)iters * 256.0 * (double)grid;
    return mb200_check_launch();
}

const char* mb200_strerror(int code) {
    switch (code) {
        case MB200_OK: return "ok";
        case MB200_EINVAL: return "invalid argument";
        case MB200_ERANGE: return "size out of supported range";
        case MB200_ELAUNCH: return "CUDA launch/runtime error";
        case MB200_EUNSUPPORTED: return "not supported by this build";
        case MB200_EIO: return "file could not be written completely";
        default: return "unknown error";
    }
}
const char* mb200_last_cuda_error(void) { return g_err; }
int mb200_version(void) { return 100; }

int mb200_env_internal_width(int We, int mode) { return mode == MB200_ENV_FILE ? We + 1 : We; }

static int log2i_ceil(unsigned x) { int l = 0; while ((1u << l) < x) ++l; return l; }

int mb200_hier_describe(int res_x, int res_y, mb200_hier_desc* d) {
    if (!d || res_x < 2 || res_y < 2) return MB200_EINVAL;
    memset(d, 0, sizeof(*d));
    d->res_x = res_x; d->res_y = res_y;
    const int npx = res_x - 1, npy = res_y - 1;
    const int max_level = log2i_ceil((unsigned)(npx > npy ? npx : npy));
    int n = 0; long long off = 0;
    d->lvl_off[n] = 0; d->lvl_w[n] = res_x; d->lvl_h[n] = res_y; off += (long long)res_x * res_y; ++n;
    int sx = npx, sy = npy;
    for (int level = max_level; level >= 0; --level) {
        sx += sx & 1; sy += sy & 1;
        if (n >= MB200_MAX_LEVELS) return MB200_ERANGE;
        off = (off + 3) & ~3ll;                       // float4-aligned levels
        d->lvl_off[n] = (int)off; d->lvl_w[n] = sx; d->lvl_h[n] = sy; off += (long long)sx * sy; ++n;
        sx >>= 1; sy >>= 1;
    }
    if (off >= (1ll << 31)) return MB200_ERANGE;
    d->n_levels = n; d->total_floats = (int)off;
    return MB200_OK;
}

size_t mb200_env_scratch_bytes(int res_x, int res_y) {
    (void)res_x;
    return sizeof(double) * (size_t)(res_y + 8);      // row sums + total/scale slots
}

}  // extern "C"
