// mb200_env.cu — envmap ingest + Hierarchical2D construction on the GPU, and the adjoint of the ingest map.
//
// Replaces EnvironmentMapEmitter::parameters_changed('data') of mitsuba 3.5.2 (src/emitters/envmap.cpp;
// reached from inverse_img_w_mi.py:63-64 on every envmap-phase iteration), which migrates the texture to the
// HOST and rebuilds the sampling hierarchy on one CPU thread.  Here it stays on the device.
// All arithmetic is IEEE round-to-nearest without contraction so the pyramid is bit-identical to the oracle's.
#include "mb200_device.cuh"
#include "mb200_host.h"

using namespace mb;

namespace {

__global__ void env_ingest_kernel(const float* __restrict__ env_in, int He, int We, int Wi, int mode,
                                  float4* __restrict__ env4, float* __restrict__ lum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= He * Wi) return;
    const int y = i / Wi, x = i % Wi;
    float t[3];
    if (mode == MB200_ENV_FILE) {
        const int xs = x == We ? 0 : x;                     // appended column = copy of column 0
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c] = env_in[((size_t)y * We + xs) * 3 + c];
    } else if (x == 0 || x == Wi - 1) {                     // horizontal continuity: average first / last
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c] = XMUL(.5f, XADD(env_in[((size_t)y * We) * 3 + c], env_in[((size_t)y * We + We - 1) * 3 + c]));
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c] = env_in[((size_t)y * We + x) * 3 + c];
    }
    env4[i] = make_float4(t[0], t[1], t[2], 0.f);
    const float theta_scale = XMUL(XDIV(1.f, (float)(He - 1)), 3.14159265358979323846f);
    const float theta = XMUL((float)y, theta_scale);
    const float sin_theta = (float)sin((double)theta);
    const float l = XADD(XADD(XMUL(t[0], 0.212671f), XMUL(t[1], 0.715160f)), XMUL(t[2], 0.072169f));
    lum[i] = XMUL(l, sin_theta);
}

__device__ __forceinline__ float patch_avg(const float* __restrict__ d, int rx, int x, int y) {
    const float v00 = d[y * rx + x], v10 = d[y * rx + x + 1], v01 = d[(y + 1) * rx + x], v11 = d[(y + 1) * rx + x + 1];
    return XMUL(.25f, XADD(XADD(XADD(v00, v10), v01), v11));
}
// one thread per patch row, sequential double accumulation (the order the oracle uses)
__global__ void env_row_sums_kernel(const float* __restrict__ lum, int rx, int ry, double* __restrict__ rowsum) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= ry - 1) return;
    double rs = 0.0;
    for (int x = 0; x < rx - 1; ++x) rs += (double)patch_avg(lum, rx, x, y);
    rowsum[y] = rs;
}
__global__ void env_scale_kernel(const double* __restrict__ rowsum, int rx, int ry, float* __restrict__ scale_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double sum = 0.0;
    for (int y = 0; y < ry - 1; ++y) sum += rowsum[y];
    *scale_out = XDIV((float)((double)(rx - 1) * (double)(ry - 1)), (float)sum);
}
__global__ void env_level1_kernel(const float* __restrict__ lum, int rx, int ry, const float* __restrict__ scale,
                                  float* __restrict__ l1, int w1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int npx = rx - 1, npy = ry - 1;
    if (i >= npx * npy) return;
    const int y = i / npx, x = i % npx;
    l1[lvl_index((uint32_t)x, (uint32_t)y, (uint32_t)w1)] = XMUL(patch_avg(lum, rx, x, y), *scale);
}
__global__ void env_scale_level0_kernel(float* __restrict__ l0, int n, const float* __restrict__ scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) l0[i] = XMUL(l0[i], *scale);
}
__global__ void env_upper_level_kernel(const float* __restrict__ child, int cw, int ch, float* __restrict__ parent, int pw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int hw = cw / 2, hh = ch / 2;
    if (i >= hw * hh) return;
    const int y = i / hw, x = i % hw;
    const float4 q = *reinterpret_cast<const float4*>(child + lvl_index((uint32_t)(2 * x), (uint32_t)(2 * y), (uint32_t)cw));
    parent[lvl_index((uint32_t)x, (uint32_t)y, (uint32_t)pw)] = XADD(XADD(XADD(q.x, q.y), q.z), q.w);
}

// sums the privatised slabs (fixed order) and applies the adjoint of the ingest map
__global__ void env_grad_finish_kernel(const float4* __restrict__ g4, int n_slabs, int He, int We, int Wi, int mode, float* __restrict__ g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= He * We) return;
    const int y = i / We, x = i % We;
    const size_t stride = (size_t)He * Wi;
    auto texel = [&](size_t idx) {
        float3 a = make_float3(0.f, 0.f, 0.f);
        for (int s = 0; s < n_slabs; ++s) { const float4 t = __ldg(g4 + (size_t)s * stride + idx); a.x += t.x; a.y += t.y; a.z += t.z; }
        return a;
    };
    float3 v = texel((size_t)y * Wi + x);
    if (mode == MB200_ENV_FILE) {
        if (x == 0) { const float3 e = texel((size_t)y * Wi + We); v.x += e.x; v.y += e.y; v.z += e.z; }
    } else if (x == 0 || x == We - 1) {
        const float3 p = texel((size_t)y * Wi), q = texel((size_t)y * Wi + Wi - 1);
        v = make_float3(.5f * (p.x + q.x), .5f * (p.y + q.y), .5f * (p.z + q.z));
    }
    g[3 * (size_t)i] = v.x; g[3 * (size_t)i + 1] = v.y; g[3 * (size_t)i + 2] = v.z;
}

// Slab reduction, parallel over slabs: a 256-thread CTA owns 8 texels; the 32 lanes of a warp each add slabs lane, lane + 32, ...
// of their texel in a fixed order, then a fixed shuffle tree combines the lanes (deterministic); the sum is written in place
// over slab 0.  (One thread per texel summing up to 1 184 slabs serially ran the 16x32 map on 4 CTAs: 203 us, profiles/r2z.)
__global__ void __launch_bounds__(256) env_slab_reduce_kernel(float4* __restrict__ g4, int n_slabs, int n_texels) {
    const int lane = threadIdx.x & 31, t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= n_texels) return;
    float3 a = make_float3(0.f, 0.f, 0.f);
    for (int s = lane; s < n_slabs; s += 32) { const float4 v = g4[(size_t)s * n_texels + t]; a.x += v.x; a.y += v.y; a.z += v.z; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_down_sync(0xffffffffu, a.x, o); a.y += __shfl_down_sync(0xffffffffu, a.y, o); a.z += __shfl_down_sync(0xffffffffu, a.z, o);
    }
    if (lane == 0) g4[t] = make_float4(a.x, a.y, a.z, 0.f);
}

}  // namespace

extern "C" {

int mb200_env_prepare(const float* env_in, int He, int We, int mode, float* env4, float* hier,
                      const mb200_hier_desc* d, void* scratch, void* stream) {
    if (!env_in || !env4 || !hier || !d || !scratch || He < 2 || We < 2) return MB200_EINVAL;
    if (mode != MB200_ENV_ASSIGNED && mode != MB200_ENV_FILE) return MB200_EINVAL;
    const int Wi = mb200_env_internal_width(We, mode);
    if (d->res_x != Wi || d->res_y != He || d->n_levels < 2) return MB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    double* rowsum = reinterpret_cast<double*>(scratch);
    float* scale = reinterpret_cast<float*>(rowsum + He + 1);
    const int tb = 256, n0 = He * Wi, np = (Wi - 1) * (He - 1);
    int rc = mb200_check(cudaMemsetAsync(hier, 0, sizeof(float) * (size_t)d->total_floats, st));
    if (rc) return rc;
    env_ingest_kernel<<<(n0 + tb - 1) / tb, tb, 0, st>>>(env_in, He, We, Wi, mode, reinterpret_cast<float4*>(env4), hier);
    env_row_sums_kernel<<<(He - 1 + 63) / 64, 64, 0, st>>>(hier, Wi, He, rowsum);
    env_scale_kernel<<<1, 32, 0, st>>>(rowsum, Wi, He, scale);
    env_level1_kernel<<<(np + tb - 1) / tb, tb, 0, st>>>(hier, Wi, He, scale, hier + d->lvl_off[1], d->lvl_w[1]);
    env_scale_level0_kernel<<<(n0 + tb - 1) / tb, tb, 0, st>>>(hier, n0, scale);
    for (int l = 2; l < d->n_levels; ++l) {
        const int n = (d->lvl_w[l - 1] / 2) * (d->lvl_h[l - 1] / 2);
        env_upper_level_kernel<<<(n + tb - 1) / tb, tb, 0, st>>>(hier + d->lvl_off[l - 1], d->lvl_w[l - 1], d->lvl_h[l - 1],
                                                              hier + d->lvl_off[l], d->lvl_w[l]);
    }
    return mb200_check_launch();
}

int mb200_env_grad_slabs(int He, int We, int mode) {
    // privatised copies of the envmap-gradient grid, sized to stay L2-resident (<= 64 MB in total): the adjoint kernel's
    // CTAs spread their texel updates over them (slab = blockIdx % n), so a sun texel that attracts half of all emitter
    // samples is n different L2 addresses instead of one serialised atomic.
    if (He < 2 || We < 2) return 1;
    const long long bytes = (long long)He * mb200_env_internal_width(We, mode) * 16;
    long long n = (64ll << 20) / bytes;
    const long long cap = (long long)mb200_sm_count() * 8;
    if (n > cap) n = cap;
    return n < 1 ? 1 : (int)n;
}

int mb200_env_grad_finish(float* g_env4, int n_slabs, int He, int We, int mode, float* g_env, void* stream) {
    if (!g_env4 || !g_env || He < 2 || We < 2 || n_slabs < 1) return MB200_EINVAL;
    const int Wi = mb200_env_internal_width(We, mode), n = He * We, tb = 128;
    if (n_slabs > 1) {      // g_env4 is scratch owned by this call sequence (the caller zeroes it before every adjoint): reduce in place
        env_slab_reduce_kernel<<<(He * Wi + 7) / 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(g_env4), n_slabs, He * Wi);
        n_slabs = 1;
    }
    env_grad_finish_kernel<<<(n + tb - 1) / tb, tb, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(g_env4), n_slabs, He, We, Wi, mode, g_env);
    return mb200_check_launch();
}

}  // extern "C"
