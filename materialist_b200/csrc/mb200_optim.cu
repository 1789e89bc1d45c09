// mb200_optim.cu — the loss and optimiser step that sit between the forward and the adjoint render of one
// BRDF-phase iteration (inverse_img_w_mi.py:388-432), fused into four small kernels so that the iteration is
// ~9 launches instead of ~100 elementwise torch launches and never syncs with the host:
//
//   image_sum      Σ pred                       -> ratio = gt.mean() / pred.detach().mean()      (:388)
//   srgb_sums      Σ diff², Σ |diff|            -> loss_mse, loss_l1, scale_raito                 (:391-395, :415)
//   srgb_grad      d loss / d pred              (autograd of :389-418 written out)
//   adam_clamped   clamp backward + aux L1 gradient + Adam + clamp forward of the next iteration (:370-376, :398-417, :428)
//
// All three are HBM-bound streaming passes over (rows*W*3) floats; reductions are two-stage with a FIXED
// summation order (per-thread strided partial -> warp shuffle -> block -> the last block sums the per-block
// partials), so results are bitwise reproducible run to run.
#include <math.h>
#include "mb200_host.h"

namespace {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 1024;
constexpr float kSrgbExp = 1.0f / 2.2f;        // myutils/misc.py:167-170  image ** (1/2.2)

struct RedScratch { float partial[2][kRedMaxBlocks]; unsigned int ticket; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-level reduce of NV values, then the last block to arrive reduces the per-block partials in index order
template <int NV>
__device__ __forceinline__ void finish_reduce(float (&v)[NV], RedScratch* sc, float* out) {
    __shared__ float s_part[NV][kRedThreads / 32];
    __shared__ bool s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const float w = warp_sum(v[k]);
        if (lane == 0) s_part[k][warp] = w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float s = 0.f;
            for (int i = 0; i < kRedThreads / 32; ++i) s += s_part[k][i];
            sc->partial[k][blockIdx.x] = s;
        }
        __threadfence();
        const unsigned int t = atomicAdd(&sc->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (warp < NV) {
        // one warp per value: lane-strided partial sums (fixed order), then a shuffle tree
        float s = 0.f;
        for (int i = lane; i < (int)gridDim.x; i += 32) s += __ldcg(&sc->partial[warp][i]);
        s = warp_sum(s);
        if (lane == 0) out[warp] = s;
    }
    if (threadIdx.x == 0) sc->ticket = 0u;           // ready for the next launch on the same stream
}

__global__ void __launch_bounds__(kRedThreads) image_sum_kernel(const float* __restrict__ img, long long n, RedScratch* sc, float* out) {
    float v[1] = {0.f};
    const long long n4 = n >> 2;
    const float4* img4 = reinterpret_cast<const float4*>(img);
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kRedThreads) {
        const float4 q = __ldg(img4 + i);
        v[0] += (q.x + q.y) + (q.z + q.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) v[0] += img[(n4 << 2) + threadIdx.x];
    finish_reduce<1>(v, sc, out);
}

// ratio = num/den read from device memory: scal[0] = Σ gt (global), scal[1] = Σ pred (global)
__device__ __forceinline__ float srgb_of(float x) { return x > 0.f ? powf(x, kSrgbExp) : 0.f; }

__global__ void __launch_bounds__(kRedThreads) srgb_sums_kernel(const float* __restrict__ img, const float* __restrict__ gt_srgb, long long n,
                                                                const float* __restrict__ scal, RedScratch* sc, float* out2,
                                                                float* __restrict__ pred_srgb) {
    const float ratio = __fdiv_rn(scal[0], scal[1]);
    float v[2] = {0.f, 0.f};
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kRedThreads) {
        const float y = srgb_of(__ldg(img + i) * ratio);
        const float d = y - __ldg(gt_srgb + i);
        v[0] = fmaf(d, d, v[0]); v[1] += fabsf(d);
        if (pred_srgb) pred_srgb[i] = y;
    }
    finish_reduce<2>(v, sc, out2);
}

// loss = 3 * (S1/S0) * Σdiff²/n_total + Σ|diff|/n_total with S1/S0 and ratio detached (:388, :415-417)
__global__ void __launch_bounds__(256) srgb_grad_kernel(const float* __restrict__ img, const float* __restrict__ gt_srgb, long long n,
                                                        const float* __restrict__ scal, const float* __restrict__ sums2, float inv_n_total,
                                                        float* __restrict__ grad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float ratio = __fdiv_rn(scal[0], scal[1]);
    const float k_mse = 6.f * __fdiv_rn(sums2[1], sums2[0]) * inv_n_total;      // 3 * scale_raito * 2 / n
    const float x = __ldg(img + i) * ratio;
    float g = 0.f;
    if (x > 0.f) {
        const float y = powf(x, kSrgbExp);
        const float d = y - __ldg(gt_srgb + i);
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        const float dy = fmaf(k_mse, d, sgn * inv_n_total);
        g = dy * (kSrgbExp * __fdiv_rn(y, x)) * ratio;                          // d(x^e)/dx = e * x^(e-1) = e * y / x
    }
    grad[i] = g;
}

struct AdamSeg { float* p; float* mat; const float* g; const float* ori; float* m; float* v; long long n; float lo, hi, aux; };
struct AdamParams { AdamSeg seg[MB200_ADAM_MAX_SEGS]; int nseg; float lr_over_bc1, inv_sqrt_bc2, beta1, beta2, eps; };

// torch.optim.Adam (single-tensor formulas) on p with gradient
//   mask(lo <= p <= hi) * (g_render + aux * sign(clamp(p) - ori))       [clamp backward + NF.l1_loss(mat, ori) * scale_delta]
// then mat = clamp(p_new) — the map the next iteration renders with.
__global__ void __launch_bounds__(256) adam_clamped_kernel(const __grid_constant__ AdamParams P) {
    const AdamSeg& s = P.seg[blockIdx.y];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += (long long)gridDim.x * blockDim.x) {
        float p = s.p[i];
        float g = 0.f;
        if (p >= s.lo && p <= s.hi) {
            const float d = p - s.ori[i];                   // inside the clamp range clamp(p) == p
            g = s.g[i] + s.aux * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        }
        float m = s.m[i], v = s.v[i];
        m = fmaf(g - m, 1.f - P.beta1, m);                  // exp_avg.lerp_(grad, 1 - beta1)
        v = fmaf(v, P.beta2, (1.f - P.beta2) * g * g);      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = __fsqrt_rn(v) * P.inv_sqrt_bc2 + P.eps;
        p = p - P.lr_over_bc1 * __fdiv_rn(m, denom);
        s.m[i] = m; s.v[i] = v; s.p[i] = p;
        s.mat[i] = fminf(fmaxf(p, s.lo), s.hi);
    }
}

int red_grid(long long n) {
    long long b = (n + kRedThreads * 8 - 1) / (kRedThreads * 8);
    const long long cap = (long long)mb200_sm_count() * 4;
    if (b > cap) b = cap;
    if (b > kRedMaxBlocks) b = kRedMaxBlocks;
    return b < 1 ? 1 : (int)b;
}

}  // namespace

extern "C" {

size_t mb200_reduce_scratch_bytes(void) { return sizeof(RedScratch); }

int mb200_image_sum(const float* img, int64_t n, float* out, void* scratch, void* stream) {
    if (!img || !out || !scratch || n <= 0) return MB200_EINVAL;
    if (((uintptr_t)img & 15) != 0) return MB200_EINVAL;
    image_sum_kernel<<<red_grid(n), kRedThreads, 0, (cudaStream_t)stream>>>(img, (long long)n, (RedScratch*)scratch, out);
    return mb200_check_launch();
}

int mb200_loss_srgb_sums(const float* img, const float* gt_srgb, int64_t n, const float* gt_pred_sums, float* out2,
                         float* pred_srgb_opt, void* scratch, void* stream) {
    if (!img || !gt_srgb || !gt_pred_sums || !out2 || !scratch || n <= 0) return MB200_EINVAL;
    srgb_sums_kernel<<<red_grid(n), kRedThreads, 0, (cudaStream_t)stream>>>(img, gt_srgb, (long long)n, gt_pred_sums, (RedScratch*)scratch,
                                                                            out2, pred_srgb_opt);
    return mb200_check_launch();
}

int mb200_loss_srgb_grad(const float* img, const float* gt_srgb, int64_t n, const float* gt_pred_sums, const float* sums2,
                         int64_t n_total, float* grad, void* stream) {
    if (!img || !gt_srgb || !gt_pred_sums || !sums2 || !grad || n <= 0 || n_total <= 0) return MB200_EINVAL;
    const int tb = 256;
    srgb_grad_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, (cudaStream_t)stream>>>(img, gt_srgb, (long long)n, gt_pred_sums, sums2,
                                                                                     1.f / (float)n_total, grad);
    return mb200_check_launch();
}

int mb200_adam_clamped(const mb200_adam_seg* segs, int nseg, float lr, float beta1, float beta2, float eps, int step, void* stream) {
    if (!segs || nseg <= 0 || nseg > MB200_ADAM_MAX_SEGS || step <= 0) return MB200_EINVAL;
    AdamParams P; memset(&P, 0, sizeof(P));
    long long nmax = 0;
    for (int k = 0; k < nseg; ++k) {
        const mb200_adam_seg& s = segs[k];
        if (!s.p || !s.mat || !s.g || !s.m || !s.v || s.n <= 0 || (s.aux_coeff != 0.f && !s.ori)) return MB200_EINVAL;
        AdamSeg& d = P.seg[k];
        d.p = s.p; d.mat = s.mat; d.g = s.g; d.ori = s.ori ? s.ori : s.p; d.m = s.m; d.v = s.v; d.n = s.n; d.lo = s.lo; d.hi = s.hi; d.aux = s.aux_coeff;
        if (s.n > nmax) nmax = s.n;
    }
    P.nseg = nseg; P.beta1 = beta1; P.beta2 = beta2; P.eps = eps;
    P.lr_over_bc1 = (float)((double)lr / (1.0 - pow((double)beta1, (double)step)));
    P.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)step)));
    long long b = (nmax + 255) / 256; const long long cap = (long long)mb200_sm_count() * 8;
    if (b > cap) b = cap;
    adam_clamped_kernel<<<dim3((unsigned)b, (unsigned)nseg), 256, 0, (cudaStream_t)stream>>>(P);
    return mb200_check_launch();
}

}  // extern "C"
