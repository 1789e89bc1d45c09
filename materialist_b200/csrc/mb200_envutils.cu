// mb200_envutils.cu — CUDA side of the helper surface of myutils/envmap_utils.py and myutils/computeSH.py.
//
//   mb200_cdf_build       build_envmap        envmap_utils.py:43-66   (marginal / conditional CDFs)
//   mb200_cdf_sample      sample_envmap       envmap_utils.py:172-201 (searchsorted importance sampling)
//   mb200_sh_project      computeSHFromImage  computeSH.py:299-347    (order-4 real SH, 25 coefficients, fp64)
//   mb200_sh_reconstruct  reconstImageFromSH  computeSH.py:226-240
//
// The CDF prefix sums accumulate sequentially in double and round each prefix to float (what torch's CPU cumsum
// does: acc_type<float> = double), one thread per row, so the integer searchsorted indices reproduce the
// reference's bit for bit; the SH kernels run in fp64 like the numpy reference.
#include "mb200_device.cuh"
#include "mb200_host.h"

namespace {

// ---------------------------------------------------------------- CDF build
__global__ void cdf_rows_kernel(const float* __restrict__ env, int h, int w, float* __restrict__ c_cdf, float* __restrict__ marg) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    const float h01 = (float)(((double)y + 0.5) / (double)h);
    const float sin_theta = sinf(XMUL(3.14159265358979323846f, h01));
    double acc = 0.0, tot = 0.0;
    for (int x = 0; x < w; ++x) {
        const float* t = env + ((size_t)y * w + x) * 3;
        const float lum = XADD(XADD(XMUL(0.299f, t[0]), XMUL(0.587f, t[1])), XMUL(0.114f, t[2]));
        acc += (double)XMUL(lum, sin_theta);
        const float c = (float)acc;
        c_cdf[(size_t)y * w + x] = c;
        tot += (double)c;                                  // marginal = sum of the CUMULATIVE row (envmap_utils.py:53)
    }
    marg[y] = (float)tot;
    const float last = XADD(c_cdf[(size_t)y * w + w - 1], 1e-6f);        // conditional_cdf / (conditional_cdf[:, -1] + 1e-6)
    for (int x = 0; x < w; ++x) c_cdf[(size_t)y * w + x] = XDIV(c_cdf[(size_t)y * w + x], last);
}
__global__ void cdf_marginal_kernel(const float* __restrict__ marg, int h, float* __restrict__ m_cdf) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double acc = 0.0;
    for (int y = 0; y < h; ++y) { acc += (double)marg[y]; m_cdf[y] = (float)acc; }
    const float last = XADD(m_cdf[h - 1], 1e-6f);
    for (int y = 0; y < h; ++y) m_cdf[y] = XDIV(m_cdf[y], last);
}
// ---------------------------------------------------------------- CDF sample
__device__ __forceinline__ int searchsorted_left(const float* __restrict__ a, int n, float x) {
    int lo = 0, hi = n;                       // first i with a[i] >= x   (torch.searchsorted default, right=False)
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}
__global__ void cdf_sample_kernel(const float* __restrict__ c_cdf, const float* __restrict__ m_cdf, int h, int w,
                                  const float* __restrict__ s2, long long n, float* __restrict__ dirs, float* __restrict__ pdf,
                                  long long* __restrict__ v_idx_out, long long* __restrict__ u_idx_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x0 = s2[i], x1 = s2[n + i];
    const int v_idx = searchsorted_left(m_cdf, h, x0);
    const int vi = min(v_idx, h - 1);
    const float prev = v_idx > 0 ? m_cdf[max(vi - 1, 0)] : 0.f;
    const float dv = v_idx > 0 ? XDIV(XSUB(x0, prev), XSUB(m_cdf[vi], prev)) : XDIV(x0, m_cdf[vi]);
    const float pdf_m = v_idx > 0 ? XSUB(m_cdf[vi], prev) : m_cdf[vi];
    const float v = XADD((float)v_idx, dv);
    const float* row = c_cdf + (size_t)vi * w;
    const int u_idx = searchsorted_left(row, w, x1);
    int ui = u_idx, ui1 = u_idx - 1;
    if (ui1 == -1) ui1 = 0;
    if (ui == 32) ui = 31;                    // hard-coded in the reference (envmap_utils.py:120)
    ui = min(ui, w - 1);
    const float pdf_c = u_idx > 0 ? XSUB(row[ui], row[ui1]) : row[ui];
    const float theta = XDIV(XMUL(v, 3.14159265358979323846f), (float)h);
    const float phi = XDIV(XMUL(XMUL(2.0f, (float)u_idx), 3.14159265358979323846f), (float)w);
    float st, ct, sp, cp; sincosf(theta, &st, &ct); sincosf(phi, &sp, &cp);
    float dx = st * cp, dy = st * sp, dz = ct;                       // angle2xyz (z-up) + F.normalize
    const float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
    dirs[3 * i] = dx * inv; dirs[3 * i + 1] = dy * inv; dirs[3 * i + 2] = dz * inv;
    const float two_pi_pi = 19.739208802178716f;
    pdf[i] = __fdiv_rn((float)(h * w) * (pdf_c * pdf_m), two_pi_pi * st + 1e-6f);
    v_idx_out[i] = v_idx; u_idx_out[i] = u_idx;
}

// ---------------------------------------------------------------- SH (fp64)
struct ShK { double K[25]; };
__device__ __forceinline__ void sh_basis(double theta, double phi, const ShK& k, double Y[25]) {
    const double c = cos(theta), s = sin(theta), c2 = c * c, s2 = s * s;
    const double P00 = 1.0, P10 = c, P11 = -s;
    const double P20 = 0.5 * (3 * c2 - 1), P21 = -3 * c * s, P22 = 3 * s2;
    const double P30 = 0.5 * (5 * c2 * c - 3 * c), P31 = -1.5 * (5 * c2 - 1) * s, P32 = 15 * c * s2, P33 = -15 * s2 * s;
    const double P40 = 0.125 * (35 * c2 * c2 - 30 * c2 + 3), P41 = -2.5 * (7 * c2 * c - 3 * c) * s, P42 = 7.5 * (7 * c2 - 1) * s2,
                 P43 = -105 * c * s2 * s, P44 = 105 * s2 * s2;
    const double r2 = 1.4142135623730951;
    double sn[5], cs[5];
    for (int m = 1; m <= 4; ++m) { sn[m] = sin(m * phi); cs[m] = cos(m * phi); }
    Y[0] = k.K[0] * P00;
    Y[1] = r2 * k.K[1] * sn[1] * P11; Y[2] = k.K[2] * P10; Y[3] = r2 * k.K[3] * cs[1] * P11;
    Y[4] = r2 * k.K[4] * sn[2] * P22; Y[5] = r2 * k.K[5] * sn[1] * P21; Y[6] = k.K[6] * P20;
    Y[7] = r2 * k.K[7] * cs[1] * P21; Y[8] = r2 * k.K[8] * cs[2] * P22;
    Y[9] = r2 * k.K[9] * sn[3] * P33; Y[10] = r2 * k.K[10] * sn[2] * P32; Y[11] = r2 * k.K[11] * sn[1] * P31; Y[12] = k.K[12] * P30;
    Y[13] = r2 * k.K[13] * cs[1] * P31; Y[14] = r2 * k.K[14] * cs[2] * P32; Y[15] = r2 * k.K[15] * cs[3] * P33;
    Y[16] = r2 * k.K[16] * sn[4] * P44; Y[17] = r2 * k.K[17] * sn[3] * P43; Y[18] = r2 * k.K[18] * sn[2] * P42;
    Y[19] = r2 * k.K[19] * sn[1] * P41; Y[20] = k.K[20] * P40; Y[21] = r2 * k.K[21] * cs[1] * P41;
    Y[22] = r2 * k.K[22] * cs[2] * P42; Y[23] = r2 * k.K[23] * cs[3] * P43; Y[24] = r2 * k.K[24] * cs[4] * P44;
}
// bilinear fetch uvToEnvmap (computeSH.py:75-85)
__device__ __forceinline__ void sh_fetch(const double* __restrict__ im, int h, int w, double u, double v, double col[3]) {
    const double c = u * (w - 1), r = (1 - v) * (h - 1);
    const int cs = (int)c, rs = (int)r, ce = min(w - 1, cs + 1), re = min(h - 1, rs + 1);
    const double wc = c - cs, wr = r - rs;
    for (int k = 0; k < 3; ++k) {
        const double c1 = (1 - wc) * im[((size_t)rs * w + cs) * 3 + k] + wc * im[((size_t)rs * w + ce) * 3 + k];
        const double c2 = (1 - wc) * im[((size_t)re * w + cs) * 3 + k] + wc * im[((size_t)re * w + ce) * 3 + k];
        col[k] = (1 - wr) * c1 + wr * c2;
    }
}
__global__ void __launch_bounds__(128) sh_project_kernel(const double* __restrict__ im, int h, int w, const double* __restrict__ angles,
                                                         long long n, const ShK k, double* __restrict__ coef) {
    __shared__ double s_acc[75];
    for (int i = threadIdx.x; i < 75; i += blockDim.x) s_acc[i] = 0.0;
    __syncthreads();
    double acc[75];
    for (int i = 0; i < 75; ++i) acc[i] = 0.0;
    const double pi = 3.141592653589793;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double theta = angles[2 * i], phi = angles[2 * i + 1];
        double Y[25], col[3];
        sh_basis(theta, phi, k, Y);
        sh_fetch(im, h, w, (phi + pi) / 2 / pi, 1 - theta / pi, col);
        for (int b = 0; b < 25; ++b) { acc[3 * b] += Y[b] * col[0]; acc[3 * b + 1] += Y[b] * col[1]; acc[3 * b + 2] += Y[b] * col[2]; }
    }
    for (int i = 0; i < 75; ++i) {
        double v = acc[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[i], v);
    }
    __syncthreads();
    const double Wt = 4 * pi / (double)n;
    for (int i = threadIdx.x; i < 75; i += blockDim.x) atomicAdd(&coef[i], s_acc[i] * Wt);
}
__global__ void sh_reconstruct_kernel(const double* __restrict__ coef, int nrows, int ncols, int clip, const ShK k, double* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * ncols) return;
    const int r = i / ncols, c = i % ncols;
    const double pi = 3.141592653589793;
    // np.linspace(-1, 1, ncols+1)[c] * pi, np.linspace(0, 1, nrows+1)[r] * pi
    const double phi = pi * (-1.0 + (double)c * (2.0 / (double)ncols)), theta = pi * ((double)r * (1.0 / (double)nrows));
    double Y[25]; sh_basis(theta, phi, k, Y);
    for (int ch = 0; ch < 3; ++ch) {
        double v = 0.0;
        for (int b = 0; b < 25; ++b) v += Y[b] * coef[3 * b + ch];
        if (clip) v = fmin(fmax(v, 0.0), 1.0);
        img[(size_t)i * 3 + ch] = v;
    }
}
ShK make_k() {
    static const int L[25] = {0, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4};
    static const int M[25] = {0, -1, 0, 1, -2, -1, 0, 1, 2, -3, -2, -1, 0, 1, 2, 3, -4, -3, -2, -1, 0, 1, 2, 3, 4};
    ShK k;
    for (int i = 0; i < 25; ++i) {                 // computeK (computeSH.py:58-68), factorials round-trip through float32
        const int l = L[i], m = M[i] < 0 ? -M[i] : M[i];
        double a = 1, b = 1;
        for (int j = 2; j <= l - m; ++j) a *= j;
        for (int j = 2; j <= l + m; ++j) b *= j;
        k.K[i] = sqrt((2 * l + 1) * (double)(float)a / (double)(float)b / 4 / 3.141592653589793);
    }
    return k;
}

}  // namespace

extern "C" {

int mb200_cdf_build(const float* env, int h, int w, float* c_cdf, float* m_cdf, void* stream) {
    if (!env || !c_cdf || !m_cdf || h < 1 || w < 1) return MB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    // the per-row sums are staged in m_cdf itself and consumed by the single-thread marginal scan
    cdf_rows_kernel<<<(h + 63) / 64, 64, 0, st>>>(env, h, w, c_cdf, m_cdf);
    cdf_marginal_kernel<<<1, 32, 0, st>>>(m_cdf, h, m_cdf);
    return mb200_check_launch();
}

int mb200_cdf_sample(const float* c_cdf, const float* m_cdf, int h, int w, const float* sample2, int64_t n,
                     float* dirs, float* pdf, int64_t* v_idx, int64_t* u_idx, void* stream) {
    if (!c_cdf || !m_cdf || !sample2 || !dirs || !pdf || !v_idx || !u_idx || h < 1 || w < 1 || n < 0) return MB200_EINVAL;
    if (n == 0) return MB200_OK;
    cdf_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c_cdf, m_cdf, h, w, sample2, n, dirs, pdf,
                                                                                  reinterpret_cast<long long*>(v_idx), reinterpret_cast<long long*>(u_idx));
    return mb200_check_launch();
}

int mb200_sh_project(const double* im, int h, int w, const double* angles, int64_t n, double* coef, void* stream) {
    if (!im || !angles || !coef || h < 1 || w < 1 || n < 1) return MB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = mb200_check(cudaMemsetAsync(coef, 0, sizeof(double) * 75, st));
    if (rc) return rc;
    long long blocks = (n + 127) / 128; if (blocks > 4 * mb200_sm_count()) blocks = 4 * mb200_sm_count();
    sh_project_kernel<<<(unsigned)blocks, 128, 0, st>>>(im, h, w, angles, n, make_k(), coef);
    return mb200_check_launch();
}

int mb200_sh_reconstruct(const double* coef, int nrows, int ncols, int clip, double* img, void* stream) {
    if (!coef || !img || nrows < 1 || ncols < 1) return MB200_EINVAL;
    const int n = nrows * ncols;
    sh_reconstruct_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(coef, nrows, ncols, clip, make_k(), img);
    return mb200_check_launch();
}

}  // extern "C"
