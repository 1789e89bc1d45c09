// mb200_lanes.cu — the inner (plugin) boundary: MatDiffBSDF.eval_pdf / .sample on arrays of lanes
// (myutils/mi_plugin.py:1429-1460).  One thread per lane; used for unit parity of a single BSDF
// evaluation against the oracle and the reference's torch BRDF sub-terms, not by the render hot loop.
#include "mb200_device.cuh"
#include "mb200_host.h"

using namespace mb;

namespace {

struct LaneParams {
    CamView cam; int use_mesh_normal; long long L;
    const float *p, *n_geo, *wi, *wo, *s1, *s2, *a, *r, *m, *n_opt;
    float *o3, *o1, *ow;
    TransView trans; long long* oflat;          // mb200_trans_* only
};
__device__ __forceinline__ float3 ld3(const float* q, long long i) { return f3(q[3 * i], q[3 * i + 1], q[3 * i + 2]); }
__device__ __forceinline__ void st3(float* q, long long i, float3 v) { q[3 * i] = v.x; q[3 * i + 1] = v.y; q[3 * i + 2] = v.z; }
__device__ __forceinline__ Material lane_material(const LaneParams& P, long long i) {
    const long long flat = texel_index(P.cam, ld3(P.p, i));
    Material mt; mt.a = ld3(P.a, flat); mt.r = P.r[flat]; mt.m = P.m[flat];
    mt.n = (P.use_mesh_normal || !P.n_opt) ? ld3(P.n_geo, i) : ld3(P.n_opt, flat);
    return mt;
}
__global__ void bsdf_eval_pdf_kernel(const __grid_constant__ LaneParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.L) return;
    const Material mt = lane_material(P, i);
    const BsdfVal v = eval_brdf(ld3(P.wo, i), ld3(P.wi, i), mt);   // eval_brdf(wi := light (wo), wo := view (si.wi))
    st3(P.o3, i, v.f); P.o1[i] = v.pdf;
}
// adjoint of eval_pdf's rgb value: cotangent in P.s2 (L,3) -> g_a (o3), g_r (o1), g_m (ow, 1 per lane), g_n (og)
__global__ void bsdf_eval_grad_kernel(const __grid_constant__ LaneParams P, float* __restrict__ og) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.L) return;
    const Material mt = lane_material(P, i);
    const BsdfGrad g = eval_brdf_grad<true>(ld3(P.wo, i), ld3(P.wi, i), mt, ld3(P.s2, i));
    st3(P.o3, i, g.ga); P.o1[i] = g.gr; P.ow[i] = g.gm; st3(og, i, g.gn);
}
__global__ void bsdf_sample_kernel(const __grid_constant__ LaneParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.L) return;
    const Material mt = lane_material(P, i);
    const Frame fs = make_frame(mt.n);
    const BsdfSample s = sample_brdf(P.s1[i], P.s2[2 * i], P.s2[2 * i + 1], ld3(P.wi, i), mt, fs);
    st3(P.o3, i, s.wi); P.o1[i] = s.pdf; st3(P.ow, i, s.weight);
}
// TransBSDF.eval_pdf / .sample / calculate_refracted_screen_coor on lanes (mi_plugin.py:1503-1544, 1748-1761)
__global__ void trans_eval_pdf_kernel(const __grid_constant__ LaneParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.L) return;
    const Material mt = lane_material(P, i);
    const float3 p = ld3(P.p, i), view = ld3(P.wi, i);
    const TransMat tm = trans_fetch(P.cam, P.trans, texel_index(P.cam, p), view, ld3(P.n_geo, i), p);
    const BsdfVal v = trans_eval_brdf(ld3(P.wo, i), view, mt, tm, P.trans);
    st3(P.o3, i, v.f); P.o1[i] = v.pdf;
}
__global__ void trans_sample_kernel(const __grid_constant__ LaneParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.L) return;
    const Material mt = lane_material(P, i);
    const float3 p = ld3(P.p, i), view = ld3(P.wi, i);
    const TransMat tm = trans_fetch(P.cam, P.trans, texel_index(P.cam, p), view, ld3(P.n_geo, i), p);
    const BsdfSample s = trans_sample_brdf(P.s1[i], P.s2[2 * i], P.s2[2 * i + 1], view, mt, tm, P.trans, make_frame(mt.n));
    st3(P.o3, i, s.wi); P.o1[i] = s.pdf; st3(P.ow, i, s.weight);
}
// the shared reproducible functions of include/mb200_exact_math.h (and the branch-free division / square root of the hierarchy
// descent) as nvcc compiles them, on arrays: tests compare them bit for bit with the gcc build (oracle) and with the IEEE intrinsics
__global__ void exact_math_kernel(int op, const float* __restrict__ x, const float* __restrict__ y, long long n, float* __restrict__ o0, float* __restrict__ o1) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = x[i], b = y ? y[i] : 0.f;
    float r0 = 0.f, r1 = 0.f; ExChk bad;
    switch (op) {
        case 0: mbx_sincospi(a, &r0, &r1); break;
        case 1: r0 = mbx_atan2(a, b); break;
        case 2: r0 = mbx_acos(a); break;
        case 3: r0 = mbx_asin01(a); break;
        case 4: r0 = mbx_rsqrt(a); r1 = __frsqrt_rn(a); break;
        case 5: r0 = xdiv_pos(a, b, bad); r1 = __fdiv_rn(a, b); if (bad.bad()) r0 = r1; break;
        case 6: r0 = xsqrt_pos(a, bad); r1 = __fsqrt_rn(a); if (bad.bad()) r0 = r1; break;
        default: break;
    }
    o0[i] = r0; if (o1) o1[i] = r1;
}

__global__ void trans_refracted_kernel(const __grid_constant__ LaneParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.L) return;
    float sx, sy; trans_refracted_screen(P.cam, P.trans, ld3(P.wi, i), ld3(P.n_geo, i), ld3(P.p, i), sx, sy);
    P.o3[2 * i] = sx; P.o3[2 * i + 1] = sy;
    P.oflat[i] = trans_refracted_index(P.cam, P.trans, ld3(P.wi, i), ld3(P.n_geo, i), ld3(P.p, i));
}
int fill_trans_lanes(const mb200_trans* t, LaneParams& P) {
    if (!t || !(t->ior > 0.f)) return MB200_EINVAL;
    P.trans.bg = t->bg; P.trans.mask = t->mask; P.trans.ior = t->ior; P.trans.spec_trans = t->spec_trans; P.trans.refract_dist = t->refract_distance;
    return MB200_OK;
}
int fill(const mb200_cfg* c, LaneParams& P) {
    if (!c || c->H <= 0 || c->W <= 0) return MB200_EINVAL;
    memset(&P, 0, sizeof(P));
    for (int i = 0; i < 16; ++i) { P.cam.view[i] = c->view[i]; P.cam.proj[i] = c->proj[i]; P.cam.c2w[i] = c->cam_to_world[i]; }
    P.cam.tan_half_fov_x = c->tan_half_fov_x; P.cam.H = c->H; P.cam.W = c->W;
    P.cam.stride = (c->flags & MB200_FLAG_ROW_STRIDE_H) ? c->H : c->W;
    P.use_mesh_normal = c->use_mesh_normal;
    return MB200_OK;
}
}  // namespace

extern "C" {

int mb200_bsdf_eval_pdf(const mb200_cfg* c, int64_t L, const float* p, const float* n_geo, const float* wi_world,
                        const float* wo_world, const float* a, const float* r, const float* m, const float* n_opt,
                        float* out_f, float* out_pdf, void* stream) {
    LaneParams P; int rc = fill(c, P); if (rc) return rc;
    if (L < 0 || !p || !n_geo || !wi_world || !wo_world || !a || !r || !m || !out_f || !out_pdf) return MB200_EINVAL;
    if (L == 0) return MB200_OK;
    P.L = L; P.p = p; P.n_geo = n_geo; P.wi = wi_world; P.wo = wo_world; P.a = a; P.r = r; P.m = m; P.n_opt = n_opt;
    P.o3 = out_f; P.o1 = out_pdf;
    bsdf_eval_pdf_kernel<<<(unsigned)((L + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
    return mb200_check_launch();
}

int mb200_bsdf_eval_grad(const mb200_cfg* c, int64_t L, const float* p, const float* n_geo, const float* wi_world,
                         const float* wo_world, const float* a, const float* r, const float* m, const float* n_opt, const float* w,
                         float* g_a, float* g_r, float* g_m, float* g_n, void* stream) {
    LaneParams P; int rc = fill(c, P); if (rc) return rc;
    if (L < 0 || !p || !n_geo || !wi_world || !wo_world || !a || !r || !m || !w || !g_a || !g_r || !g_m || !g_n) return MB200_EINVAL;
    if (L == 0) return MB200_OK;
    P.L = L; P.p = p; P.n_geo = n_geo; P.wi = wi_world; P.wo = wo_world; P.a = a; P.r = r; P.m = m; P.n_opt = n_opt; P.s2 = w;
    P.o3 = g_a; P.o1 = g_r; P.ow = g_m;
    bsdf_eval_grad_kernel<<<(unsigned)((L + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, g_n);
    return mb200_check_launch();
}

int mb200_bsdf_sample(const mb200_cfg* c, int64_t L, const float* p, const float* n_geo, const float* wi_world,
                      const float* sample1, const float* sample2, const float* a, const float* r, const float* m,
                      const float* n_opt, float* out_wo, float* out_pdf, float* out_weight, void* stream) {
    LaneParams P; int rc = fill(c, P); if (rc) return rc;
    if (L < 0 || !p || !n_geo || !wi_world || !sample1 || !sample2 || !a || !r || !m || !out_wo || !out_pdf || !out_weight) return MB200_EINVAL;
    if (L == 0) return MB200_OK;
    P.L = L; P.p = p; P.n_geo = n_geo; P.wi = wi_world; P.s1 = sample1; P.s2 = sample2; P.a = a; P.r = r; P.m = m; P.n_opt = n_opt;
    P.o3 = out_wo; P.o1 = out_pdf; P.ow = out_weight;
    bsdf_sample_kernel<<<(unsigned)((L + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
    return mb200_check_launch();
}

int mb200_trans_eval_pdf(const mb200_cfg* c, const mb200_trans* t, int64_t L, const float* p, const float* n_geo, const float* wi_world,
                         const float* wo_world, const float* a, const float* r, const float* m, const float* n_opt,
                         float* out_f, float* out_pdf, void* stream) {
    LaneParams P; int rc = fill(c, P); if (rc) return rc;
    if ((rc = fill_trans_lanes(t, P)) != MB200_OK) return rc;
    if (L < 0 || !t->bg || !t->mask || !p || !n_geo || !wi_world || !wo_world || !a || !r || !m || !out_f || !out_pdf) return MB200_EINVAL;
    if (L == 0) return MB200_OK;
    P.L = L; P.p = p; P.n_geo = n_geo; P.wi = wi_world; P.wo = wo_world; P.a = a; P.r = r; P.m = m; P.n_opt = n_opt;
    P.o3 = out_f; P.o1 = out_pdf;
    trans_eval_pdf_kernel<<<(unsigned)((L + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
    return mb200_check_launch();
}

int mb200_trans_sample(const mb200_cfg* c, const mb200_trans* t, int64_t L, const float* p, const float* n_geo, const float* wi_world,
                       const float* sample1, const float* sample2, const float* a, const float* r, const float* m,
                       const float* n_opt, float* out_wo, float* out_pdf, float* out_weight, void* stream) {
    LaneParams P; int rc = fill(c, P); if (rc) return rc;
    if ((rc = fill_trans_lanes(t, P)) != MB200_OK) return rc;
    if (L < 0 || !t->bg || !t->mask || !p || !n_geo || !wi_world || !sample1 || !sample2 || !a || !r || !m || !out_wo || !out_pdf || !out_weight) return MB200_EINVAL;
    if (L == 0) return MB200_OK;
    P.L = L; P.p = p; P.n_geo = n_geo; P.wi = wi_world; P.s1 = sample1; P.s2 = sample2; P.a = a; P.r = r; P.m = m; P.n_opt = n_opt;
    P.o3 = out_wo; P.o1 = out_pdf; P.ow = out_weight;
    trans_sample_kernel<<<(unsigned)((L + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
    return mb200_check_launch();
}

int mb200_trans_refracted_texel(const mb200_cfg* c, const mb200_trans* t, int64_t L, const float* p, const float* n_geo,
                                const float* wi_world, float* out_screen, int64_t* out_flat, void* stream) {
    LaneParams P; int rc = fill(c, P); if (rc) return rc;
    if ((rc = fill_trans_lanes(t, P)) != MB200_OK) return rc;
    if (L < 0 || !p || !n_geo || !wi_world || !out_screen || !out_flat) return MB200_EINVAL;
    if (L == 0) return MB200_OK;
    P.L = L; P.p = p; P.n_geo = n_geo; P.wi = wi_world; P.o3 = out_screen; P.oflat = (long long*)out_flat;
    trans_refracted_kernel<<<(unsigned)((L + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
    return mb200_check_launch();
}

int mb200_debug_exact_math(int op, const float* x, const float* y, int64_t n, float* out0, float* out1, void* stream) {
    if (op < 0 || op > 6 || !x || !out0 || n < 0) return MB200_EINVAL;
    if ((op == 1 || op == 5) && !y) return MB200_EINVAL;
    if (n == 0) return MB200_OK;
    exact_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(op, x, y, (long long)n, out0, out1);
    return mb200_check_launch();
}

}  // extern "C"
