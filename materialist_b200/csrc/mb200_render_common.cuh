// mb200_render_common.cuh — kernel-parameter block and host helpers shared by the G-buffer shade kernels
// (mb200_render.cu) and the mesh-mode path-tracing kernels (mb200_mesh.cu).
#pragma once
#include <string.h>
#include "mb200_device.cuh"
#include "mb200_host.h"

using namespace mb;

namespace {

// resident CTAs per SM the shade kernels are compiled for (register cap = 65536 / (256 * N)); tuned in profiles/
#ifndef MB_MIN_BLOCKS_FWD
#define MB_MIN_BLOCKS_FWD 4
#endif
#ifndef MB_MIN_BLOCKS_BWD
#define MB_MIN_BLOCKS_BWD 3
#endif
constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kRecStride = 20;            // floats per staged sample record (bank-conflict-free for 16B stores)

struct RenderParams {
    CamView cam; HierView hier; EnvView env;
    const float4* gpos; const float4* gnrm;
    const float* a; const float* r; const float* m; const float* n_opt;
    int H, W, spp; uint32_t seed; int flags; int use_mesh_normal; int max_depth;
    int prow0, prows;                    // rows this launch processes (shard rows + film halo)
    // forward
    float* partials;
    // adjoint
    const float4* gadj; int grow0, grows; // G image rows
    float* g_a; float* g_r; float* g_m; float* g_n; float4* g_env4; int env_slabs; long long env_slab_stride;
    TransView trans;                     // TransBSDF kernels only (mb200_trans_*)
};

inline int fill_trans(const mb200_trans* t, RenderParams& P) {
    if (!t || !t->bg || !t->mask || !(t->ior > 0.f)) return MB200_EINVAL;
    P.trans.bg = t->bg; P.trans.mask = t->mask; P.trans.ior = t->ior; P.trans.spec_trans = t->spec_trans; P.trans.refract_dist = t->refract_distance;
    return MB200_OK;
}

__device__ __forceinline__ void env_scatter(float4* g, int Wi, const Bilerp& b, float3 cot) {
    const float w00 = b.w0y * b.w0x, w10 = b.w0y * b.w1x, w01 = b.w1y * b.w0x, w11 = b.w1y * b.w1x;
    atomicAdd(g + b.i00,          make_float4(w00 * cot.x, w00 * cot.y, w00 * cot.z, 0.f));
    atomicAdd(g + b.i00 + 1,      make_float4(w10 * cot.x, w10 * cot.y, w10 * cot.z, 0.f));
    atomicAdd(g + b.i00 + Wi,     make_float4(w01 * cot.x, w01 * cot.y, w01 * cot.z, 0.f));
    atomicAdd(g + b.i00 + Wi + 1, make_float4(w11 * cot.x, w11 * cot.y, w11 * cot.z, 0.f));
}

// ---------------------------------------------------------------- host side
// need_gbuf = false: mesh mode (gpos / gnrm unused, and n_opt may be NULL even without use_mesh_normal: geometric normal)
inline int fill_params(const mb200_cfg* c, const float* gpos, const float* gnrm, const float* a, const float* r, const float* m,
                       const float* n_opt, const float* env4, const float* hier, const mb200_hier_desc* d, RenderParams& P,
                       bool need_gbuf = true) {
    if (!c || !a || !r || !m || !env4 || !hier || !d) return MB200_EINVAL;
    if (need_gbuf && (!gpos || !gnrm)) return MB200_EINVAL;
    if (c->H <= 0 || c->W <= 0 || c->spp <= 0 || c->rows <= 0 || c->row0 < 0 || c->row0 + c->rows > c->H) return MB200_EINVAL;
    if (c->filter != MB200_FILTER_BOX && c->filter != MB200_FILTER_GAUSSIAN) return MB200_EINVAL;
    if (need_gbuf && !c->use_mesh_normal && !n_opt) return MB200_EINVAL;
    if ((double)c->H * (double)c->W * (double)c->spp >= 4294967296.0) return MB200_ERANGE;
    if (d->n_levels < 2 || d->n_levels > MB200_MAX_LEVELS) return MB200_EINVAL;
    memset(&P, 0, sizeof(P));
    for (int i = 0; i < 16; ++i) { P.cam.view[i] = c->view[i]; P.cam.proj[i] = c->proj[i]; P.cam.c2w[i] = c->cam_to_world[i]; }
    P.cam.tan_half_fov_x = c->tan_half_fov_x; P.cam.H = c->H; P.cam.W = c->W;
    P.cam.stride = (c->flags & MB200_FLAG_ROW_STRIDE_H) ? c->H : c->W;
    P.hier.data = hier; P.hier.res_x = d->res_x; P.hier.res_y = d->res_y; P.hier.n_levels = d->n_levels;
    P.hier.psx = 1.f / (float)(d->res_x - 1); P.hier.psy = 1.f / (float)(d->res_y - 1);
    for (int l = 0; l < d->n_levels; ++l) { P.hier.lvl_off[l] = d->lvl_off[l]; P.hier.lvl_w[l] = d->lvl_w[l]; }
    P.env.tex = reinterpret_cast<const float4*>(env4); P.env.Wi = d->res_x; P.env.He = d->res_y; P.env.u_shift = c->env_u_shift;
    P.gpos = reinterpret_cast<const float4*>(gpos); P.gnrm = reinterpret_cast<const float4*>(gnrm);
    P.a = a; P.r = r; P.m = m; P.n_opt = n_opt;
    P.H = c->H; P.W = c->W; P.spp = c->spp; P.seed = c->seed; P.flags = c->flags; P.use_mesh_normal = c->use_mesh_normal;
    P.max_depth = c->max_depth;
    return MB200_OK;
}

inline void halo_rows(const mb200_cfg* c, int halo, int* first, int* count) {
    int r0 = c->row0 - halo, r1 = c->row0 + c->rows + halo;
    if (r0 < 0) r0 = 0; if (r1 > c->H) r1 = c->H;
    *first = r0; *count = r1 - r0;
}

inline int grid_for(int npix) {
    const int blocks_needed = (npix + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int cap = mb200_sm_count() * 8;           // persistent-style grid: a multiple of the SM count
    return blocks_needed < cap ? blocks_needed : cap;
}

}  // namespace
