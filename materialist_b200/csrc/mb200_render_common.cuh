// mb200_render_common.cuh — kernel-parameter block and host helpers shared by the G-buffer shade kernels
// (mb200_render.cu) and the mesh-mode path-tracing kernels (mb200_mesh.cu).
#pragma once
#include <stdlib.h>
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
// (measured and rejected, profiles/r5m: row weight folded into the radiance by the staging lane — records of 5 x (L.rgb * wy_j, wy_j) +
// wx[5], 6 instead of 8 instructions per record in the tap loop — shade_fwd 1.166 -> 1.202 ms: 7 instead of 4 16-byte stores per
// sample and more spills at the 64-register cap outweigh the shorter loop; a 28-float record stride alone costs +0.65 %)
// (measured and rejected, profiles/r5c: weights staged as (w,w) pairs + FMUL2 / 2 x FFMA2 per record in the forward tap loop — 6
// instead of 8 instructions per record — made shade_fwd SLOWER, 1.205 -> 1.270 ms: the records grow 20 -> 28 floats, 8.7 KB more
// shared memory per CTA taken from the L1 that serves the two un-staged pyramid levels and the envmap texels)

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
    int env_smem_texels;                 // > 0: the float4 texels are staged in shared memory too (small envmaps)
    int env_grad_smem;                   // adjoint: 1 = the envmap-gradient map is accumulated in shared memory (3 floats / texel) and flushed per CTA
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

// sums x over the lanes of `peers` (all 32 lanes of the warp must call); result valid in the lowest lane of each group
template <int N>
__device__ __forceinline__ void reduce_peers(unsigned peers, float (&x)[N]) {
    const int lane = threadIdx.x & 31;
    int rel_pos = __popc(peers << (32 - lane));
    if (lane == 0) rel_pos = 0;
    peers &= (0xfffffffeu << lane);
    while (__any_sync(0xffffffffu, peers)) {
        const int next = __ffs(peers);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const float t = __shfl_sync(0xffffffffu, x[i], (next - 1) & 31);
            if (next) x[i] += t;
        }
        const unsigned done = rel_pos & 1;
        peers &= ~__ballot_sync(0xffffffffu, done);
        rel_pos >>= 1;
    }
}

// Envmap-gradient scatter of one bilinear footprint per lane, WARP-AGGREGATED (all 32 lanes call; `active` = this lane has an
// update): lanes that hit the same envmap cell (match.any on the cell index) first add their 4 x rgb contributions in registers,
// and only the group's lowest lane issues the updates — into the CTA's shared-memory copy of a small gradient map (`senv`, 3 floats
// per texel, flushed once per CTA: block-privatised) or, for maps that do not fit, as four 16-byte red.global.add.v4.f32 into the
// CTA's L2-resident slab.  With a sun texel half of a warp's emitter samples share a few cells: one update instead of ~16.
__device__ __forceinline__ void env_scatter_agg(float4* g, float* senv, int Wi, const Bilerp& b, float3 cot, bool active) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned key = active ? b.i00 : (0xffffffe0u | lane);          // inactive lanes: singleton groups
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const bool leader = (peers & ((1u << lane) - 1u)) == 0u;
    float x[12];
    if (active) {
        const float w00 = b.w0y * b.w0x, w10 = b.w0y * b.w1x, w01 = b.w1y * b.w0x, w11 = b.w1y * b.w1x;
        x[0] = w00 * cot.x; x[1] = w00 * cot.y; x[2] = w00 * cot.z; x[3] = w10 * cot.x; x[4] = w10 * cot.y; x[5] = w10 * cot.z;
        x[6] = w01 * cot.x; x[7] = w01 * cot.y; x[8] = w01 * cot.z; x[9] = w11 * cot.x; x[10] = w11 * cot.y; x[11] = w11 * cot.z;
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = 0.f;
    }
    if (__any_sync(0xffffffffu, peers != (1u << lane))) reduce_peers(peers, x);
    if (active && leader) {
        if (senv) {
            float* t = senv + 3 * b.i00;
            atomicAdd(t, x[0]); atomicAdd(t + 1, x[1]); atomicAdd(t + 2, x[2]); atomicAdd(t + 3, x[3]); atomicAdd(t + 4, x[4]); atomicAdd(t + 5, x[5]);
            t += 3 * Wi;
            atomicAdd(t, x[6]); atomicAdd(t + 1, x[7]); atomicAdd(t + 2, x[8]); atomicAdd(t + 3, x[9]); atomicAdd(t + 4, x[10]); atomicAdd(t + 5, x[11]);
        } else {
            atomicAdd(g + b.i00,          make_float4(x[0], x[1], x[2], 0.f));
            atomicAdd(g + b.i00 + 1,      make_float4(x[3], x[4], x[5], 0.f));
            atomicAdd(g + b.i00 + Wi,     make_float4(x[6], x[7], x[8], 0.f));
            atomicAdd(g + b.i00 + Wi + 1, make_float4(x[9], x[10], x[11], 0.f));
        }
    }
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
    P.hier.smem_from = d->n_levels; P.hier.smem_off0 = d->total_floats;      // no shared-memory staging unless plan_env_staging() says so
    P.env.tex = reinterpret_cast<const float4*>(env4); P.env.Wi = d->res_x; P.env.He = d->res_y; P.env.u_shift = c->env_u_shift;
    P.gpos = reinterpret_cast<const float4*>(gpos); P.gnrm = reinterpret_cast<const float4*>(gnrm);
    P.a = a; P.r = r; P.m = m; P.n_opt = n_opt;
    P.H = c->H; P.W = c->W; P.spp = c->spp; P.seed = c->seed; P.flags = c->flags; P.use_mesh_normal = c->use_mesh_normal;
    P.max_depth = c->max_depth;
    return MB200_OK;
}

// Shared-memory staging plan for a kernel with `budget` bytes of dynamic shared memory: the deepest suffix of pyramid levels
// (smallest first: levels n-1, n-2, ...) that fits, then the texels if everything fits.  Returns the dynamic bytes to launch with.
inline size_t plan_env_staging(RenderParams& P, const mb200_hier_desc* d, size_t budget) {
    P.hier.smem_from = d->n_levels; P.hier.smem_off0 = d->total_floats; P.hier.smem_floats = 0; P.env_smem_texels = 0;
    for (int l = d->n_levels - 1; l >= 0; --l) {
        const size_t bytes = sizeof(float) * (size_t)(d->total_floats - d->lvl_off[l]);
        if (bytes > budget) break;
        P.hier.smem_from = l; P.hier.smem_off0 = d->lvl_off[l]; P.hier.smem_floats = d->total_floats - d->lvl_off[l];
    }
    for (int l = 0; l < d->n_levels; ++l) P.hier.lvl_sm[l] = make_int2(4 * (d->lvl_off[l] - P.hier.smem_off0), 4 * d->lvl_w[l]);
    size_t used = sizeof(float) * (size_t)((P.hier.smem_floats + 3) & ~3);
    const size_t tex = sizeof(float4) * (size_t)d->res_x * d->res_y;
    if (P.hier.smem_from == 0 && used + tex <= budget) { P.env_smem_texels = d->res_x * d->res_y; used += tex; }
    return used;
}
// dynamic shared-memory budget of the staging (bytes per CTA); MB200_HIER_SMEM_KB overrides (0 = staging off) for measurements
inline size_t env_staging_budget(size_t dflt) {
    static long kb = -2;
    if (kb == -2) { const char* e = getenv("MB200_HIER_SMEM_KB"); kb = e ? atol(e) : -1; }
    return kb >= 0 ? (size_t)kb * 1024 : dflt;
}

// MB200_ENV_SCATTER: "direct" (default) | "agg" (warp-aggregated) | "smem" (warp-aggregated + block-privatised in shared memory)
inline int env_scatter_aggregated() {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("MB200_ENV_SCATTER"); mode = !e ? 0 : (!strcmp(e, "agg") ? 1 : (!strcmp(e, "smem") ? 2 : 0)); }
    return mode;
}
// MB200_GRID_CTAS_PER_SM: CTAs per SM of the persistent-style grids (default 32; 0 = exactly the resident count, one wave)
inline int grid_ctas_per_sm() {
    static int n = -1;
    if (n < 0) { const char* e = getenv("MB200_GRID_CTAS_PER_SM"); n = e ? atoi(e) : 32; }
    return n;
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
// Persistent-style grid: SMs x 32 CTAs (the pixel loop strides over the grid, so any size is correct; each CTA stages the envmap
// pyramid once and then shades ~55 pixels at C2).  Measured on the C2 step (profiles/r4c_grid_smem_ab.log, r4d_*): exactly ONE
// wave (SMs x resident CTAs; MB200_GRID_CTAS_PER_SM=0) 3.12 ms, x8 3.05, x16 2.98, x32 2.92, x64 2.92, x128 2.96, x256 3.05.
// One static wave ends with its slowest CTA (the SMs do not run at identical speed); with many short CTAs the hardware scheduler
// refills an SM the moment a CTA retires, and the partial last wave is short.
template <typename K>
inline int persistent_grid(K kernel, int npix, size_t dyn_smem = 0) {
    static const void* keys[32]; static int vals[32]; static int n_keys = 0;     // (per function-pointer TYPE; keyed by the pointer)
    int resident = 0;
    for (int i = 0; i < n_keys; ++i) if (keys[i] == (const void*)kernel) { resident = vals[i]; break; }
    if (resident <= 0) {
        // static + dynamic shared memory may pass 48 KB (adjoint, 8 lanes per pixel: 12.5 KB of film cotangents + 44 KB of staging)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, dyn_smem) != cudaSuccess || n <= 0) n = 2;
        resident = n;
        if (n_keys < 32) { keys[n_keys] = (const void*)kernel; vals[n_keys] = n; ++n_keys; }
    }
    const int blocks_needed = (npix + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int per_sm = grid_ctas_per_sm();
    const int cap = mb200_sm_count() * (per_sm > 0 ? per_sm : resident);
    return blocks_needed < cap ? blocks_needed : cap;
}

}  // namespace
