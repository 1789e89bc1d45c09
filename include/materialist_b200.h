/*
 * materialist_b200.h — C-ABI of the B200-native differentiable envmap-shading path.
 *
 * This is the drop-in boundary.  Every entry point takes plain
 * pointers and sizes; there are no torch / C++ types in any signature.  All
 * pointers are DEVICE pointers unless the parameter name ends in `_host`.
 * Every function returns MB200_OK (0) or a negative MB200_E* code, never
 * throws, never allocates device memory and enqueues its work on the stream
 * passed as `void* stream` (a cudaStream_t; NULL = legacy default stream).
 *
 * What each entry point replaces in the reference (lez-s/Materialist @ bdd146f):
 *
 *   mb200_env_prepare      Mitsuba EnvironmentMapEmitter::parameters_changed('data')
 *                          reached from  inverse_img_w_mi.py:63-64  (params['emitter.data']=...;
 *                          params.update())  and the envmap file load at inverse_img_w_mi.py:54,
 *                          render_final.py:52.  [upstream mitsuba==3.5.2 src/emitters/envmap.cpp,
 *                          include/mitsuba/core/distr_2d.h Hierarchical2D]
 *   mb200_shade_fwd        mi.render(scene, params, spp, seed) forward —
 *                          inverse_img_w_mi.py:65, :79; render_final.py:194, :227, :391 —
 *                          with the BSDF virtual calls MatDiffBSDF.eval_pdf / .sample
 *                          (myutils/mi_plugin.py:1429-1460) fused in.
 *   mb200_film_develop     HDRFilm::develop (rgb / weight) at the end of mi.render.
 *   mb200_film_weights     the ImageBlock::put weight channel of the seed_grad re-render
 *                          (Integrator::render_backward re-renders).
 *   mb200_shade_bwd        the loss.backward() leg of  render_w_brdf / render_envmap
 *                          (inverse_img_w_mi.py:59-80, :248, :419): Mitsuba
 *                          render_backward(seed_grad) + dr.backward, returning
 *                          d/d{a,r,m,n} and d/d{emitter.data}.
 *   mb200_env_grad_finish  adjoint of the column-average / appended-column map of
 *                          parameters_changed.
 *   mb200_bsdf_eval_pdf    MatDiffBSDF.eval_pdf  (myutils/mi_plugin.py:1449-1460) on an array of lanes.
 *   mb200_bsdf_sample      MatDiffBSDF.sample    (myutils/mi_plugin.py:1429-1446) on an array of lanes.
 *   mb200_trans_*          TransBSDF (myutils/mi_plugin.py:1477-1770), the transparency-editing plugin of
 *                          trans_edit.py:16-49: forward render in G-buffer and mesh mode + lane-level eval_pdf / sample.
 *   mb200_debug_sample_indices
 *                          not in the reference: exposes the integer decisions of the path
 *                          (hierarchy offsets, texel index, lobe) so tests can assert them bit-exact.
 *   mb200_mesh_build       the scene side of mi.load_dict({... 'shape': {'type': 'ply', 'filename': mesh_path}})
 *                          inverse_img_w_mi.py:40-56 / render_final.py:32-53: Mitsuba's PLY loader (vertex normals
 *                          recomputed, angle-weighted) + the OptiX acceleration structure, here a GPU-built
 *                          Morton-ordered implicit 4-ary BVH.
 *   mb200_mesh_shade_fwd / _bwd
 *                          the same mi.render / render_backward calls as mb200_shade_fwd / _bwd, but tracing the
 *                          triangle mesh the reference really renders: per-sample primary hits, shadow rays, up to
 *                          max_depth-1 bounces with the material fetched at every vertex through
 *                          mi_world_to_screen(si.p) (mi_plugin.py:1435,1456).  Same film kernels either way.
 *   mb200_mesh_primary     primary visibility of the mesh -> the G-buffer mb200_shade_* consume (the fast path);
 *                          also exposes the primary triangle / texel indices so tests can assert them bit-exact.
 *   mb200_mesh_intersect   debug: closest / any hit for caller-supplied rays (bit-exact vs the oracle).
 *   mb200_posmlp_*         PosMLP.forward (+autograd backward) mymodels/mlps.py:211-251 as used at
 *                          inverse_img_w_mi.py:163,:493 (brdf_net) and :117,:238 (envmap_net).
 *   mb200_cdf_build        build_envmap   myutils/envmap_utils.py:43-66
 *   mb200_cdf_sample       sample_envmap  myutils/envmap_utils.py:172-201
 *   mb200_sh_project       computeSHFromImage-style order-4 projection  myutils/computeSH.py:299-347
 *                          (deterministic sample positions supplied by the caller)
 *   mb200_sh_reconstruct   reconstImageFromSH  myutils/computeSH.py:226-240
 */
#ifndef MATERIALIST_B200_H
#define MATERIALIST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- errors */
#define MB200_OK             0
#define MB200_EINVAL        -1   /* bad argument (null pointer, non-positive size, ...) */
#define MB200_ERANGE        -2   /* size out of supported range (e.g. H*W*spp >= 2^32)   */
#define MB200_ELAUNCH       -3   /* CUDA launch / runtime error (see mb200_last_cuda_error) */
#define MB200_EUNSUPPORTED  -4   /* valid request, not implemented by this build */
#define MB200_EIO           -5   /* a file could not be written completely (fwrite / fclose failed) */

/* ---------------------------------------------------------------- flags  */
/* cfg.flags — reference-exact quirks, all ON by default in the Python host. */
#define MB200_FLAG_WO_WORLD_QUIRK   1  /* bs.wo returned in world space is pushed through si.to_world
                                          again by the integrator (mi_plugin.py:1444) */
#define MB200_FLAG_ROW_STRIDE_H     2  /* texel index = x + y*shape[0] (mi_plugin.py:1310,1381);
                                          off: x + y*W (the intended behaviour for W != H) */
#define MB200_FLAG_ENV_HALF_TEXEL   4  /* u -= .5/(res.x-1) in envmap eval / pdf, += in sample
                                          (mitsuba 3.5 envmap.cpp) */
#define MB200_FLAG_AD_WEIGHTS       8  /* use the AD-pass BSDF-sample weight f2/p2
                                          instead of the primal f/(p+1e-6); set by shade_bwd itself,
                                          may be set on shade_fwd to render the primal of the AD pass */

#define MB200_FILTER_BOX       0
#define MB200_FILTER_GAUSSIAN  1

/* envmap ingest modes for mb200_env_prepare */
#define MB200_ENV_ASSIGNED  0  /* tensor assigned through params['emitter.data']: width kept,
                                  first/last column averaged (Appendix A5) */
#define MB200_ENV_FILE      1  /* bitmap loaded from file: one column appended = copy of col 0 */

#define MB200_MAX_LEVELS 24
#define MB200_FILM_TAPS  25   /* 5x5 footprint of the radius-2 gaussian */

/* ---------------------------------------------------------------- types  */
typedef struct mb200_cfg {
    int32_t  H, W;            /* full image size (pixels)                                   */
    int32_t  spp;             /* samples per pixel                                          */
    int32_t  max_depth;       /* path integrator max_depth (>= 2)            */
    uint32_t seed;            /* sampler seed (mi.render seed, or seed_grad for the adjoint) */
    int32_t  filter;          /* MB200_FILTER_*                                             */
    int32_t  flags;           /* MB200_FLAG_*                                               */
    int32_t  use_mesh_normal; /* 1: shade with G-buffer normal, 0: with the n map           */
    int32_t  row0, rows;      /* pixel shard: image rows [row0, row0+rows)                  */
    float    view[16];        /* world -> camera, row-major (MatDiffBSDF.view_matrix)       */
    float    proj[16];        /* camera -> clip, row-major (MatDiffBSDF.persp_proj_matx)    */
    float    cam_to_world[16];/* sensor to_world, row-major                                 */
    float    tan_half_fov_x;  /* tan(x_fov/2) of the perspective sensor                     */
    float    env_u_shift;     /* .5/(res.x-1) when MB200_FLAG_ENV_HALF_TEXEL else 0 (filled by host) */
} mb200_cfg;

/* Layout of the Hierarchical2D sampling pyramid inside one float buffer.
 * level 0 = vertex values (res_x * res_y, row-major, normalised);
 * level l>=1 = patch sums, 2x2-swizzled, size lvl_w[l] x lvl_h[l] (even). */
typedef struct mb200_hier_desc {
    int32_t res_x, res_y;                 /* internal envmap resolution (vertices) */
    int32_t n_levels;                     /* number of levels incl. level 0        */
    int32_t lvl_off[MB200_MAX_LEVELS];    /* offset (floats) of each level         */
    int32_t lvl_w[MB200_MAX_LEVELS];
    int32_t lvl_h[MB200_MAX_LEVELS];
    int32_t total_floats;
} mb200_hier_desc;

/* ---------------------------------------------------------------- host helpers (no GPU needed) */
const char* mb200_strerror(int code);
const char* mb200_last_cuda_error(void);            /* text of the last CUDA error seen by this thread */
int  mb200_version(void);
/* internal width for a user envmap of width We under `mode` */
int  mb200_env_internal_width(int We, int mode);
/* fills `out` for an internal resolution (res_x, res_y) */
int  mb200_hier_describe(int res_x, int res_y, mb200_hier_desc* out_host);
/* bytes of scratch mb200_env_prepare needs */
size_t mb200_env_scratch_bytes(int res_x, int res_y);
/* number of rows (incl. film halo) shade_fwd writes partials for, and the first such row */
int  mb200_fwd_partial_rows(const mb200_cfg* cfg_host, int* first_row_host);
/* floats per pixel in the `partials` buffer for cfg.filter */
int  mb200_partial_stride(int filter);

/* ---------------------------------------------------------------- envmap */
/* env_in  : (He, We, 3) fp32 user envmap
 * env4    : (He, Wi, 4) fp32 internal texels (rgb, 0), Wi = mb200_env_internal_width
 * hier    : desc.total_floats fp32
 * scratch : mb200_env_scratch_bytes bytes (8-byte aligned) */
int mb200_env_prepare(const float* env_in, int He, int We, int mode,
                      float* env4, float* hier, const mb200_hier_desc* desc_host,
                      void* scratch, void* stream);

/* Number of privatised copies ("slabs") of the envmap-gradient grid mb200_shade_bwd should scatter into: the CTAs of the
 * adjoint kernel spread over them (slab = blockIdx % n) so that hot texels (a sun) are n L2 addresses instead of one.
 * Sized to keep n * He * Wi * 16 bytes <= 64 MB (L2-resident). */
int mb200_env_grad_slabs(int He, int We, int mode);
/* g_env4 (n_slabs, He, Wi, 4) -> g_env (He, We, 3): sums the slabs in a fixed order (in place: slab 0 of g_env4 receives the sum,
 * g_env4 is scratch the caller re-zeroes before the next adjoint), then the adjoint of the ingest map; g_env is OVERWRITTEN */
int mb200_env_grad_finish(float* g_env4, int n_slabs, int He, int We, int mode, float* g_env, void* stream);

/* ---------------------------------------------------------------- render */
/* G-buffer: gpos (H,W,4) = (x,y,z,valid?1:0), gnrm (H,W,4) = (nx,ny,nz,0)  — full image, fp32.
 * a (H,W,3), r (H,W,1), m (H,W,1), n_opt (H,W,3) or NULL — full image.
 * partials : (prow_count, W, stride) with stride = mb200_partial_stride(filter),
 *            prow_count / first row from mb200_fwd_partial_rows. */
int mb200_shade_fwd(const mb200_cfg* cfg_host,
                    const float* gpos, const float* gnrm,
                    const float* a, const float* r, const float* m, const float* n_opt,
                    const float* env4, const float* hier, const mb200_hier_desc* desc_host,
                    float* partials, void* stream);

/* partials -> img (rows, W, 3) for the shard rows; */
int mb200_film_develop(const mb200_cfg* cfg_host, const float* partials, float* img, void* stream);

/* Weight channel only (uses nothing but the RNG): wpart (prow_count, W, 25) for gaussian.
 * Then G[q] = grad_img[q] / W_q is formed by mb200_film_adjoint for rows [row0-2, row0+rows+2) ∩ image:
 * grad_img_halo : (grow_count, W, 3) incl. the 2-row halo (rows outside the image absent),
 * gadj          : (grow_count, W, 4) = (G.rgb, 0). For the box filter wpart may be NULL and gadj = grad/spp. */
int mb200_film_weights(const mb200_cfg* cfg_host, float* wpart, void* stream);
int mb200_film_adjoint(const mb200_cfg* cfg_host, const float* wpart, const float* grad_img_halo,
                       float* gadj, void* stream);
/* rows of wpart (needs a 4-row halo) and of gadj / grad_img_halo (2-row halo) */
int mb200_bwd_wpart_rows(const mb200_cfg* cfg_host, int* first_row_host);
int mb200_bwd_gadj_rows(const mb200_cfg* cfg_host, int* first_row_host);

/* Adjoint render (cfg.seed = seed_grad).  Accumulates (+=) into the gradient buffers:
 * g_a (H,W,3), g_r (H,W,1), g_m (H,W,1), g_n (H,W,3) full-image, any may be NULL;
 * g_env4 (n_env_slabs, He, Wi, 4) or NULL (n_env_slabs from mb200_env_grad_slabs, >= 1).  Caller zeroes them. */
int mb200_shade_bwd(const mb200_cfg* cfg_host,
                    const float* gpos, const float* gnrm,
                    const float* a, const float* r, const float* m, const float* n_opt,
                    const float* env4, const float* hier, const mb200_hier_desc* desc_host,
                    const float* gadj,
                    float* g_a, float* g_r, float* g_m, float* g_n, float* g_env4, int n_env_slabs,
                    void* stream);

/* (S,4) int32 per lane: (hier off.x, off.y, texel flat index, lobe 0=specular 1=diffuse); shard rows only */
int mb200_debug_sample_indices(const mb200_cfg* cfg_host, const float* gpos, const float* r,
                               const float* hier, const mb200_hier_desc* desc_host,
                               int32_t* out, void* stream);

/* Per-lane decision record of the PRODUCTION forward sample function (the kernels' own shade_sample, instrumented):
 * (S,12) int32 per lane = hier off.x, off.y, texel flat index (-1: no surface), lobe, envmap cell (flat index of the
 * upper-left texel) of the emitter sample, envmap cell of the BSDF-sampled direction (of the primary ray for pixels
 * without a surface), IEEE bits of the emitter direction (3) and of the BSDF-sampled direction (3).  out_radiance
 * (S,3) float or NULL: the lane's radiance.  cfg.flags & MB200_FLAG_AD_WEIGHTS selects the AD-pass weights.
 * north_star "CDF/sample indices bit-exact": every word is compared for equality with the oracle's. */
int mb200_debug_sample_record(const mb200_cfg* cfg_host, const float* gpos, const float* gnrm,
                              const float* a, const float* r, const float* m, const float* n_opt,
                              const float* env4, const float* hier, const mb200_hier_desc* desc_host,
                              int32_t* out, float* out_radiance, void* stream);

/* ---------------------------------------------------------------- mesh mode (triangle mesh instead of a G-buffer) */
#ifndef MB200_MESH_LEAF
#define MB200_MESH_LEAF        1    /* triangles per BVH leaf (compile-time; mb200_mesh_describe reports the layout) */
#endif
#define MB200_MESH_MAX_LEVELS  16
/* Layout of a built mesh inside ONE device buffer `mesh_buf` (byte offsets), filled by mb200_mesh_describe:
 *   header : centre.xyz, radius of the scene bounding sphere, bbox lo.xyz, hi.xyz (written by the build, read by kernels)
 *   tv     : (n_slots, 3) float4 — triangles in Morton order: (p0.xyz, triangle id as int bits), (p1.xyz,0), (p2.xyz,0)
 *   tn     : (n_slots, 3) float4 — their three vertex normals (absent when face_normals)
 *   nodes  : n_groups x 6 float4 — implicit 4-ary tree: group g of level l holds the boxes of nodes 4g..4g+3 of level l
 *            as (lo.x[4], lo.y[4], lo.z[4], hi.x[4], hi.y[4], hi.z[4]); node i of level l has children = group i of
 *            level l-1; level 0 nodes are leaves of MB200_MESH_LEAF consecutive slots. */
typedef struct mb200_mesh_desc {
    int32_t nv, nt;
    int32_t face_normals;      /* 1: flat shading frames; 0: Mitsuba's default for a PLY without normals (angle-weighted vertex normals) */
    int32_t n_slots, n_levels, n_groups;
    int32_t lvl_nodes[MB200_MESH_MAX_LEVELS];
    int32_t lvl_group_off[MB200_MESH_MAX_LEVELS];
    int64_t off_header, off_tv, off_tn, off_nodes, total_bytes;
} mb200_mesh_desc;
int    mb200_mesh_describe(int nv, int nt, int face_normals, mb200_mesh_desc* out_host);        /* host only */
size_t mb200_mesh_scratch_bytes(int nv, int nt, int face_normals);                             /* needs a CUDA device; 0 on error */
/* verts (nv,3) fp32, tris (nt,3) int32 (device) -> mesh_buf (desc.total_bytes, 256-byte aligned); scratch may be freed
 * once the stream has passed the call.  Everything runs on the GPU (bounds, Morton sort, vertex normals, boxes). */
int mb200_mesh_build(const float* verts, const int32_t* tris, const mb200_mesh_desc* desc_host,
                     void* mesh_buf, void* scratch, void* stream);
/* as mb200_shade_fwd / mb200_shade_bwd with the mesh in place of the G-buffer; n_opt may be NULL (geometric normal).
 * cfg.max_depth <= 8.  Film buffers (partials, gadj) and mb200_film_* are shared with the G-buffer path. */
int mb200_mesh_shade_fwd(const mb200_cfg* cfg_host, const mb200_mesh_desc* desc_host, const void* mesh_buf,
                         const float* a, const float* r, const float* m, const float* n_opt,
                         const float* env4, const float* hier, const mb200_hier_desc* hdesc_host,
                         float* partials, void* stream);
int mb200_mesh_shade_bwd(const mb200_cfg* cfg_host, const mb200_mesh_desc* desc_host, const void* mesh_buf,
                         const float* a, const float* r, const float* m, const float* n_opt,
                         const float* env4, const float* hier, const mb200_hier_desc* hdesc_host,
                         const float* gadj,
                         float* g_a, float* g_r, float* g_m, float* g_n, float* g_env4, int n_env_slabs,
                         void* stream);
/* closest hit (out_tri = triangle id or -1, out_tuv (n,3) = t,u,v) or any hit (out_tri = 1/0) for n rays (o, d: (n,3));
 * maxt (n) or NULL = infinity */
int mb200_mesh_intersect(const mb200_mesh_desc* desc_host, const void* mesh_buf, const float* o, const float* d,
                         const float* maxt, int n, int any_hit, int32_t* out_tri, float* out_tuv, void* stream);
/* primary visibility through film offset (jx, jy) of every pixel: gpos (H,W,4) = (p, hit?1:0), gnrm (H,W,4) = (n_geo, 0),
 * tri (H,W) int32 or NULL, flat (H,W) int32 texel index of the hit point or NULL */
int mb200_mesh_primary(const mb200_cfg* cfg_host, const mb200_mesh_desc* desc_host, const void* mesh_buf, float jx, float jy,
                       float* gpos, float* gnrm, int32_t* tri, int32_t* flat, void* stream);

/* ---------------------------------------------------------------- BSDF plugin on lanes */
/* All lane arrays are (L,3) / (L) fp32.  cfg supplies view/proj/H/W/flags/use_mesh_normal.
 * eval_pdf: wo_world = light direction, wi_world = view direction (mi_plugin.py:1449-1460). */
int mb200_bsdf_eval_pdf(const mb200_cfg* cfg_host, int64_t L,
                        const float* p, const float* n_geo, const float* wi_world, const float* wo_world,
                        const float* a, const float* r, const float* m, const float* n_opt,
                        float* out_f /*(L,3)*/, float* out_pdf /*(L)*/, void* stream);
/* Adjoint of mb200_bsdf_eval_pdf's rgb value (what dr.backward through MatDiffBSDF.eval_pdf, mi_plugin.py:1449-1460, returns for
 * a cotangent w on f; the pdf is detached by the path integrator): per lane, at the lane's texel. */
int mb200_bsdf_eval_grad(const mb200_cfg* cfg_host, int64_t L,
                         const float* p, const float* n_geo, const float* wi_world, const float* wo_world,
                         const float* a, const float* r, const float* m, const float* n_opt, const float* w /*(L,3)*/,
                         float* g_a /*(L,3)*/, float* g_r /*(L)*/, float* g_m /*(L)*/, float* g_n /*(L,3)*/, void* stream);
int mb200_bsdf_sample(const mb200_cfg* cfg_host, int64_t L,
                      const float* p, const float* n_geo, const float* wi_world,
                      const float* sample1 /*(L)*/, const float* sample2 /*(L,2)*/,
                      const float* a, const float* r, const float* m, const float* n_opt,
                      float* out_wo /*(L,3) world*/, float* out_pdf /*(L)*/, float* out_weight /*(L,3)*/,
                      void* stream);

/* ---------------------------------------------------------------- TransBSDF (transparency editing; forward only) */
/* Replaces the `TransBSDF` plugin (myutils/mi_plugin.py:1477-1770) that trans_edit.py:16-49 renders with: texels under
 * `mask` get the edited (diffuse + metal + glass reflection / transmission) BSDF with the background `bg` looked up at
 * the twice-refracted screen position (calculate_refracted_screen_coor, :1503-1519); all other texels keep the
 * MatDiffBSDF value.  pdf clamps VoH at 1e-4 and the sample weight is f / (pdf + 1e-4) (:1610, :1659).  The reference
 * never differentiates this plugin, so there is no adjoint entry point. */
typedef struct mb200_trans {
    float ior;               /* 'shape.bsdf.ior' (props['ior'], default 1.3)                                  */
    float spec_trans;        /* 'shape.bsdf.specTrans' (default 0.8)                                          */
    float refract_distance;  /* 100 when the plugin was given keep_albedo_color, else 1 (mi_plugin.py:1484-1489) */
    int32_t reserved;
    const float*   bg;       /* (H,W,3) fp32 'shape.bsdf.bg'                                                   */
    const uint8_t* mask;     /* (H,W) bytes, nonzero = edited texel ('shape.bsdf.mask')                        */
} mb200_trans;
/* as mb200_shade_fwd / mb200_mesh_shade_fwd with the TransBSDF in place of MatDiffBSDF */
int mb200_trans_shade_fwd(const mb200_cfg* cfg_host, const mb200_trans* trans_host,
                          const float* gpos, const float* gnrm,
                          const float* a, const float* r, const float* m, const float* n_opt,
                          const float* env4, const float* hier, const mb200_hier_desc* desc_host,
                          float* partials, void* stream);
int mb200_trans_mesh_shade_fwd(const mb200_cfg* cfg_host, const mb200_trans* trans_host,
                               const mb200_mesh_desc* desc_host, const void* mesh_buf,
                               const float* a, const float* r, const float* m, const float* n_opt,
                               const float* env4, const float* hier, const mb200_hier_desc* hdesc_host,
                               float* partials, void* stream);
/* TransBSDF.eval_pdf (:1748-1761) / .sample (:1521-1544) on arrays of lanes, arguments as mb200_bsdf_eval_pdf / _sample */
int mb200_trans_eval_pdf(const mb200_cfg* cfg_host, const mb200_trans* trans_host, int64_t L,
                         const float* p, const float* n_geo, const float* wi_world, const float* wo_world,
                         const float* a, const float* r, const float* m, const float* n_opt,
                         float* out_f, float* out_pdf, void* stream);
int mb200_trans_sample(const mb200_cfg* cfg_host, const mb200_trans* trans_host, int64_t L,
                       const float* p, const float* n_geo, const float* wi_world,
                       const float* sample1, const float* sample2,
                       const float* a, const float* r, const float* m, const float* n_opt,
                       float* out_wo, float* out_pdf, float* out_weight, void* stream);
/* calculate_refracted_screen_coor on lanes: out_screen (L,2) fp32, out_flat (L) int64 = x + y*stride of the bg texel */
int mb200_trans_refracted_texel(const mb200_cfg* cfg_host, const mb200_trans* trans_host, int64_t L,
                                const float* p, const float* n_geo, const float* wi_world,
                                float* out_screen, int64_t* out_flat, void* stream);

/* Debug / parity: the shared reproducible float32 functions (include/mb200_exact_math.h) and the branch-free exact division /
 * square root of the hierarchy descent, evaluated on the device over arrays.  op: 0 sincospi(x) -> (out0 = sin, out1 = cos),
 * 1 atan2(x, y), 2 acos(x), 3 asin01(x), 4 rsqrt(x) (out0 = the kernels' mbx_rsqrt, out1 = __frsqrt_rn),
 * 5 x / y (out0 = fast path with its deferred fallback, out1 = __fdiv_rn), 6 sqrt(x) (likewise vs __fsqrt_rn).  out1 may be NULL. */
int mb200_debug_exact_math(int op, const float* x, const float* y, int64_t n, float* out0, float* out1, void* stream);

/* Wavefront formulation of mb200_mesh_shade_fwd / mb200_trans_mesh_shade_fwd (trans_host may be NULL = MatDiffBSDF): the path
 * loop is cut into kernels at its rays (traversal-only kernels at high occupancy, coherent shading kernels); path state lives
 * in `scratch` (device, 256-byte aligned, mb200_mesh_fwd_wf_scratch_bytes(cfg) bytes: 156 bytes per path, <= 4 Mi paths at a
 * time).  Same film partials as the one-kernel formulation. */
size_t mb200_mesh_fwd_wf_scratch_bytes(const mb200_cfg* cfg_host);
/* Primary-visibility index of a (mesh, camera) pair: every triangle binned into the pixels its projection (dilated by 0.01 pixel)
 * touches, so that a primary ray tests — with the same exact Moeller-Trumbore, same closest-hit tie rule — only the candidates of
 * its pixel instead of walking the BVH (bit-identical hits at ~1/5 of the cost; the reference's meshes are depth maps seen from the
 * camera they were unprojected with, mesh_recon.py:184-258 / inverse_img_w_mi.py:40-56).  Pixels with too many candidates and
 * meshes the index cannot represent (a triangle at or behind the camera plane, or covering > 64 pixels) fall back to the BVH
 * transparently.  index: mb200_mesh_primary_index_bytes bytes of device memory (256-byte aligned); uses cfg's camera, H and W only;
 * pass it as `primary_index` to the two wavefront entry points below (or NULL). */
size_t mb200_mesh_primary_index_bytes(const mb200_cfg* cfg_host, const mb200_mesh_desc* desc_host);
int mb200_mesh_primary_index_build(const mb200_cfg* cfg_host, const mb200_mesh_desc* desc_host, const void* mesh_buf, void* index, void* stream);
int mb200_mesh_shade_fwd_wf(const mb200_cfg* cfg_host, const mb200_trans* trans_host,
                            const mb200_mesh_desc* desc_host, const void* mesh_buf,
                            const float* a, const float* r, const float* m, const float* n_opt,
                            const float* env4, const float* hier, const mb200_hier_desc* hdesc_host,
                            float* partials, void* scratch, size_t scratch_bytes, const void* primary_index, void* stream);

/* Wavefront formulation of mb200_mesh_shade_bwd (same arguments + scratch: mb200_mesh_bwd_wf_scratch_bytes(cfg) bytes, device,
 * 256-byte aligned; 224 + 128 * (max_depth - 1) bytes per path, <= 8 Mi paths at a time). */
size_t mb200_mesh_bwd_wf_scratch_bytes(const mb200_cfg* cfg_host);
int mb200_mesh_shade_bwd_wf(const mb200_cfg* cfg_host, const mb200_mesh_desc* desc_host, const void* mesh_buf,
                            const float* a, const float* r, const float* m, const float* n_opt,
                            const float* env4, const float* hier, const mb200_hier_desc* hdesc_host,
                            const float* gadj,
                            float* g_a, float* g_r, float* g_m, float* g_n, float* g_env4, int n_env_slabs,
                            void* scratch, size_t scratch_bytes, const void* primary_index, void* stream);

/* ---------------------------------------------------------------- PosMLP */
#define MB200_POSMLP_TCGEN05 0   /* 256-wide layers on tcgen05 tensor cores, FP16x2-split operands, FP32 TMEM accumulators */
#define MB200_POSMLP_FFMA    1   /* all layers in FP32 FFMA (first-generation kernels; kept for A/B measurement)           */
typedef struct mb200_posmlp_desc {
    int32_t n_color;      /* colour / feature channels of the input image (5 for 'arm', 3 for envmap_net) */
    int32_t n_out;        /* output channels                                                   */
    int32_t hidden;       /* 256                                                               */
    int32_t n_freq;       /* multires_view (2)                                                 */
    int32_t output_type;  /* 0 = 'envmap' (softplus), 1 = 'arm' (1.3*tanh + img, STE clamp)    */
    int32_t H, W;         /* pixel grid the N rows enumerate (row-major)                       */
    int32_t impl;         /* MB200_POSMLP_*                                                    */
    int32_t row0;         /* first image row of the N pixels (row shard of an (H, W) image): pixel n has row = row0 + n / W,
                             col = n % W; 0 for the whole image                                */
} mb200_posmlp_desc;
/* parameter packing: [W0 (h0 x d0) | b0 | W1 | b1 | W2 | b2 | W3 | b3 | W4 | b4], nn.Linear row-major (out,in) */
int64_t mb200_posmlp_param_count(const mb200_posmlp_desc* d_host);
size_t  mb200_posmlp_cache_bytes(const mb200_posmlp_desc* d_host, int64_t N);
/* device scratch (16-byte aligned) the forward needs for the pre-split weight images; 0 for MB200_POSMLP_FFMA */
size_t  mb200_posmlp_workspace_bytes(const mb200_posmlp_desc* d_host);
int mb200_posmlp_fwd(const mb200_posmlp_desc* d_host, const float* params, const float* img /*(N,n_color)*/,
                     int64_t N, float* out /*(N,n_out)*/, void* cache /* or NULL: inference */, void* workspace, void* stream);
int mb200_posmlp_bwd(const mb200_posmlp_desc* d_host, const float* params, const float* img, int64_t N,
                     const void* cache, const float* g_out /*(N,n_out)*/,
                     float* g_params /* += */, float* g_img /*(N,n_color) or NULL*/, void* workspace, void* stream);

/* ---------------------------------------------------------------- envmap_utils / computeSH */
/* build_envmap: env (h,w,3) -> c_cdf (h,w), m_cdf (h) */
int mb200_cdf_build(const float* env, int h, int w, float* c_cdf, float* m_cdf, void* stream);
/* sample_envmap: sample2 (2,n) -> dirs (n,3), pdf (n,1), v_idx (n) int64, u_idx (n) int64 */
int mb200_cdf_sample(const float* c_cdf, const float* m_cdf, int h, int w, const float* sample2, int64_t n,
                     float* dirs, float* pdf, int64_t* v_idx, int64_t* u_idx, void* stream);
/* order-4 real SH (25 coef).  angles (n,2) = (theta, phi) fp64; im (h,w,3) fp64; coef (25,3) fp64 */
int mb200_sh_project(const double* im, int h, int w, const double* angles, int64_t n, double* coef, void* stream);
int mb200_sh_reconstruct(const double* coef, int nrows, int ncols, int clip, double* img /*(nrows,ncols,3)*/, void* stream);

/* ---------------------------------------------------------------- on-disk image formats (HOST functions, host pointers) */
/* What mi.Bitmap(path) / mi.util.write_bitmap(path, img) do in the reference (myutils/misc.py:99-111,
 * myutils/mi_plugin.py:701-739, render_final.py:182-202, inverse_img_w_mi.py:54).  Radiance RGBE .hdr (read flat + RLE,
 * write RLE), OpenEXR scanline files (read NONE / ZIPS / ZIP / PIZ, HALF / FLOAT / UINT; write ZIP FLOAT) and PNG (8 / 16 bit,
 * non-interlaced; values / 255 or / 65535 as plt.imread returns them for bg.png / mask.png; write 8 bit).
 * Images are (H, W, C) fp32 row-major in R,G,B(,A) order; single-channel files are C = 1. */
int mb200_image_info(const char* path, int* H, int* W, int* C);
int mb200_image_read(const char* path, float* out_host, int H, int W, int C);
int mb200_image_write(const char* path, const float* img_host, int H, int W, int C);   /* by extension: .hdr (C = 3), .exr, .png (raw values) */
/* flags: MB200_IMG_SRGB = 8-bit PNG colour channels through the sRGB transfer curve, alpha linear — what mi.util.write_bitmap does for
 * 8-bit files (Bitmap::convert(..., srgb_gamma = true); trans_edit.py:47); ignored for .hdr / .exr (linear float formats). */
#define MB200_IMG_SRGB 1
int mb200_image_write_ex(const char* path, const float* img_host, int H, int W, int C, int flags);

/* ---------------------------------------------------------------- fused loss + optimiser step (BRDF phase) */
/* What sits between the forward and the adjoint render of one iteration of `optimize_envmap_ARMN`
 * (inverse_img_w_mi.py:388-432, model_name == 'none'), as four streaming kernels with fixed-order reductions
 * and device-resident scalars (no `.item()` host syncs):
 *   mb200_image_sum       Σ pred                   -> ratio = gt.mean() / pred.detach().mean()            (:388)
 *   mb200_loss_srgb_sums  Σ diff², Σ|diff|, diff = (pred*ratio)^(1/2.2) - gt_srgb  (misc.py:167-170; :391-395)
 *   mb200_loss_srgb_grad  d/d pred of  3*(S1/S0).detach()*mse + l1                                       (:415-419)
 *   mb200_adam_clamped    clamp backward (:370-376) + aux L1 gradient NF.l1_loss(mat, ori)*scale_delta (:398-417)
 *                         + torch.optim.Adam step (:428) + the clamp of the next iteration's maps.
 * `scratch` = mb200_reduce_scratch_bytes() bytes, zero-initialised ONCE by the caller, reusable launch after launch
 * on one stream.  gt_pred_sums = device float[2] = (Σ gt, Σ pred) over ALL ranks; sums2 = device float[2] =
 * (Σ diff², Σ|diff|) over ALL ranks (the caller all-reduces between the calls when the image is sharded);
 * n = this rank's element count, n_total = elements of the whole image. */
#define MB200_ADAM_MAX_SEGS 4
typedef struct mb200_adam_seg {
    float*       p;          /* parameter (updated in place)                                  */
    float*       mat;        /* out: clamp(p_new, lo, hi) — what the next iteration renders   */
    const float* g;          /* gradient of the render loss w.r.t. mat (from mb200_shade_bwd) */
    const float* ori;        /* the map the aux L1 term pulls towards (may be NULL if aux_coeff == 0) */
    float*       m;          /* Adam exp_avg                                                  */
    float*       v;          /* Adam exp_avg_sq                                               */
    int64_t      n;          /* elements                                                      */
    float        lo, hi;     /* clamp range (albedo/metallic 0..1, roughness 0.07..1)         */
    float        aux_coeff;  /* scale_delta / (H*W*C): d(aux L1 mean * scale_delta)/d element */
} mb200_adam_seg;
size_t mb200_reduce_scratch_bytes(void);
int mb200_image_sum(const float* img, int64_t n, float* out /*1*/, void* scratch, void* stream);
int mb200_loss_srgb_sums(const float* img, const float* gt_srgb, int64_t n, const float* gt_pred_sums, float* out2,
                         float* pred_srgb_opt /* (n) or NULL */, void* scratch, void* stream);
int mb200_loss_srgb_grad(const float* img, const float* gt_srgb, int64_t n, const float* gt_pred_sums, const float* sums2,
                         int64_t n_total, float* grad /*(n)*/, void* stream);
int mb200_adam_clamped(const mb200_adam_seg* segs_host, int nseg, float lr, float beta1, float beta2, float eps,
                       int step /* 1-based */, void* stream);

/* ---------------------------------------------------------------- multi-GPU: exchange steps over peer memory (NVLink / NVSwitch)
 * One process per GPU.  Every rank owns a mailbox of mb200_peer_box_bytes() bytes in peer-visible device memory
 * (mb200_peer_alloc; the 64-byte IPC handle goes to the other ranks through any host channel, who map it with mb200_peer_open).
 * The *_peer variants of the loss kernels exchange the three scalar sums of an iteration THROUGH those mailboxes from inside the
 * kernels (the producer's last block stores its partial into every peer's mailbox and raises a flag; the consumer — the next kernel
 * of the iteration — waits for all flags and adds the partials in rank order): what the reference's single process gets from
 * `gt.mean() / pred.mean()`, `loss_mse`, `loss_l1` (inverse_img_w_mi.py:388-395) without a collective library in the loop.
 * mb200_peer_push / mb200_peer_wait move the 2-row film halo of d(loss)/d(image) and the stepped boundary rows of the material
 * maps between neighbouring row shards the same way.  peer->seq: > 0, incremented by the caller once per iteration, identical on
 * all ranks. */
#define MB200_MAX_PEERS 16
#define MB200_PEER_MAX_PUSH 8
#define MB200_PEER_HALO 0
#define MB200_PEER_MAP  1
typedef struct mb200_peer {
    int32_t  rank, world;
    uint32_t seq, reserved;
    void*    box[MB200_MAX_PEERS];       /* mailbox of every rank as mapped into THIS process (box[rank] = its own) */
} mb200_peer;
typedef struct mb200_push_seg { const void* src; void* dst; int64_t n_float4; } mb200_push_seg;
size_t mb200_peer_box_bytes(void);
int mb200_peer_alloc(size_t bytes, void** dev_ptr_host, void* ipc_handle_64_host);   /* cudaMalloc + zero + IPC handle */
int mb200_peer_open(const void* ipc_handle_64_host, void** dev_ptr_host);
int mb200_peer_close(void* dev_ptr);
int mb200_peer_free(void* dev_ptr);
int mb200_image_sum_peer(const float* img, int64_t n, float* out_local /*1*/, void* scratch, const mb200_peer* peer_host, void* stream);
int mb200_loss_srgb_sums_peer(const float* img, const float* gt_srgb, int64_t n, float* gt_pred_sums /* [1] receives the global sum */,
                              float* out2_local, float* pred_srgb_opt, void* scratch, const mb200_peer* peer_host, void* stream);
int mb200_loss_srgb_grad_peer(const float* img, const float* gt_srgb, int64_t n, const float* gt_pred_sums, float* sums2 /* receives the global sums */,
                              int64_t n_total, float* grad, const mb200_peer* peer_host, void* stream);
int mb200_peer_push(const mb200_peer* peer_host, int which, const mb200_push_seg* segs_host, int nseg, int to_up, int to_down,
                    void* ticket, void* stream);
int mb200_peer_wait(const mb200_peer* peer_host, int which, int from_up, int from_down, void* stream);

/* ---------------------------------------------------------------- measurement aid (bench.py only) */
/* Runs `iters` rounds of 16 independent FFMA chains per thread on SMs*8 blocks of 256 threads and writes a
 * checksum to out[0..]; returns the number of FLOPs issued (2 per FFMA) through *flops_host.  Used to MEASURE
 * the non-tensor FP32 peak that the shading kernels are bounded by. */
int mb200_probe_ffma(float* out, int iters, double* flops_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MATERIALIST_B200_H */
