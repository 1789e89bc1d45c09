/*
 * mb_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's differentiable envmap-shading path, used ONLY by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker.  Nothing under materialist_b200/ may import, link or execute it.
 *
 * PARITY STATUS
 *   - BSDF (B-rows): restated line by line from the reference's own source
 *       myutils/mi_plugin.py:60-97 (G1/G_Smith/D_GGX), :217-281 (samplers), :585-595, :645-671
 *       (projection), :1296-1341 (sample_brdf), :1372-1427 (eval_brdf), :1429-1460 (sample/eval_pdf),
 *       :1477-1770 (TransBSDF: refraction, refracted screen coordinate, eval_brdf, sample_brdf).
 *     PINNED: tests/golden/make_bsdf_golden.py EXECUTES that Dr.Jit-typed source (imported from /root/reference) on numpy
 *     stand-ins for the drjit / mitsuba array types (drjit_np_shim.py) -> matdiff_bsdf.npz, trans_bsdf.npz; this file matches
 *     them with median error 0 / 99 % <= 4e-6 and identical texel indices (tests/test_bsdf_plugin_golden.py).  The shared
 *     sub-terms are also pinned against the reference's torch functions (make_golden.py -> bsdf_terms.npz).
 *   - Render operator (P-rows): the arithmetic lives in the un-vendored dependency
 *     mitsuba==3.5.2 / drjit==0.4.6 (requirements.txt:7,:1), absent from /root/reference and from this
 *     image.  Restated from its published algorithm: src/integrators/path.cpp, src/emitters/envmap.cpp,
 *     include/mitsuba/core/distr_2d.h (Hierarchical2D), core/warp.h (square_to_bilinear),
 *     src/samplers/independent.cpp, render/sampler.h, core/random.h (sample_tea_32), drjit random.h
 *     (PCG32), src/render/imageblock.cpp, src/films/hdrfilm.cpp, src/rfilters/gaussian.cpp,
 *     src/python/python/util.py (render / seed_grad); mesh mode (mb_oracle_mesh.c): render/mesh.h, render/interaction.h,
 *     render/scene.cpp.  PINNED on the forward render by the reference's OWN saved output: with the recovered sampler seed
 *     (993) the mesh-mode oracle reproduces output_imgs/indoor/best_results/rendered_img.exr down to its Monte-Carlo noise
 *     pattern (0.5 % / 1.2 % rel-L2 on absolute radiance vs 6-7 % for any other seed; tests/test_reference_render_pin.py),
 *     plus the official PCG32 and Mitsuba TEA known answers; a second shipped render (output_imgs/jinjya, seed 705) confirms it.
 *   - ADJOINT: the BSDF derivatives (eval_brdf_grad) are PINNED against the Jacobian of the reference's own MatDiffBSDF.eval_pdf
 *     source, taken by float64 central differences through the numpy Dr.Jit stand-ins (tests/golden/make_bsdf_grad_golden.py ->
 *     matdiff_bsdf_grad.npz, tests/test_bsdf_grad_pin.py).  The operator-level structure of render_backward (second render with
 *     seed_grad, f2 / detach(p2) weights, detached pdf / MIS / film weights) is restated from Mitsuba's published source and
 *     checked by adjoint identities and a float64 autograd mirror (tests/torch_mirror.py); it has no Mitsuba OUTPUT to pin
 *     against here (not installable) ==> "parity unpinned" for that structure only, until tests/upstream_check.py (guarded,
 *     unit by unit against real mitsuba) can run on a box that has it.
 *
 * Build: see oracle/Makefile.  -ffp-contract=off is REQUIRED (integer decisions must not depend on
 * FMA contraction); fmaf() is used only where upstream writes fmadd.
 * Compile with -DMBO_DOUBLE for a float64 variant used to measure fp32 error and finite differences.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/materialist_b200.h"   /* cfg / hier_desc struct definitions only */
#include "../include/mb200_exact_math.h"   /* reproducible float32 sincospi / atan2 / acos shared with the kernels */

#ifdef MBO_DOUBLE
typedef double real;
#define R(x) x
#define SQRT sqrt
#define SIN sin
#define COS cos
#define ASIN asin
#define ACOS acos
#define ATAN2 atan2
#define EXP exp
#define FLOOR floor
#define FABS fabs
#define FMA fma
#define FMAX fmax
#define FMIN fmin
#define COPYSIGN copysign
/* float64 build (error measurement / finite differences): libm */
static inline void SINCOSPI(double x, double* s, double* c) { *s = sin(x * 3.14159265358979323846); *c = cos(x * 3.14159265358979323846); }
#define RSQRT(x) (1.0 / sqrt(x))
#else
typedef float real;
#define R(x) x##f
#define SQRT sqrtf
#define SIN sinf
#define COS cosf
#define ASIN asinf
#define ACOS acosf
#define ATAN2 atan2f
#define EXP expf
#define FLOOR floorf
#define FABS fabsf
#define FMA fmaf
#define FMAX fmaxf
#define FMIN fminf
#define COPYSIGN copysignf
/* float32 build: the transcendental functions that feed directions (and through them integer decisions and the GGX peak)
 * are the shared reproducible implementations of include/mb200_exact_math.h, bit-identical to the kernels' */
#undef ACOS
#undef ATAN2
#undef ASIN
#define ASIN mbx_asin01      /* only dr::unit_angle uses it (argument in [0, 1]) */
#define ACOS mbx_acos
#define ATAN2 mbx_atan2
#define SINCOSPI mbx_sincospi
#define RSQRT mbx_rsqrt      /* correctly rounded 1/sqrt: one rounding, bit-identical to the device's __frsqrt_rn */
#endif

#define PI_R      ((real)3.14159265358979323846)
#define INV_PI    ((real)0.31830988618379067154)
#define INV_2PI   ((real)0.15915494309189533577)
#define TWO_PI    ((real)6.28318530717958647692)

typedef struct { real x, y, z; } v3;
static inline v3 V3(real x, real y, real z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, real s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline real vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* dr::dot on a 3-vector: a0 b0, then two fmadd (the shading geometry uses this form; Moeller-Trumbore keeps the plain vdot) */
static inline real vdotf(v3 a, v3 b) { return FMA(a.z, b.z, FMA(a.y, b.y, a.x * b.x)); }
static inline v3 vnormalize(v3 a) { real inv = RSQRT(vdotf(a, a)); return vmul(a, inv); }
static inline real safe_sqrt(real x) { return SQRT(FMAX(x, R(0.0))); }
static inline real safe_acos(real x) { return ACOS(FMIN(FMAX(x, R(-1.0)), R(1.0))); }
static inline real pow5(real x) { real x2 = x * x; return x * (x2 * x2); }   /* drjit int pow: square-and-multiply */
static inline real pow4(real x) { real x2 = x * x; return x2 * x2; }

/* ======================================================================== RNG */
/* mitsuba core/random.h sample_tea_32 */
void mbo_tea32(uint32_t v0, uint32_t v1, int rounds, uint32_t* o0, uint32_t* o1) {
    uint32_t sum = 0;
    for (int i = 0; i < rounds; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    *o0 = v0; *o1 = v1;
}

typedef struct { uint64_t state, inc; } pcg32;
static inline uint32_t pcg_next_u32(pcg32* g) {
    uint64_t old = g->state;
    g->state = old * 0x5851f42d4c957f2dull + g->inc;
    uint32_t xorshift = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshift >> rot) | (xorshift << ((~rot + 1u) & 31u));
}
/* drjit PCG32::seed(size=1, initstate, initseq) */
static inline void pcg_seed(pcg32* g, uint64_t initstate, uint64_t initseq) {
    g->state = 0; g->inc = (initseq << 1u) | 1u;
    pcg_next_u32(g); g->state += initstate; pcg_next_u32(g);
}
static inline float pcg_next_float(pcg32* g) {
    union { uint32_t u; float f; } x; x.u = (pcg_next_u32(g) >> 9) | 0x3f800000u; return x.f - 1.0f;
}
/* IndependentSampler::seed(seed, wavefront): per-lane stream from TEA(seed, lane) */
static inline void sampler_seed(pcg32* g, uint32_t seed, uint32_t lane) {
    uint32_t s0, s1; mbo_tea32(seed, lane, 4, &s0, &s1); pcg_seed(g, s0, s1);
}
/* exported unit helpers */
void mbo_pcg32_stream(uint64_t initstate, uint64_t initseq, int n, uint32_t* out) {
    pcg32 g; pcg_seed(&g, initstate, initseq); for (int i = 0; i < n; ++i) out[i] = pcg_next_u32(&g);
}
void mbo_sampler_floats(uint32_t seed, uint32_t lane, int n, float* out) {
    pcg32 g; sampler_seed(&g, seed, lane); for (int i = 0; i < n; ++i) out[i] = pcg_next_float(&g);
}
void mbo_sampler_floats_n(uint32_t seed, uint32_t lane0, int nlanes, int n, float* out) {
    for (int l = 0; l < nlanes; ++l) { pcg32 g; sampler_seed(&g, seed, lane0 + (uint32_t)l); for (int i = 0; i < n; ++i) out[(size_t)l * n + i] = pcg_next_float(&g); }
}
/* mi.render: seed_grad = sample_tea_32(seed, 1)[0]  (util.py) */
uint32_t mbo_seed_grad(uint32_t seed) { uint32_t a, b; mbo_tea32(seed, 1, 4, &a, &b); return a; }

/* ======================================================================== Hierarchical2D */
static int log2i_ceil(uint32_t x) { int l = 0; while ((1u << l) < x) ++l; return l; }

int mbo_hier_describe(int res_x, int res_y, mb200_hier_desc* d) {
    if (res_x < 2 || res_y < 2) return -1;
    memset(d, 0, sizeof(*d));
    d->res_x = res_x; d->res_y = res_y;
    int npx = res_x - 1, npy = res_y - 1;
    int max_level = log2i_ceil((uint32_t)(npx > npy ? npx : npy));
    int n = 0, off = 0;
    d->lvl_off[n] = 0; d->lvl_w[n] = res_x; d->lvl_h[n] = res_y; off += res_x * res_y; ++n;
    int sx = npx, sy = npy;
    for (int level = max_level; level >= 0; --level) {
        sx += sx & 1; sy += sy & 1;
        if (n >= MB200_MAX_LEVELS) return -2;
        off = (off + 3) & ~3;          /* 16-byte aligned levels (layout detail shared with the kernels) */
        d->lvl_off[n] = off; d->lvl_w[n] = sx; d->lvl_h[n] = sy; off += sx * sy; ++n;
        sx >>= 1; sy >>= 1;
    }
    d->n_levels = n; d->total_floats = off;
    return 0;
}
static inline uint32_t lvl_index(uint32_t x, uint32_t y, uint32_t width) {
    return ((x & 1u) | (((x & ~1u) | (y & 1u)) << 1u)) + (y & ~1u) * width;
}

/* data: res_y*res_x non-negative values (luminance*sin_theta) */
void mbo_hier_build(const float* data, const mb200_hier_desc* d, float* hier) {
    const int rx = d->res_x, ry = d->res_y, npx = rx - 1, npy = ry - 1;
    memset(hier, 0, sizeof(float) * (size_t)d->total_floats);
    /* normalisation: double accumulation; rows first, then the row sums (fixed order) */
    double sum = 0.0;
    for (int y = 0; y < npy; ++y) {
        double rs = 0.0;
        for (int x = 0; x < npx; ++x) {
            float v00 = data[y * rx + x], v10 = data[y * rx + x + 1], v01 = data[(y + 1) * rx + x], v11 = data[(y + 1) * rx + x + 1];
            float avg = .25f * (v00 + v10 + v01 + v11);
            rs += (double)avg;
        }
        sum += rs;
    }
    float scale = (float)((double)npx * (double)npy) / (float)sum;
    for (int i = 0; i < rx * ry; ++i) hier[i] = data[i] * scale;
    /* level 1 = patch integrals */
    float* l1 = hier + d->lvl_off[1]; int w1 = d->lvl_w[1];
    for (int y = 0; y < npy; ++y)
        for (int x = 0; x < npx; ++x) {
            float v00 = data[y * rx + x], v10 = data[y * rx + x + 1], v01 = data[(y + 1) * rx + x], v11 = data[(y + 1) * rx + x + 1];
            float avg = .25f * (v00 + v10 + v01 + v11) * scale;
            l1[lvl_index((uint32_t)x, (uint32_t)y, (uint32_t)w1)] = avg;
        }
    /* upper levels: sums of the 2x2 children (contiguous thanks to the swizzle) */
    for (int l = 2; l < d->n_levels; ++l) {
        const float* c = hier + d->lvl_off[l - 1]; int cw = d->lvl_w[l - 1], ch = d->lvl_h[l - 1];
        float* p = hier + d->lvl_off[l]; int pw = d->lvl_w[l];
        for (int y = 0; y < ch / 2; ++y)
            for (int x = 0; x < cw / 2; ++x) {
                const float* q = c + lvl_index((uint32_t)(2 * x), (uint32_t)(2 * y), (uint32_t)cw);
                p[lvl_index((uint32_t)x, (uint32_t)y, (uint32_t)pw)] = q[0] + q[1] + q[2] + q[3];
            }
    }
}

/* warp.h square_to_bilinear */
static inline real square_to_bilinear(real v00, real v10, real v01, real v11, real* sx, real* sy) {
    real r0 = v00 + v10, r1 = v01 + v11;
    if (FABS(r0 - r1) > R(1e-4) * (r0 + r1))
        *sy = (r0 - safe_sqrt(r0 * r0 + *sy * (r1 * r1 - r0 * r0))) / (r0 - r1);
    real c0 = FMA(R(1.0) - *sy, v00, *sy * v01), c1 = FMA(R(1.0) - *sy, v10, *sy * v11);
    if (FABS(c0 - c1) > R(1e-4) * (c0 + c1))
        *sx = (c0 - safe_sqrt(c0 * c0 + *sx * (c1 * c1 - c0 * c0))) / (c0 - c1);
    return FMA(R(1.0) - *sx, c0, *sx * c1);
}

typedef struct { real u, v, pdf; uint32_t offx, offy; } hsample;
static inline hsample hier_sample(const float* hier, const mb200_hier_desc* d, real sx, real sy) {
    uint32_t ox = 0, oy = 0;
    for (int l = d->n_levels - 2; l > 0; --l) {
        const float* lv = hier + d->lvl_off[l];
        ox <<= 1; oy <<= 1;
        uint32_t i = lvl_index(ox, oy, (uint32_t)d->lvl_w[l]);
        real v00 = lv[i], v10 = lv[i + 1], v01 = lv[i + 2], v11 = lv[i + 3];
        sx = FMIN(FMAX(sx, R(0.0)), R(1.0)); sy = FMIN(FMAX(sy, R(0.0)), R(1.0));
        real r0 = v00 + v10, r1 = v01 + v11;
        sy *= r0 + r1;
        int m = sy > r0;
        if (m) { oy += 1; sy -= r0; }
        sy /= m ? r1 : r0;
        real c0 = m ? v01 : v00, c1 = m ? v11 : v10;
        sx *= c0 + c1;
        m = sx > c0;
        if (m) sx -= c0;
        sx /= m ? c1 : c0;
        if (m) ox += 1;
    }
    const int rx = d->res_x;
    uint32_t i = ox + oy * (uint32_t)rx;
    real pdf = square_to_bilinear(hier[i], hier[i + 1], hier[i + rx], hier[i + rx + 1], &sx, &sy);
    hsample h;
    real psx = R(1.0) / (real)(d->res_x - 1), psy = R(1.0) / (real)(d->res_y - 1);
    h.u = ((real)ox + sx) * psx; h.v = ((real)oy + sy) * psy; h.pdf = pdf; h.offx = ox; h.offy = oy;
    return h;
}
static inline real hier_eval(const float* hier, const mb200_hier_desc* d, real u, real v) {
    const int rx = d->res_x, npx = d->res_x - 1, npy = d->res_y - 1;
    real px = u * (real)npx, py = v * (real)npy;
    int ix = (int)px, iy = (int)py;
    uint32_t ox = (uint32_t)ix, oy = (uint32_t)iy;
    if (ox > (uint32_t)(npx - 1)) ox = (uint32_t)(npx - 1);
    if (oy > (uint32_t)(npy - 1)) oy = (uint32_t)(npy - 1);
    real w1x = px - (real)(int)ox, w1y = py - (real)(int)oy, w0x = R(1.0) - w1x, w0y = R(1.0) - w1y;
    uint32_t i = ox + oy * (uint32_t)rx;
    real v00 = hier[i], v10 = hier[i + 1], v01 = hier[i + rx], v11 = hier[i + rx + 1];
    return FMA(w0y, FMA(w0x, v00, w1x * v10), w1y * FMA(w0x, v01, w1x * v11));
}
/* exported units */
void mbo_hier_sample_n(const float* hier, const mb200_hier_desc* d, const float* s, int n,
                       float* uv, float* pdf, int32_t* off) {
    for (int i = 0; i < n; ++i) {
        hsample h = hier_sample(hier, d, (real)s[2 * i], (real)s[2 * i + 1]);
        uv[2 * i] = (float)h.u; uv[2 * i + 1] = (float)h.v; pdf[i] = (float)h.pdf;
        off[2 * i] = (int32_t)h.offx; off[2 * i + 1] = (int32_t)h.offy;
    }
}
void mbo_hier_eval_n(const float* hier, const mb200_hier_desc* d, const float* uv, int n, float* out) {
    for (int i = 0; i < n; ++i) out[i] = (float)hier_eval(hier, d, (real)uv[2 * i], (real)uv[2 * i + 1]);
}

/* ======================================================================== envmap */
int mbo_env_internal_width(int We, int mode) { return mode == MB200_ENV_FILE ? We + 1 : We; }

/* EnvironmentMapEmitter::parameters_changed('data') / file-load constructor */
void mbo_env_prepare(const float* env_in, int He, int We, int mode, float* env_int /*(He,Wi,3)*/,
                     float* hier, const mb200_hier_desc* d) {
    const int Wi = mbo_env_internal_width(We, mode);
    for (int y = 0; y < He; ++y) {
        for (int x = 0; x < We; ++x)
            for (int c = 0; c < 3; ++c) env_int[(y * Wi + x) * 3 + c] = env_in[(y * We + x) * 3 + c];
        if (mode == MB200_ENV_FILE) {
            for (int c = 0; c < 3; ++c) env_int[(y * Wi + We) * 3 + c] = env_in[(y * We) * 3 + c];
        } else {   /* enforce horizontal continuity */
            for (int c = 0; c < 3; ++c) {
                float v01 = .5f * (env_in[(y * We) * 3 + c] + env_in[(y * We + We - 1) * 3 + c]);
                env_int[(y * Wi) * 3 + c] = v01; env_int[(y * Wi + Wi - 1) * 3 + c] = v01;
            }
        }
    }
    float* lum = (float*)malloc(sizeof(float) * (size_t)He * Wi);
    float theta_scale = 1.f / (float)(He - 1) * 3.14159265358979323846f;
    for (int y = 0; y < He; ++y) {
        float theta = (float)y * theta_scale;
        float sin_theta = (float)sin((double)theta);
        for (int x = 0; x < Wi; ++x) {
            const float* t = env_int + (y * Wi + x) * 3;
            float l = t[0] * 0.212671f + t[1] * 0.715160f + t[2] * 0.072169f;
            lum[y * Wi + x] = l * sin_theta;
        }
    }
    mbo_hier_build(lum, d, hier);
    free(lum);
}
/* adjoint of the ingest map */
void mbo_env_grad_finish(const float* g_int, int He, int We, int mode, float* g_env) {
    const int Wi = mbo_env_internal_width(We, mode);
    for (int y = 0; y < He; ++y)
        for (int c = 0; c < 3; ++c) {
            for (int x = 0; x < We; ++x) g_env[(y * We + x) * 3 + c] = g_int[(y * Wi + x) * 3 + c];
            if (mode == MB200_ENV_FILE) g_env[(y * We) * 3 + c] += g_int[(y * Wi + We) * 3 + c];
            else {
                float h = .5f * (g_int[(y * Wi) * 3 + c] + g_int[(y * Wi + Wi - 1) * 3 + c]);
                g_env[(y * We) * 3 + c] = h; g_env[(y * We + We - 1) * 3 + c] = h;
            }
        }
}

typedef struct { uint32_t idx[4]; real w[4]; } bilerp;   /* 4 texel indices + weights */
/* eval_spectrum(uv): includes the half-texel un-shift */
static inline bilerp env_lookup(real u, real v, int Wi, int He, real u_shift) {
    u -= u_shift;
    u -= FLOOR(u); v -= FLOOR(v);
    u *= (real)(Wi - 1); v *= (real)(He - 1);
    uint32_t px = (uint32_t)u, py = (uint32_t)v;
    if (px > (uint32_t)(Wi - 2)) px = (uint32_t)(Wi - 2);
    if (py > (uint32_t)(He - 2)) py = (uint32_t)(He - 2);
    real w1x = u - (real)px, w1y = v - (real)py, w0x = R(1.0) - w1x, w0y = R(1.0) - w1y;
    bilerp b; uint32_t i = py * (uint32_t)Wi + px;
    b.idx[0] = i; b.idx[1] = i + 1; b.idx[2] = i + (uint32_t)Wi; b.idx[3] = i + (uint32_t)Wi + 1;
    /* weights kept per axis: the value is an fmadd chain, see env_value() */
    b.w[0] = w0x; b.w[1] = w1x; b.w[2] = w0y; b.w[3] = w1y;
    return b;
}
static inline v3 env_value(const float* env, const bilerp* b) {
    real w0x = b->w[0], w1x = b->w[1], w0y = b->w[2], w1y = b->w[3];
    real o[3];
    for (int c = 0; c < 3; ++c) {
        real v00 = env[b->idx[0] * 3 + c], v10 = env[b->idx[1] * 3 + c], v01 = env[b->idx[2] * 3 + c], v11 = env[b->idx[3] * 3 + c];
        real v0 = FMA(w0x, v00, w1x * v10), v1 = FMA(w0x, v01, w1x * v11);
        o[c] = FMA(w0y, v0, w1y * v1);
    }
    return V3(o[0], o[1], o[2]);
}
static inline void env_scatter(float* g_env, const bilerp* b, v3 cot) {
    real w0x = b->w[0], w1x = b->w[1], w0y = b->w[2], w1y = b->w[3];
    real ww[4] = { w0y * w0x, w0y * w1x, w1y * w0x, w1y * w1x };
    real cc[3] = { cot.x, cot.y, cot.z };
    for (int k = 0; k < 4; ++k)
        for (int c = 0; c < 3; ++c) {
            float add = (float)(ww[k] * cc[c]);
#pragma omp atomic
            g_env[b->idx[k] * 3 + c] += add;
        }
}
static inline void dir_to_uv(v3 d, real* u, real* v) {
    *u = ATAN2(d.x, -d.z) * INV_2PI; *v = safe_acos(d.y) * INV_PI;
}
static inline real inv_sin_theta(v3 d) {
    const real eps = R(5.9604644775390625e-08);   /* dr::Epsilon<float> */
    return R(1.0) / SQRT(FMAX(d.x * d.x + d.z * d.z, eps * eps));
}
typedef struct { v3 d; real pdf; bilerp b; uint32_t offx, offy; } emsample;
static inline emsample env_sample_direction(const float* hier, const mb200_hier_desc* d, real u_shift, real s0, real s1) {
    emsample e; hsample h = hier_sample(hier, d, s0, s1);
    real u = h.u + u_shift, v = h.v;
    /* theta = v pi, phi = u 2pi: sin / cos of (pi x) evaluated directly (exact range reduction) */
    real st, ct, sp, cp; SINCOSPI(v, &st, &ct); SINCOSPI(R(2.0) * u, &sp, &cp);
    v3 d0 = V3(st * cp, st * sp, ct);
    e.d = V3(d0.y, d0.z, -d0.x);
    e.pdf = h.pdf * inv_sin_theta(e.d) * (R(1.0) / (R(2.0) * PI_R * PI_R));
    e.b = env_lookup(u, v, d->res_x, d->res_y, u_shift);
    e.offx = h.offx; e.offy = h.offy;
    return e;
}
static inline real env_pdf_direction(const float* hier, const mb200_hier_desc* d, real u_shift, v3 dir) {
    real u, v; dir_to_uv(dir, &u, &v);
    u -= u_shift; u -= FLOOR(u); v -= FLOOR(v);
    return hier_eval(hier, d, u, v) * inv_sin_theta(dir) * (R(1.0) / (R(2.0) * PI_R * PI_R));
}
/* exported units */
void mbo_env_eval_n(const float* env_int, int He, int Wi, float u_shift, const float* dirs, int n, float* out) {
    for (int i = 0; i < n; ++i) {
        real u, v; dir_to_uv(V3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), &u, &v);
        bilerp b = env_lookup(u, v, Wi, He, (real)u_shift); v3 e = env_value(env_int, &b);
        out[3 * i] = (float)e.x; out[3 * i + 1] = (float)e.y; out[3 * i + 2] = (float)e.z;
    }
}
void mbo_env_sample_n(const float* env_int, const float* hier, const mb200_hier_desc* d, float u_shift,
                      const float* s, int n, float* dirs, float* pdf, float* weight) {
    for (int i = 0; i < n; ++i) {
        emsample e = env_sample_direction(hier, d, (real)u_shift, (real)s[2 * i], (real)s[2 * i + 1]);
        v3 le = env_value(env_int, &e.b);
        dirs[3 * i] = (float)e.d.x; dirs[3 * i + 1] = (float)e.d.y; dirs[3 * i + 2] = (float)e.d.z;
        pdf[i] = (float)e.pdf;
        real inv = e.pdf != R(0.0) ? R(1.0) / e.pdf : R(0.0);
        weight[3 * i] = (float)(le.x * inv); weight[3 * i + 1] = (float)(le.y * inv); weight[3 * i + 2] = (float)(le.z * inv);
    }
}
void mbo_env_pdf_n(const float* hier, const mb200_hier_desc* d, float u_shift, const float* dirs, int n, float* out) {
    for (int i = 0; i < n; ++i)
        out[i] = (float)env_pdf_direction(hier, d, (real)u_shift, V3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
}

/* ======================================================================== frame */
typedef struct { v3 s, t, n; } frame;
static inline frame make_frame(v3 n) {
    frame f; real sign = COPYSIGN(R(1.0), n.z), a = R(-1.0) / (sign + n.z), b = n.x * n.y * a;
    f.s = V3(sign * (n.x * n.x * a) + R(1.0), sign * b, -sign * n.x);
    f.t = V3(b, FMA(n.y, n.y * a, sign), -n.y);
    f.n = n; return f;
}
static inline v3 to_world(const frame* f, v3 v) {
    return V3(FMA(f->n.x, v.z, FMA(f->t.x, v.y, v.x * f->s.x)),
              FMA(f->n.y, v.z, FMA(f->t.y, v.y, v.x * f->s.y)),
              FMA(f->n.z, v.z, FMA(f->t.z, v.y, v.x * f->s.z)));
}

/* ======================================================================== BSDF (mi_plugin.py) */
/* mi_world_to_screen  mi_plugin.py:645-671 ; matrices row-major */
static inline void world_to_screen(const mb200_cfg* c, v3 p, real* sx, real* sy) {
    const float* V = c->view; const float* P = c->proj;
    real cam[4], clip[4];
    for (int i = 0; i < 4; ++i) cam[i] = (real)V[4 * i] * p.x + (real)V[4 * i + 1] * p.y + (real)V[4 * i + 2] * p.z + (real)V[4 * i + 3] * R(1.0);
    for (int i = 0; i < 4; ++i) clip[i] = (real)P[4 * i] * cam[0] + (real)P[4 * i + 1] * cam[1] + (real)P[4 * i + 2] * cam[2] + (real)P[4 * i + 3] * cam[3];
    real ndcx = clip[0] / clip[3], ndcy = clip[1] / clip[3];
    *sx = (ndcx + R(1.0)) * R(0.5) * (real)c->W;
    *sy = (ndcy + R(1.0)) * R(0.5) * (real)c->H;
}
/* texel fetch index  mi_plugin.py:1378-1381 (row stride = shape[0] = H under the quirk flag) */
static inline int64_t texel_index(const mb200_cfg* c, v3 p) {
    real sx, sy; world_to_screen(c, p, &sx, &sy);
    int64_t x = (int64_t)FLOOR(sx), y = (int64_t)FLOOR(sy);
    int64_t stride = (c->flags & MB200_FLAG_ROW_STRIDE_H) ? c->H : c->W;
    int64_t flat = x + y * stride, last = (int64_t)c->H * c->W - 1;
    /* the reference gathers out of range here (UB); both oracle and kernels clamp instead */
    return flat < 0 ? 0 : (flat > last ? last : flat);
}

typedef struct { real a[3], r, m; v3 n; real bg[3]; int edit; } material;   /* bg / edit: TransBSDF only */

/* TransBSDF mode (mi_plugin.py:1477-1770): when set, eval_brdf / sample_brdf / fetch_material below follow the
 * TransBSDF overrides in EVERY path of this oracle (lanes, G-buffer render, mesh render).  Forward only. */
static mb200_trans g_trans; static int g_trans_on = 0;
void mbo_set_trans(const mb200_trans* t) { if (t) { g_trans = *t; g_trans_on = 1; } else g_trans_on = 0; }
typedef struct { real f[3], pdf; } bsdf_val;
typedef struct { real ga[3], gr, gm; v3 gn; } bsdf_grad;

/* D_GGX mi_plugin.py:89-97 */
static inline real D_GGX(real cos_h, real eta) {
    real alpha = eta * eta, alpha2 = alpha * alpha;
    real denom = (cos_h * cos_h * (alpha2 - R(1.0)) + R(1.0)) + R(1e-6);
    denom = PI_R * denom * denom;
    return alpha2 / denom;
}
/* G1_GGX_Schlick mi_plugin.py:60-68 */
static inline real G1_GGX_Schlick(real NoV, real eta) {
    real k = eta + R(1.0); k = k * k / R(8.0);
    real denom = NoV * (R(1.0) - k) + k + R(1e-6);
    return R(1.0) / denom;
}
/* MatDiffBSDF.eval_brdf mi_plugin.py:1372-1427 (disney_brdf branch); wi = light, wo = view */
static inline bsdf_val trans_eval_brdf(v3 wi, v3 wo, const material* mt);
static inline bsdf_val eval_brdf(v3 wi, v3 wo, const material* mt) {
    if (g_trans_on) return trans_eval_brdf(wi, wo, mt);
    v3 n = mt->n; bsdf_val o;
    v3 h = vnormalize(vadd(wi, wo));
    real NoL = FMAX(vdotf(n, wi), R(0.0)), NoV = FMAX(vdotf(n, wo), R(0.0));
    real VoH = FMAX(vdotf(wo, h), R(0.0)), NoH = FMAX(vdotf(n, h), R(0.0));
    real D = D_GGX(NoH, mt->r);
    real pdf_spec = D / (R(4.0) * FMAX(VoH, R(1e-6))) * NoH;
    real pdf_diff = NoL / PI_R;
    o.pdf = R(0.5) * pdf_spec + R(0.5) * pdf_diff;
    real F_D90 = R(0.5) + R(2.0) * (VoH * VoH) * mt->r;
    real F_D_w_out = R(1.0) + (F_D90 - R(1.0)) * pow5(R(1.0) - NoV);
    real F_D_w_in = R(1.0) + (F_D90 - R(1.0)) * pow5(R(1.0) - NoL);
    real G = G1_GGX_Schlick(NoL, mt->r) * G1_GGX_Schlick(NoV, mt->r);   /* G_Smith :70-76 */
    real X = pow5(R(1.0) - VoH);
    for (int c = 0; c < 3; ++c) {
        real baseColor_d = mt->a[c] * (R(1.0) - mt->m);
        real brdf_diff = baseColor_d / PI_R * F_D_w_out * F_D_w_in * NoL;
        real C_0 = (R(1.0) - mt->m) * R(0.04) + mt->m * mt->a[c];
        real F_m = C_0 + (R(1.0) - C_0) * X;
        real brdf_metal = D * G * F_m / R(4.0) * NoL;
        o.f[c] = brdf_diff + brdf_metal;
    }
    return o;
}
/* hand-derived adjoint of eval_brdf's rgb value w.r.t. (a, r, m, n), cotangent `w` on rgb. pdf is never differentiated. */
static inline void eval_brdf_grad(v3 wi, v3 wo, const material* mt, const real w[3], bsdf_grad* g) {
    v3 n = mt->n; real r = mt->r, m = mt->m;
    v3 h = vnormalize(vadd(wi, wo));
    real dNL = vdotf(n, wi), dNV = vdotf(n, wo), dNH = vdotf(n, h);
    real NoL = FMAX(dNL, R(0.0)), NoV = FMAX(dNV, R(0.0)), VoH = FMAX(vdotf(wo, h), R(0.0)), NoH = FMAX(dNH, R(0.0));
    real alpha = r * r, alpha2 = alpha * alpha;
    real den0 = (NoH * NoH * (alpha2 - R(1.0)) + R(1.0)) + R(1e-6);
    real D = alpha2 / (PI_R * den0 * den0);
    real dD_dalpha2 = (den0 - R(2.0) * alpha2 * NoH * NoH) / (PI_R * den0 * den0 * den0);
    real dD_dr = dD_dalpha2 * R(4.0) * r * r * r;
    real dD_dNoH = R(-2.0) * alpha2 / (PI_R * den0 * den0 * den0) * (R(2.0) * NoH * (alpha2 - R(1.0)));
    real k = (r + R(1.0)); k = k * k / R(8.0);
    real dk_dr = (r + R(1.0)) / R(4.0);
    real G1L = R(1.0) / (NoL * (R(1.0) - k) + k + R(1e-6)), G1V = R(1.0) / (NoV * (R(1.0) - k) + k + R(1e-6));
    real G = G1L * G1V;
    real dG1L_dk = -G1L * G1L * (R(1.0) - NoL), dG1V_dk = -G1V * G1V * (R(1.0) - NoV);
    real dG_dr = dk_dr * (dG1L_dk * G1V + G1L * dG1V_dk);
    real dG_dNoL = -G1L * G1L * (R(1.0) - k) * G1V, dG_dNoV = -G1V * G1V * (R(1.0) - k) * G1L;
    real VoH2 = VoH * VoH;
    real FD90 = R(0.5) + R(2.0) * VoH2 * r;
    real A = pow5(R(1.0) - NoV), B = pow5(R(1.0) - NoL);
    real Fout = R(1.0) + (FD90 - R(1.0)) * A, Fin = R(1.0) + (FD90 - R(1.0)) * B;
    real dFout_dr = R(2.0) * VoH2 * A, dFin_dr = R(2.0) * VoH2 * B;
    real dFout_dNoV = (FD90 - R(1.0)) * R(-5.0) * pow4(R(1.0) - NoV);
    real dFin_dNoL = (FD90 - R(1.0)) * R(-5.0) * pow4(R(1.0) - NoL);
    real X = pow5(R(1.0) - VoH);
    real gr = 0, gm = 0, gNoL = 0, gNoV = 0, gNoH = 0;
    for (int c = 0; c < 3; ++c) {
        real a = mt->a[c];
        real bd = a * (R(1.0) - m);
        real C0 = (R(1.0) - m) * R(0.04) + m * a;
        real Fm = C0 + (R(1.0) - C0) * X;
        real diff_core = Fout * Fin * NoL / PI_R;         /* d diff / d bd */
        real metal_core = D * G / R(4.0) * NoL;             /* d metal / d Fm */
        g->ga[c] = w[c] * ((R(1.0) - m) * diff_core + metal_core * m * (R(1.0) - X));
        gm += w[c] * (-a * diff_core + metal_core * (a - R(0.04)) * (R(1.0) - X));
        gr += w[c] * (bd / PI_R * NoL * (dFout_dr * Fin + Fout * dFin_dr) + Fm / R(4.0) * NoL * (dD_dr * G + D * dG_dr));
        gNoL += w[c] * (bd / PI_R * Fout * (dFin_dNoL * NoL + Fin) + D * Fm / R(4.0) * (dG_dNoL * NoL + G));
        gNoV += w[c] * (bd / PI_R * Fin * NoL * dFout_dNoV + D * Fm / R(4.0) * NoL * dG_dNoV);
        gNoH += w[c] * (G * Fm / R(4.0) * NoL * dD_dNoH);
    }
    g->gr = gr; g->gm = gm;
    v3 gn = V3(0, 0, 0);
    if (dNL > R(0.0)) gn = vadd(gn, vmul(wi, gNoL));
    if (dNV > R(0.0)) gn = vadd(gn, vmul(wo, gNoV));
    if (dNH > R(0.0)) gn = vadd(gn, vmul(h, gNoH));
    g->gn = gn;
}
/* mi_diffuse_sampler mi_plugin.py:255-281 */
/* TransBSDF.calculate_refraction mi_plugin.py:1494-1501 */
static inline v3 trans_refraction(v3 wi, v3 normal, real ior_ratio) {
    real cos_theta_i = vdot(wi, normal);
    real sin2_theta_i = FMAX(R(0.0), R(1.0) - cos_theta_i * cos_theta_i);
    real sin2_theta_t = ior_ratio * ior_ratio * sin2_theta_i;
    real cos_theta_t = safe_sqrt(R(1.0) - sin2_theta_t);
    v3 d = vsub(vmul(vsub(vmul(normal, cos_theta_i), wi), ior_ratio), vmul(normal, cos_theta_t));
    return vnormalize(d);
}
/* TransBSDF.calculate_refracted_screen_coor mi_plugin.py:1503-1519, called with ior_ratio = 1/ior (:1528) and inverted
 * again on entry (:1504), so the first refraction uses `ior` and the second 1/ior.  Both screen coordinates are clamped
 * to [0, width-1] (the WIDTH for both axes, as written). */
static inline void trans_refracted_screen(const mb200_cfg* c, v3 wi, v3 normal, v3 position, real* sx, real* sy) {
    real ior_ratio = R(1.0) / (R(1.0) / (real)g_trans.ior);
    real dist = (real)g_trans.refract_distance;
    v3 d1 = trans_refraction(wi, normal, ior_ratio);
    v3 p1 = vadd(position, vmul(d1, R(0.3) * dist));
    v3 d2 = trans_refraction(vmul(d1, R(-1.0)), normal, R(1.0) / ior_ratio);
    v3 p2 = vadd(p1, vmul(d2, dist));
    real x, y; world_to_screen(c, p2, &x, &y);
    real hi = (real)(c->W - 1);
    x = FMIN(FMAX(x, R(0.0)), hi); y = FMIN(FMAX(y, R(0.0)), hi);        /* NaN -> propagates in Dr.Jit's clamp ... */
    *sx = x > R(0.0) ? x : R(0.0); *sy = y > R(0.0) ? y : R(0.0);        /* ... and select(x > 0, x, 0) sends it to 0 */
}
static inline int64_t trans_refracted_index(const mb200_cfg* c, v3 wi, v3 normal, v3 position) {
    real sx, sy; trans_refracted_screen(c, wi, normal, position, &sx, &sy);
    if (sx != sx) sx = R(0.0); if (sy != sy) sy = R(0.0);
    int64_t stride = (c->flags & MB200_FLAG_ROW_STRIDE_H) ? c->H : c->W;
    int64_t flat = (int64_t)FLOOR(sx) + (int64_t)FLOOR(sy) * stride, last = (int64_t)c->H * c->W - 1;
    return flat < 0 ? 0 : (flat > last ? last : flat);
}
/* TransBSDF.eval_brdf mi_plugin.py:1618-1724; wi = light, wo = view */
static inline bsdf_val trans_eval_brdf(v3 wi, v3 wo, const material* mt) {
    v3 n = mt->n; bsdf_val o;
    const real ior = (real)g_trans.ior, st = (real)g_trans.spec_trans;
    v3 h = vnormalize(vadd(wi, wo));
    real dNL = vdotf(n, wi), dNV = vdotf(n, wo);
    real NoL = FMAX(dNL, R(0.0)), NoV = FMAX(dNV, R(0.0));
    real VoH = FMAX(vdotf(wo, h), R(0.0)), NoH = FMAX(vdotf(n, h), R(0.0));
    real D = D_GGX(NoH, mt->r);
    real pdf_spec = D / (R(4.0) * FMAX(VoH, R(1e-4))) * NoH;
    real pdf_diff = NoL / PI_R;
    o.pdf = R(0.5) * pdf_spec + R(0.5) * pdf_diff;
    real G = G1_GGX_Schlick(NoL, mt->r) * G1_GGX_Schlick(NoV, mt->r);
    real X = pow5(R(1.0) - VoH);
    if (!mt->edit) {                      /* brdf_ori: the MatDiffBSDF (disney) value */
        real F_D90 = R(0.5) + R(2.0) * (VoH * VoH) * mt->r;
        real F_D_w_out = R(1.0) + (F_D90 - R(1.0)) * pow5(R(1.0) - NoV);
        real F_D_w_in = R(1.0) + (F_D90 - R(1.0)) * pow5(R(1.0) - NoL);
        for (int c = 0; c < 3; ++c) {
            real baseColor_d = mt->a[c] * (R(1.0) - mt->m);
            real brdf_diff = baseColor_d / PI_R * F_D_w_out * F_D_w_in * NoL;
            real C_0 = (R(1.0) - mt->m) * R(0.04) + mt->m * mt->a[c];
            real F_m = C_0 + (R(1.0) - C_0) * X;
            o.f[c] = brdf_diff + D * G * F_m / R(4.0) * NoL;
        }
    } else {                              /* bsdf_edit */
        real LoH = FMAX(vdotf(wi, h), R(0.0));
        real hw_in = R(1.0) / (LoH + R(1e-6)), hw_out = R(1.0) / (VoH + R(1e-6));
        real nw_in = R(1.0) / (NoL + R(1e-6)), nw_out = R(1.0) / (NoV + R(1e-6));
        real R_s = (hw_in - ior * hw_out) / (hw_in + ior * hw_out);
        real R_p = (ior * hw_in - hw_out) / (ior * hw_in + hw_out);
        real F_glass = R(0.5) * (R_s * R_s + R_p * R_p);
        real D_hacking = D_GGX(NoH, mt->r * R(0.0) + R(1.0));
        real den = ior * hw_in + hw_out;
        int reflect = NoL * NoV > R(0.0);              /* glass_mask */
        for (int c = 0; c < 3; ++c) {
            real kd = mt->a[c] * (R(1.0) - mt->m) * (R(1.0) - st);
            real glass = (R(1.0) - mt->m) * (mt->bg[c] * st);          /* baseColor_glass */
            real C_0 = (R(1.0) - mt->m) * R(0.04) + mt->m * mt->a[c];
            real F_m = C_0 + (R(1.0) - C_0) * X;
            real brdf_diff = kd / PI_R * NoL;
            real brdf_metal = D * G * F_m / R(4.0) * NoL;
            real btdf = SQRT(glass) * G * D_hacking * (R(1.0) - F_glass) * (ior * ior * hw_in * hw_out) / (nw_in * nw_out * (den * den));
            real spec_edit = glass * D * G / (R(4.0) * nw_in);
            o.f[c] = brdf_diff + brdf_metal + (reflect ? spec_edit : btdf);
        }
    }
    for (int c = 0; c < 3; ++c) o.f[c] = o.f[c] > R(0.0) ? o.f[c] : R(0.0);   /* dr.select(bsdf > 0, bsdf, 0): also NaN -> 0 */
    o.pdf = o.pdf > R(0.0) ? o.pdf : R(0.0);
    return o;
}
static inline v3 nan_to_zero(v3 v) { return V3(v.x != v.x ? 0 : v.x, v.y != v.y ? 0 : v.y, v.z != v.z ? 0 : v.z); }
static inline v3 diffuse_sampler(real u0, real u1, v3 normal) {
    /* theta = asin(sqrt(u0)): sin(theta) = sqrt(u0), cos(theta) = sqrt(1 - u0) (<= 1 ulp from the literal asin -> sin / cos,
     * and reproducible); phi = 2 pi u1 through sincospi */
    real sin_theta = safe_sqrt(u0), cos_theta = safe_sqrt(R(1.0) - u0), sp, cp; SINCOSPI(R(2.0) * u1, &sp, &cp);
    v3 wi = V3(sin_theta * cp, sin_theta * sp, cos_theta);
    frame f = make_frame(normal);
    return nan_to_zero(to_world(&f, wi));
}
/* mi_specular_sampler mi_plugin.py:217-253 */
static inline v3 specular_sampler(real u0, real u1, real roughness, v3 wo, v3 normal) {
    real alpha = roughness * roughness;
    real cos_theta = safe_sqrt((R(1.0) - u0) / (u0 * (alpha * alpha - R(1.0)) + R(1.0)));
    real sin_theta = safe_sqrt(FMAX(R(0.0), R(1.0) - cos_theta * cos_theta));
    real sp, cp; SINCOSPI(R(2.0) * u1, &sp, &cp);            /* phi = 2 pi u1 */
    v3 wh = V3(sin_theta * cp, sin_theta * sp, cos_theta);
    frame f = make_frame(normal);
    wh = to_world(&f, wh);
    v3 wi = vsub(vmul(wh, R(2.0) * vdotf(wo, wh)), wo);
    wi = nan_to_zero(wi);
    return vnormalize(wi);
}
typedef struct { v3 wi; real pdf; real weight[3]; int lobe; } bsdf_smp;
/* MatDiffBSDF.sample_brdf mi_plugin.py:1296-1341 */
static inline bsdf_smp sample_brdf(real s1, real s2x, real s2y, v3 wo, const material* mt) {
    bsdf_smp o; int diffuse = s1 > R(0.5);
    /* both lobes evaluated per lane, then select()ed */
    v3 wd = diffuse_sampler(s2x, s2y, mt->n), ws = specular_sampler(s2x, s2y, mt->r, wo, mt->n);
    o.wi = diffuse ? wd : ws; o.lobe = diffuse;
    bsdf_val bv = eval_brdf(o.wi, wo, mt);
    for (int c = 0; c < 3; ++c) {
        if (g_trans_on) {                 /* TransBSDF.sample_brdf :1609-1613 */
            real w = bv.f[c] / (bv.pdf + R(1e-4));
            o.weight[c] = bv.pdf > R(0.0) ? w : R(0.0);
        } else {
            real w = bv.f[c] / (bv.pdf + R(1e-6));
            o.weight[c] = bv.pdf > R(1e-6) ? w : R(0.0);
        }
    }
    o.pdf = bv.pdf > R(0.0) ? bv.pdf : R(0.0);
    return o;
}
/* exported lane-array units (MatDiffBSDF.eval_pdf / .sample) */
static inline void fetch_material(const mb200_cfg* c, v3 p, v3 n_geo, v3 view, const float* a, const float* r, const float* m,
                                  const float* n_opt, material* mt, int64_t* flat_out) {
    int64_t flat = texel_index(c, p);
    mt->edit = 0; mt->bg[0] = mt->bg[1] = mt->bg[2] = R(0.0);
    if (g_trans_on) {                     /* :1624-1640: mask at the texel, bg at the refracted texel (own texel when unmasked) */
        mt->edit = g_trans.mask[flat] != 0;
        int64_t fr = mt->edit ? trans_refracted_index(c, view, n_geo, p) : flat;
        mt->bg[0] = g_trans.bg[3 * fr]; mt->bg[1] = g_trans.bg[3 * fr + 1]; mt->bg[2] = g_trans.bg[3 * fr + 2];
    }
    mt->a[0] = a[3 * flat]; mt->a[1] = a[3 * flat + 1]; mt->a[2] = a[3 * flat + 2];
    mt->r = r[flat]; mt->m = m[flat];
    mt->n = (c->use_mesh_normal || !n_opt) ? n_geo : V3(n_opt[3 * flat], n_opt[3 * flat + 1], n_opt[3 * flat + 2]);
    if (flat_out) *flat_out = flat;
}
void mbo_bsdf_eval_pdf(const mb200_cfg* c, int64_t L, const float* p, const float* n_geo, const float* wi_w, const float* wo_w,
                       const float* a, const float* r, const float* m, const float* n_opt, float* out_f, float* out_pdf) {
    for (int64_t i = 0; i < L; ++i) {
        material mt; fetch_material(c, V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), V3(n_geo[3 * i], n_geo[3 * i + 1], n_geo[3 * i + 2]),
                                    V3(wi_w[3 * i], wi_w[3 * i + 1], wi_w[3 * i + 2]), a, r, m, n_opt, &mt, 0);
        /* eval_pdf: eval_brdf(wi:=wo (light), wo:=wi (view))  mi_plugin.py:1458 */
        bsdf_val bv = eval_brdf(V3(wo_w[3 * i], wo_w[3 * i + 1], wo_w[3 * i + 2]), V3(wi_w[3 * i], wi_w[3 * i + 1], wi_w[3 * i + 2]), &mt);
        out_f[3 * i] = (float)bv.f[0]; out_f[3 * i + 1] = (float)bv.f[1]; out_f[3 * i + 2] = (float)bv.f[2]; out_pdf[i] = (float)bv.pdf;
    }
}
/* adjoint of MatDiffBSDF.eval_pdf's rgb value on lanes: cotangent w (L,3) -> d/d albedo (L,3), roughness (L), metallic (L), normal (L,3)
 * at each lane's texel (pinned against finite differences of the reference's own source: tests/golden/make_bsdf_grad_golden.py) */
void mbo_bsdf_eval_grad(const mb200_cfg* c, int64_t L, const float* p, const float* n_geo, const float* wi_w, const float* wo_w,
                        const float* a, const float* r, const float* m, const float* n_opt, const float* w,
                        float* g_a, float* g_r, float* g_m, float* g_n) {
    for (int64_t i = 0; i < L; ++i) {
        material mt; fetch_material(c, V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), V3(n_geo[3 * i], n_geo[3 * i + 1], n_geo[3 * i + 2]),
                                    V3(wi_w[3 * i], wi_w[3 * i + 1], wi_w[3 * i + 2]), a, r, m, n_opt, &mt, 0);
        real ww[3] = { (real)w[3 * i], (real)w[3 * i + 1], (real)w[3 * i + 2] };
        bsdf_grad bg; eval_brdf_grad(V3(wo_w[3 * i], wo_w[3 * i + 1], wo_w[3 * i + 2]), V3(wi_w[3 * i], wi_w[3 * i + 1], wi_w[3 * i + 2]), &mt, ww, &bg);
        for (int k = 0; k < 3; ++k) g_a[3 * i + k] = (float)bg.ga[k];
        g_r[i] = (float)bg.gr; g_m[i] = (float)bg.gm;
        g_n[3 * i] = (float)bg.gn.x; g_n[3 * i + 1] = (float)bg.gn.y; g_n[3 * i + 2] = (float)bg.gn.z;
    }
}
void mbo_bsdf_sample(const mb200_cfg* c, int64_t L, const float* p, const float* n_geo, const float* wi_w,
                     const float* s1, const float* s2, const float* a, const float* r, const float* m, const float* n_opt,
                     float* out_wo, float* out_pdf, float* out_w) {
    for (int64_t i = 0; i < L; ++i) {
        material mt; fetch_material(c, V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), V3(n_geo[3 * i], n_geo[3 * i + 1], n_geo[3 * i + 2]),
                                    V3(wi_w[3 * i], wi_w[3 * i + 1], wi_w[3 * i + 2]), a, r, m, n_opt, &mt, 0);
        bsdf_smp bs = sample_brdf((real)s1[i], (real)s2[2 * i], (real)s2[2 * i + 1], V3(wi_w[3 * i], wi_w[3 * i + 1], wi_w[3 * i + 2]), &mt);
        out_wo[3 * i] = (float)bs.wi.x; out_wo[3 * i + 1] = (float)bs.wi.y; out_wo[3 * i + 2] = (float)bs.wi.z;
        out_pdf[i] = (float)bs.pdf;
        for (int k = 0; k < 3; ++k) out_w[3 * i + k] = (float)bs.weight[k];
    }
}
void mbo_trans_refracted_texel(const mb200_cfg* c, int64_t L, const float* p, const float* n_geo, const float* wi_w, float* out_screen, int64_t* out_flat) {
    for (int64_t i = 0; i < L; ++i) {
        v3 P = V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), N = V3(n_geo[3 * i], n_geo[3 * i + 1], n_geo[3 * i + 2]), Wv = V3(wi_w[3 * i], wi_w[3 * i + 1], wi_w[3 * i + 2]);
        real sx, sy; trans_refracted_screen(c, Wv, N, P, &sx, &sy);
        out_screen[2 * i] = (float)sx; out_screen[2 * i + 1] = (float)sy; out_flat[i] = trans_refracted_index(c, Wv, N, P);
    }
}
/* sub-term unit exports (pinned against the reference's torch D_GGX / G_Smith / fresnelSchlick) */
void mbo_terms(int n, const float* cos_h, const float* NoV, const float* NoL, const float* VoH, const float* rough, const float* F0,
               float* D, float* G, float* F) {
    for (int i = 0; i < n; ++i) {
        D[i] = (float)D_GGX((real)cos_h[i], (real)rough[i]);
        G[i] = (float)(G1_GGX_Schlick((real)NoL[i], (real)rough[i]) * G1_GGX_Schlick((real)NoV[i], (real)rough[i]));
        real x = pow5(R(1.0) - (real)VoH[i]); F[i] = (float)((real)F0[i] + (R(1.0) - (real)F0[i]) * x);
    }
}
void mbo_world_to_screen(const mb200_cfg* c, int n, const float* p, float* out, int64_t* flat) {
    for (int i = 0; i < n; ++i) {
        real sx, sy; v3 q = V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        world_to_screen(c, q, &sx, &sy); out[2 * i] = (float)sx; out[2 * i + 1] = (float)sy; flat[i] = texel_index(c, q);
    }
}

/* ======================================================================== film */
static inline real gauss_w(real x) {   /* gaussian rfilter, stddev .5, radius 2: alpha = -1/(2 sigma^2) = -2 */
    const real bias = EXP(R(-2.0) * R(4.0));
    return FMAX(R(0.0), EXP(R(-2.0) * (x * x)) - bias);
}
/* weights of the 5 columns (or rows) px-2..px+2 for jitter j in [0,1): rel = (px+o+.5) - (px+j) */
static inline void film_taps(real j, real w[5]) {
    for (int o = -2; o <= 2; ++o) {
        real rel = ((real)o + R(0.5)) - j;
        w[o + 2] = (FABS(rel) <= R(2.0)) ? gauss_w(rel) : R(0.0);
    }
}

/* ======================================================================== sensor */
static inline v3 primary_dir(const mb200_cfg* c, real sx, real sy) {
    /* perspective sensor, x_fov along x; camera space looks along +z with x to the left */
    real t = (real)c->tan_half_fov_x, aspect = (real)c->W / (real)c->H;
    v3 l = V3((R(1.0) - R(2.0) * sx / (real)c->W) * t, (R(1.0) - R(2.0) * sy / (real)c->H) * t / aspect, R(1.0));
    l = vnormalize(l);
    const float* M = c->cam_to_world;
    return V3((real)M[0] * l.x + (real)M[1] * l.y + (real)M[2] * l.z,
              (real)M[4] * l.x + (real)M[5] * l.y + (real)M[6] * l.z,
              (real)M[8] * l.x + (real)M[9] * l.y + (real)M[10] * l.z);
}
static inline v3 cam_origin(const mb200_cfg* c) { return V3(c->cam_to_world[3], c->cam_to_world[7], c->cam_to_world[11]); }

static inline real mis_weight(real a, real b) {
    a *= a; b *= b; real w = a / (a + b);
    return isfinite(w) ? w : R(0.0);
}

/* ======================================================================== one path */
typedef struct {
    const mb200_cfg* c; const float *gpos, *gnrm, *a, *r, *m, *n_opt, *env; const float* hier; const mb200_hier_desc* d;
} scene_t;

typedef struct {
    real L[3]; real jx, jy;
    /* everything the adjoint needs */
    int valid; int64_t flat; material mt; v3 view;
    emsample em; int active_em; bsdf_val f_em; real mis_em; v3 le_em;
    v3 d_bs; bilerp b_bs; v3 le_bs; real mis_bs; real w_bs[3]; int bs_active; int lobe;
    bilerp b_miss;
} path_rec;

static void trace_path(const scene_t* S, int px, int py, int s, int ad_weights, path_rec* o) {
    const mb200_cfg* c = S->c;
    const int64_t pixel = (int64_t)py * c->W + px;
    pcg32 rng; sampler_seed(&rng, c->seed, (uint32_t)(pixel * c->spp + s));
    o->jx = (real)pcg_next_float(&rng); o->jy = (real)pcg_next_float(&rng);
    o->L[0] = o->L[1] = o->L[2] = 0; o->active_em = 0; o->bs_active = 0; o->lobe = -1; o->flat = -1;
    memset(&o->em, 0, sizeof(o->em));
    const float* gp = S->gpos + 4 * pixel; const float* gn = S->gnrm + 4 * pixel;
    const real u_shift = (real)c->env_u_shift; const int Wi = S->d->res_x, He = S->d->res_y;
    o->valid = gp[3] != 0.0f;
    if (!o->valid) {     /* primary ray escapes: L = Le(ray.d), MIS weight 1 */
        v3 dir = primary_dir(c, (real)px + o->jx, (real)py + o->jy);
        real u, v; dir_to_uv(dir, &u, &v);
        o->b_miss = env_lookup(u, v, Wi, He, u_shift);
        v3 le = env_value(S->env, &o->b_miss);
        o->L[0] = le.x; o->L[1] = le.y; o->L[2] = le.z;
        return;
    }
    if (c->max_depth < 2) return;
    real uex = (real)pcg_next_float(&rng), uey = (real)pcg_next_float(&rng);
    real s1 = (real)pcg_next_float(&rng);
    real s2x = (real)pcg_next_float(&rng), s2y = (real)pcg_next_float(&rng);
    (void)pcg_next_float(&rng);   /* russian-roulette draw: consumed, never applied (rr_depth 5 > max_depth) */
    v3 p = V3(gp[0], gp[1], gp[2]), n_geo = V3(gn[0], gn[1], gn[2]);
    o->view = vnormalize(vsub(cam_origin(c), p));
    fetch_material(c, p, n_geo, o->view, S->a, S->r, S->m, S->n_opt, &o->mt, &o->flat);
    /* ---- emitter sampling */
    o->em = env_sample_direction(S->hier, S->d, u_shift, uex, uey);
    o->active_em = o->em.pdf != R(0.0);
    o->le_em = env_value(S->env, &o->em.b);
    o->f_em = eval_brdf(o->em.d, o->view, &o->mt);
    o->mis_em = mis_weight(o->em.pdf, o->f_em.pdf);
    if (o->active_em) {
        real inv = R(1.0) / o->em.pdf;
        o->L[0] += o->f_em.f[0] * (o->le_em.x * inv) * o->mis_em;
        o->L[1] += o->f_em.f[1] * (o->le_em.y * inv) * o->mis_em;
        o->L[2] += o->f_em.f[2] * (o->le_em.z * inv) * o->mis_em;
    }
    /* ---- BSDF sampling */
    bsdf_smp bs = sample_brdf(s1, s2x, s2y, o->view, &o->mt);
    o->lobe = bs.lobe;
    frame F = make_frame(n_geo);
    o->d_bs = (c->flags & MB200_FLAG_WO_WORLD_QUIRK) ? to_world(&F, bs.wi) : bs.wi;   /* mi_plugin.py:1444 */
    for (int k = 0; k < 3; ++k) o->w_bs[k] = bs.weight[k];
    if (ad_weights) {   /* path.cpp: re-evaluate with the detached direction, weight = f2 / detach(p2) */
        bsdf_val b2 = eval_brdf(o->d_bs, o->view, &o->mt);
        if (b2.pdf > R(0.0)) for (int k = 0; k < 3; ++k) o->w_bs[k] = b2.f[k] / b2.pdf;
    }
    real tmax = FMAX(o->w_bs[0], FMAX(o->w_bs[1], o->w_bs[2]));
    if (tmax == R(0.0)) return;           /* active = active_next && throughput_max != 0 */
    /* next iteration: the spawned ray escapes (no-occlusion G-buffer) */
    real em_pdf = env_pdf_direction(S->hier, S->d, u_shift, o->d_bs);
    o->mis_bs = mis_weight(bs.pdf, em_pdf);
    real u, v; dir_to_uv(o->d_bs, &u, &v);
    o->b_bs = env_lookup(u, v, Wi, He, u_shift);
    o->le_bs = env_value(S->env, &o->b_bs);
    o->bs_active = bs.pdf > R(0.0);       /* emitter->eval(si, prev_bsdf_pdf > 0) */
    if (o->bs_active) {
        o->L[0] += o->w_bs[0] * o->le_bs.x * o->mis_bs;
        o->L[1] += o->w_bs[1] * o->le_bs.y * o->mis_bs;
        o->L[2] += o->w_bs[2] * o->le_bs.z * o->mis_bs;
    }
}

/* ======================================================================== render forward */
/* img: (rows, W, 3) for the shard rows. indices (optional): (rows*W*spp, 4) int32. */
int mbo_render_fwd(const mb200_cfg* c, const float* gpos, const float* gnrm, const float* a, const float* r, const float* m,
                   const float* n_opt, const float* env_int, const float* hier, const mb200_hier_desc* d,
                   float* img, int32_t* indices) {
    if ((double)c->H * c->W * c->spp >= 4294967296.0) return MB200_ERANGE;
    scene_t S = { c, gpos, gnrm, a, r, m, n_opt, env_int, hier, d };
    const int H = c->H, W = c->W, spp = c->spp;
    const int ad = (c->flags & MB200_FLAG_AD_WEIGHTS) != 0;
    const int gaussian = c->filter == MB200_FILTER_GAUSSIAN;
    const int halo = gaussian ? 2 : 0;
    int r0 = c->row0 - halo, r1 = c->row0 + c->rows + halo; if (r0 < 0) r0 = 0; if (r1 > H) r1 = H;
    const int prow = r1 - r0, taps = gaussian ? 25 : 1;
    real* part = (real*)calloc((size_t)prow * W * taps * 4, sizeof(real));
    if (!part) return MB200_EINVAL;
#pragma omp parallel for schedule(dynamic, 1)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            real* P = part + ((size_t)(py - r0) * W + px) * taps * 4;
            for (int s = 0; s < spp; ++s) {
                path_rec o; trace_path(&S, px, py, s, ad, &o);
                if (indices && py >= c->row0 && py < c->row0 + c->rows) {
                    int32_t* I = indices + (((size_t)(py - c->row0) * W + px) * spp + s) * 4;
                    I[0] = (int32_t)o.em.offx; I[1] = (int32_t)o.em.offy; I[2] = (int32_t)o.flat; I[3] = o.lobe;
                }
                if (gaussian) {
                    real wx[5], wy[5]; film_taps(o.jx, wx); film_taps(o.jy, wy);
                    for (int j = 0; j < 5; ++j) for (int i = 0; i < 5; ++i) {
                        real w = wx[i] * wy[j]; real* q = P + (j * 5 + i) * 4;
                        q[0] += w * o.L[0]; q[1] += w * o.L[1]; q[2] += w * o.L[2]; q[3] += w;
                    }
                } else { P[0] += o.L[0]; P[1] += o.L[1]; P[2] += o.L[2]; P[3] += R(1.0); }
            }
        }
    /* gather + develop */
    for (int qy = c->row0; qy < c->row0 + c->rows; ++qy)
        for (int qx = 0; qx < W; ++qx) {
            real acc[4] = {0, 0, 0, 0};
            if (gaussian) {
                for (int j = 0; j < 5; ++j) for (int i = 0; i < 5; ++i) {
                    int sy = qy - (j - 2), sx = qx - (i - 2);      /* source pixel whose tap (i,j) lands on q */
                    if (sx < 0 || sx >= W || sy < r0 || sy >= r1) continue;
                    const real* q = part + (((size_t)(sy - r0) * W + sx) * 25 + (j * 5 + i)) * 4;
                    for (int k = 0; k < 4; ++k) acc[k] += q[k];
                }
            } else { const real* q = part + ((size_t)(qy - r0) * W + qx) * 4; for (int k = 0; k < 4; ++k) acc[k] = q[k]; }
            real wsum = acc[3] == R(0.0) ? R(1.0) : acc[3];
            float* o = img + ((size_t)(qy - c->row0) * W + qx) * 3;
            o[0] = (float)(acc[0] / wsum); o[1] = (float)(acc[1] / wsum); o[2] = (float)(acc[2] / wsum);
        }
    free(part);
    return MB200_OK;
}

/* Per-lane decision record (the oracle side of mb200_debug_sample_record): 12 int32 words per lane =
 * hier off.x, off.y, texel index (-1: no surface), lobe, envmap cell of the emitter sample, envmap cell of the BSDF-sampled
 * direction (of the primary ray where there is no surface; computed whether or not the sample carries weight), IEEE bits
 * of the emitter direction and of the BSDF-sampled direction.  out_L (S,3) or NULL: the lane's radiance. */
int mbo_sample_record(const mb200_cfg* c, const float* gpos, const float* gnrm, const float* a, const float* r, const float* m,
                      const float* n_opt, const float* env_int, const float* hier, const mb200_hier_desc* d,
                      int32_t* out, float* out_L) {
    scene_t S = { c, gpos, gnrm, a, r, m, n_opt, env_int, hier, d };
    const int W = c->W, spp = c->spp, ad = (c->flags & MB200_FLAG_AD_WEIGHTS) != 0;
    const real u_shift = (real)c->env_u_shift;
#pragma omp parallel for schedule(dynamic, 1)
    for (int py = c->row0; py < c->row0 + c->rows; ++py)
        for (int px = 0; px < W; ++px)
            for (int s = 0; s < spp; ++s) {
                path_rec o; memset(&o.d_bs, 0, sizeof(o.d_bs)); trace_path(&S, px, py, s, ad, &o);
                const size_t i = ((size_t)(py - c->row0) * W + px) * spp + s;
                int32_t* I = out + 12 * i;
                union { float f; int32_t i; } cv;
                I[0] = I[1] = 0; I[2] = -1; I[3] = -1; I[4] = I[5] = -1; for (int k = 6; k < 12; ++k) I[k] = 0;
                if (!o.valid) {
                    v3 dir = primary_dir(c, (real)px + o.jx, (real)py + o.jy);
                    I[5] = (int32_t)o.b_miss.idx[0];
                    cv.f = (float)dir.x; I[9] = cv.i; cv.f = (float)dir.y; I[10] = cv.i; cv.f = (float)dir.z; I[11] = cv.i;
                } else if (c->max_depth >= 2) {
                    I[0] = (int32_t)o.em.offx; I[1] = (int32_t)o.em.offy; I[2] = (int32_t)o.flat; I[3] = o.lobe; I[4] = (int32_t)o.em.b.idx[0];
                    real u, v; dir_to_uv(o.d_bs, &u, &v);
                    bilerp b = env_lookup(u, v, d->res_x, d->res_y, u_shift);
                    I[5] = (int32_t)b.idx[0];
                    cv.f = (float)o.em.d.x; I[6] = cv.i; cv.f = (float)o.em.d.y; I[7] = cv.i; cv.f = (float)o.em.d.z; I[8] = cv.i;
                    cv.f = (float)o.d_bs.x; I[9] = cv.i; cv.f = (float)o.d_bs.y; I[10] = cv.i; cv.f = (float)o.d_bs.z; I[11] = cv.i;
                }
                if (out_L) { out_L[3 * i] = (float)o.L[0]; out_L[3 * i + 1] = (float)o.L[1]; out_L[3 * i + 2] = (float)o.L[2]; }
            }
    return MB200_OK;
}

/* ======================================================================== render backward */
/* c->seed must be seed_grad.  grad_img: FULL image (H,W,3).  Gradient buffers are full-size and accumulated (+=). */
int mbo_render_bwd(const mb200_cfg* c, const float* gpos, const float* gnrm, const float* a, const float* r, const float* m,
                   const float* n_opt, const float* env_int, const float* hier, const mb200_hier_desc* d,
                   const float* grad_img, float* g_a, float* g_r, float* g_m, float* g_n, float* g_env_int) {
    if ((double)c->H * c->W * c->spp >= 4294967296.0) return MB200_ERANGE;
    scene_t S = { c, gpos, gnrm, a, r, m, n_opt, env_int, hier, d };
    const int H = c->H, W = c->W, spp = c->spp;
    const int gaussian = c->filter == MB200_FILTER_GAUSSIAN;
    const int want_mat = g_a || g_r || g_m || g_n;
    /* 1. film weights of the seed_grad render -> G[q] = grad[q] / W_q on rows [row0-2, row0+rows+2) */
    int q0 = c->row0 - (gaussian ? 2 : 0), q1 = c->row0 + c->rows + (gaussian ? 2 : 0); if (q0 < 0) q0 = 0; if (q1 > H) q1 = H;
    real* G = (real*)calloc((size_t)(q1 - q0) * W * 3, sizeof(real));
    if (gaussian) {
        int w0 = q0 - 2, w1 = q1 + 2; if (w0 < 0) w0 = 0; if (w1 > H) w1 = H;
        real* wp = (real*)calloc((size_t)(w1 - w0) * W * 25, sizeof(real));
#pragma omp parallel for schedule(static)
        for (int py = w0; py < w1; ++py)
            for (int px = 0; px < W; ++px) {
                real* P = wp + ((size_t)(py - w0) * W + px) * 25;
                for (int s = 0; s < spp; ++s) {
                    pcg32 rng; sampler_seed(&rng, c->seed, (uint32_t)(((int64_t)py * W + px) * spp + s));
                    real jx = (real)pcg_next_float(&rng), jy = (real)pcg_next_float(&rng);
                    real wx[5], wy[5]; film_taps(jx, wx); film_taps(jy, wy);
                    for (int j = 0; j < 5; ++j) for (int i = 0; i < 5; ++i) P[j * 5 + i] += wx[i] * wy[j];
                }
            }
        for (int qy = q0; qy < q1; ++qy)
            for (int qx = 0; qx < W; ++qx) {
                real ws = 0;
                for (int j = 0; j < 5; ++j) for (int i = 0; i < 5; ++i) {
                    int sy = qy - (j - 2), sx = qx - (i - 2);
                    if (sx < 0 || sx >= W || sy < w0 || sy >= w1) continue;
                    ws += wp[((size_t)(sy - w0) * W + sx) * 25 + j * 5 + i];
                }
                if (ws == R(0.0)) ws = R(1.0);
                for (int k = 0; k < 3; ++k) G[((size_t)(qy - q0) * W + qx) * 3 + k] = (real)grad_img[((size_t)qy * W + qx) * 3 + k] / ws;
            }
        free(wp);
    } else {
        for (int qy = q0; qy < q1; ++qy) for (int qx = 0; qx < W; ++qx) for (int k = 0; k < 3; ++k)
            G[((size_t)(qy - q0) * W + qx) * 3 + k] = (real)grad_img[((size_t)qy * W + qx) * 3 + k] / (real)spp;
    }
    /* 2. adjoint render */
#pragma omp parallel for schedule(dynamic, 1)
    for (int py = c->row0; py < c->row0 + c->rows; ++py)
        for (int px = 0; px < W; ++px) {
            real ga[3] = {0, 0, 0}, gr = 0, gm = 0; v3 gn = V3(0, 0, 0); int64_t flat = -1;
            for (int s = 0; s < spp; ++s) {
                path_rec o; trace_path(&S, px, py, s, 1, &o);
                /* film adjoint: delta = sum_q w_q(s) G[q] */
                real dl[3] = {0, 0, 0};
                if (gaussian) {
                    real wx[5], wy[5]; film_taps(o.jx, wx); film_taps(o.jy, wy);
                    for (int j = 0; j < 5; ++j) for (int i = 0; i < 5; ++i) {
                        int qy = py + (j - 2), qx = px + (i - 2);
                        if (qx < 0 || qx >= W || qy < 0 || qy >= H) continue;
                        real w = wx[i] * wy[j]; const real* g = G + ((size_t)(qy - q0) * W + qx) * 3;
                        dl[0] += w * g[0]; dl[1] += w * g[1]; dl[2] += w * g[2];
                    }
                } else { const real* g = G + ((size_t)(py - q0) * W + px) * 3; dl[0] = g[0]; dl[1] = g[1]; dl[2] = g[2]; }
                if (!o.valid) {
                    if (g_env_int) env_scatter(g_env_int, &o.b_miss, V3(dl[0], dl[1], dl[2]));
                    continue;
                }
                if (c->max_depth < 2) continue;
                flat = o.flat;
                if (o.active_em) {
                    real inv = R(1.0) / o.em.pdf;
                    if (want_mat) {
                        real w[3] = { dl[0] * (o.le_em.x * inv) * o.mis_em, dl[1] * (o.le_em.y * inv) * o.mis_em, dl[2] * (o.le_em.z * inv) * o.mis_em };
                        bsdf_grad bg; eval_brdf_grad(o.em.d, o.view, &o.mt, w, &bg);
                        for (int k = 0; k < 3; ++k) ga[k] += bg.ga[k];
                        gr += bg.gr; gm += bg.gm; gn = vadd(gn, bg.gn);
                    }
                    if (g_env_int) {
                        v3 cot = V3(dl[0] * o.f_em.f[0] * inv * o.mis_em, dl[1] * o.f_em.f[1] * inv * o.mis_em, dl[2] * o.f_em.f[2] * inv * o.mis_em);
                        env_scatter(g_env_int, &o.em.b, cot);
                    }
                }
                if (o.bs_active) {
                    if (want_mat) {
                        bsdf_val b2 = eval_brdf(o.d_bs, o.view, &o.mt);
                        if (b2.pdf > R(0.0)) {   /* weight = f2/detach(p2); else the (zero) primal weight carries no usable gradient */
                            real ip = R(1.0) / b2.pdf;
                            real w[3] = { dl[0] * o.le_bs.x * o.mis_bs * ip, dl[1] * o.le_bs.y * o.mis_bs * ip, dl[2] * o.le_bs.z * o.mis_bs * ip };
                            bsdf_grad bg; eval_brdf_grad(o.d_bs, o.view, &o.mt, w, &bg);
                            for (int k = 0; k < 3; ++k) ga[k] += bg.ga[k];
                            gr += bg.gr; gm += bg.gm; gn = vadd(gn, bg.gn);
                        }
                    }
                    if (g_env_int) {
                        v3 cot = V3(dl[0] * o.w_bs[0] * o.mis_bs, dl[1] * o.w_bs[1] * o.mis_bs, dl[2] * o.w_bs[2] * o.mis_bs);
                        env_scatter(g_env_int, &o.b_bs, cot);
                    }
                }
            }
            if (flat >= 0 && want_mat) {
                /* scatter per texel (a plain per-pixel store when texel == pixel) */
                if (g_a) for (int k = 0; k < 3; ++k) {
                    float add = (float)ga[k];
#pragma omp atomic
                    g_a[3 * flat + k] += add;
                }
                if (g_r) { float add = (float)gr;
#pragma omp atomic
                    g_r[flat] += add; }
                if (g_m) { float add = (float)gm;
#pragma omp atomic
                    g_m[flat] += add; }
                if (g_n && !c->use_mesh_normal) {
                    float add[3] = { (float)gn.x, (float)gn.y, (float)gn.z };
                    for (int k = 0; k < 3; ++k) {
#pragma omp atomic
                        g_n[3 * flat + k] += add[k];
                    }
                }
            }
        }
    free(G);
    return MB200_OK;
}

#ifdef _OPENMP
#include <omp.h>
int mbo_omp_max_threads(void) { return omp_get_max_threads(); }
#else
int mbo_omp_max_threads(void) { return 1; }
#endif

int mbo_is_double(void) {
#ifdef MBO_DOUBLE
    return 1;
#else
    return 0;
#endif
}

/* ======================================================================== mesh mode */
#include "mb_oracle_mesh.c"
