// mb200_mesh.cu — mesh mode of the render operator (sm_100a): what the reference REALLY renders for its scene
// (inverse_img_w_mi.py:30-56: one `ply` shape with MatDiffBSDF, `path` integrator max_depth 4, envmap emitter):
// jittered primary rays against the depth-derived triangle mesh (myutils/mesh_recon.py:86-331), per-sample triangle
// hits, shadow rays for the emitter samples, up to max_depth-1 scattering vertices with the material looked up at
// EVERY vertex through mi_world_to_screen(si.p) (mi_plugin.py:1435,1456), forward and adjoint.
//
// Upstream units restated (mitsuba==3.5.2, un-vendored): src/integrators/path.cpp (loop, draw order, MIS, AD-pass
// weight), render/mesh.h (Moeller-Trumbore ray_intersect_triangle, compute_surface_interaction without UVs),
// render/interaction.h (offset_p / spawn_ray / spawn_ray_to, RayEpsilon, ShadowEpsilon), emitters/envmap.cpp
// (sample_direction target point), render/scene.cpp (sample_emitter_direction visibility test),
// Mesh::recompute_vertex_normals (angle-weighted).
//
// Acceleration structure (this file's own, built ON THE GPU, no pointers): triangles are sorted by the 30-bit Morton
// code of their bounding-box centre (cub radix sort), 4 consecutive sorted triangles form a leaf, and an IMPLICIT
// complete 4-ary tree is laid over the leaves: node i of level l has children 4i..4i+3 of level l-1.  The four
// child boxes of a node are stored together as six float4 (lo.x[4], lo.y[4], lo.z[4], hi.x[4], hi.y[4], hi.z[4])
// = 96 contiguous bytes, so one traversal step is six 16-byte loads and four slab tests.  Sorted triangles carry
// their vertices (3 float4 per slot, original triangle id in p0.w) and, for smooth shading frames, their three
// vertex normals.  The closest hit is defined as (smallest t, then smallest triangle id) — independent of traversal
// order and of the BVH — and the Moeller-Trumbore test is written with non-contracted IEEE operations in exactly the
// oracle's order, so for bit-identical rays (all primary rays) the hit triangle, t, u, v are bit-identical to the CPU
// oracle's (tests/test_gpu_mesh_parity.py).
//
// Shading kernels: ONE WARP PER PIXEL, lanes stride over the pixel's samples, exactly as the G-buffer kernels, same
// film partials / develop / adjoint kernels.  The adjoint walks the path forward (emitter-term and miss-term envmap
// gradients scattered on the way), keeps a 30-word record per scattering vertex in local memory, then walks back
// with the suffix radiance R_{k+1} to add the BSDF-weight term; material gradients of lanes that hit the same texel
// are combined with match.any + a peer reduction before they leave as REDs.
#include <cub/device/device_radix_sort.cuh>
#include "mb200_render_common.cuh"

namespace {

constexpr int kLeaf = MB200_MESH_LEAF;          // triangles per leaf
constexpr int kMaxVerts = 7;                    // scattering vertices per path (max_depth <= 8), as the oracle
constexpr int kStack = 56;
constexpr float kRayEps = 1500.f * 5.9604644775390625e-08f;
constexpr float kShadowEps = 15000.f * 5.9604644775390625e-08f;
constexpr float kInf = __builtin_huge_valf();

struct MeshView {
    const float4* tv;        // (slots, 3): (p0.xyz, tri id as int bits), (p1.xyz, 0), (p2.xyz, 0)
    const float4* tn;        // (slots, 3) vertex normals of the slot's triangle, or nullptr (flat shading frames)
    const float4* nodes;     // child-box groups, 6 float4 each
    const float* header;     // centre.xyz, radius (scene bounding sphere)
    int n_levels;            // group levels 0 .. n_levels-1 (level 0 = leaves)
    int lvl_off[MB200_MESH_MAX_LEVELS];   // first group of each level
};

struct Hit { int slot, tri; float t, u, v; };

// exact (non-contracted) vector helpers and the exact sensor ray: mb200_device.cuh (xsub3, xdot3, xcross3, xnormalize3, primary_dir)
__device__ __forceinline__ float3 cross3(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// Mesh::ray_intersect_triangle (Moeller-Trumbore), operation order of the oracle's tri_intersect
__device__ __forceinline__ bool tri_intersect(float3 p0, float3 p1, float3 p2, float3 o, float3 d, float maxt, float& tt, float& uu, float& vv) {
    const float3 e1 = xsub3(p1, p0), e2 = xsub3(p2, p0);
    const float3 pvec = xcross3(d, e2);
    const float inv_det = XDIV(1.f, xdot3(e1, pvec));
    const float3 tvec = xsub3(o, p0);
    const float u = XMUL(xdot3(tvec, pvec), inv_det);
    if (!(u >= 0.f && u <= 1.f)) return false;
    const float3 qvec = xcross3(tvec, e1);
    const float v = XMUL(xdot3(d, qvec), inv_det);
    if (!(v >= 0.f && XADD(u, v) <= 1.f)) return false;
    const float th = XMUL(xdot3(e2, qvec), inv_det);
    if (!(th >= 0.f && th <= maxt)) return false;
    tt = th; uu = u; vv = v; return true;
}

#define MB_CSWAP(i, j) { if (ct[j] < ct[i]) { float tf = ct[i]; ct[i] = ct[j]; ct[j] = tf; uint32_t tc = cc[i]; cc[i] = cc[j]; cc[j] = tc; } }

// closest hit = (smallest t, then smallest triangle id); ANY: true as soon as one triangle is hit within maxt.
// Measured alternatives (profiles/r1r_mesh_traversal_variants.log, C2m step): 4-triangle leaves 549 ms, 2-triangle 433 ms,
// 1-triangle leaves (every triangle has its own box in a level-0 group) 380 ms; a "while-while" loop that postpones the
// leaf tests until every lane holds one: 780-790 ms for all leaf sizes (lanes with long inner descents stall the rest)
template <bool ANY>
__device__ __forceinline__ bool mesh_intersect(const MeshView& M, float3 o, float3 d, float maxt, Hit& h) {
    const float3 inv = f3(__fdiv_rn(1.f, d.x), __fdiv_rn(1.f, d.y), __fdiv_rn(1.f, d.z));
    const float3 oi = f3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
    uint2 stack[kStack]; int sp = 0;
    float best = maxt; bool found = false;
    h.slot = -1; h.tri = -1; h.t = maxt; h.u = h.v = 0.f;
    uint32_t cur = (uint32_t)M.n_levels << 27;          // virtual root: its children are group 0 of the top level
    for (;;) {
        const uint32_t level = cur >> 27, idx = cur & 0x7ffffffu;
        if (level == 0) {
#pragma unroll
            for (int k = 0; k < kLeaf; ++k) {
                const int slot = (int)idx * kLeaf + k;
                const float4 q0 = __ldg(M.tv + 3 * (size_t)slot);
                const int tri = __float_as_int(q0.w);
                if (tri < 0) continue;
                const float4 q1 = __ldg(M.tv + 3 * (size_t)slot + 1), q2 = __ldg(M.tv + 3 * (size_t)slot + 2);
                float tt, uu, vv;
                if (!tri_intersect(f3(q0.x, q0.y, q0.z), f3(q1.x, q1.y, q1.z), f3(q2.x, q2.y, q2.z), o, d, best, tt, uu, vv)) continue;
                if (ANY) return true;
                if (!found || tt < h.t || (tt == h.t && tri < h.tri)) { h.slot = slot; h.tri = tri; h.t = tt; h.u = uu; h.v = vv; best = tt; found = true; }
            }
        } else {
            const float4* g = M.nodes + (size_t)(M.lvl_off[level - 1] + (int)idx) * 6;
            const float4 lx = __ldg(g), ly = __ldg(g + 1), lz = __ldg(g + 2), hx = __ldg(g + 3), hy = __ldg(g + 4), hz = __ldg(g + 5);
            const float lox[4] = {lx.x, lx.y, lx.z, lx.w}, loy[4] = {ly.x, ly.y, ly.z, ly.w}, loz[4] = {lz.x, lz.y, lz.z, lz.w};
            const float hix[4] = {hx.x, hx.y, hx.z, hx.w}, hiy[4] = {hy.x, hy.y, hy.z, hy.w}, hiz[4] = {hz.x, hz.y, hz.z, hz.w};
            float ct[4]; uint32_t cc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // slab test; a NaN (0 * inf) is ignored by fminf / fmaxf, i.e. treated conservatively
                const float ax = fmaf(lox[k], inv.x, -oi.x), bx = fmaf(hix[k], inv.x, -oi.x);
                const float ay = fmaf(loy[k], inv.y, -oi.y), by = fmaf(hiy[k], inv.y, -oi.y);
                const float az = fmaf(loz[k], inv.z, -oi.z), bz = fmaf(hiz[k], inv.z, -oi.z);
                const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
                const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) * 1.0000004f;
                const bool hit = t0 <= fminf(t1, best) && lox[k] <= hix[k];
                ct[k] = hit ? t0 : kInf;
                cc[k] = ((level - 1) << 27) | (idx * 4u + (uint32_t)k);
            }
            MB_CSWAP(0, 1) MB_CSWAP(2, 3) MB_CSWAP(0, 2) MB_CSWAP(1, 3) MB_CSWAP(1, 2)
            if (ct[0] < kInf) {
                if (ct[1] < kInf) {
                    if (ct[2] < kInf) {
                        if (ct[3] < kInf) stack[sp++] = make_uint2(cc[3], __float_as_uint(ct[3]));
                        stack[sp++] = make_uint2(cc[2], __float_as_uint(ct[2]));
                    }
                    stack[sp++] = make_uint2(cc[1], __float_as_uint(ct[1]));
                }
                cur = cc[0];
                continue;
            }
        }
        // pop the next node that can still contain a closer (or equally close) hit
        bool got = false;
        while (sp > 0) {
            const uint2 e = stack[--sp];
            if (__uint_as_float(e.y) <= best) { cur = e.x; got = true; break; }
        }
        if (!got) break;
    }
    return found;
}

struct SurfacePoint { float3 p, ng; Frame sh; };
// Mesh::compute_surface_interaction for a mesh without UVs; shading frame from interpolated vertex normals when present
__device__ __forceinline__ SurfacePoint hit_point(const MeshView& M, const Hit& h) {
    const float4 q0 = __ldg(M.tv + 3 * (size_t)h.slot), q1 = __ldg(M.tv + 3 * (size_t)h.slot + 1), q2 = __ldg(M.tv + 3 * (size_t)h.slot + 2);
    const float b1 = h.u, b2 = h.v, b0 = XSUB(XSUB(1.f, b1), b2);
    SurfacePoint s;
    s.p = f3(XFMA(q0.x, b0, XFMA(q1.x, b1, XMUL(q2.x, b2))), XFMA(q0.y, b0, XFMA(q1.y, b1, XMUL(q2.y, b2))), XFMA(q0.z, b0, XFMA(q1.z, b1, XMUL(q2.z, b2))));
    const float3 p0 = f3(q0.x, q0.y, q0.z);
    s.ng = xnormalize3(xcross3(xsub3(f3(q1.x, q1.y, q1.z), p0), xsub3(f3(q2.x, q2.y, q2.z), p0)));
    const Frame g = make_frame(s.ng);
    if (!M.tn) { s.sh = g; return s; }
    const float4 a0 = __ldg(M.tn + 3 * (size_t)h.slot), a1 = __ldg(M.tn + 3 * (size_t)h.slot + 1), a2 = __ldg(M.tn + 3 * (size_t)h.slot + 2);
    // exact chain (the oracle's hit_point, operation for operation): the shading frame turns the sampled lobe direction into the next
    // ray, so an ulp here is an ulp of the next hit — and a different radiance wherever the next vertex sits on a GGX peak
    float3 ns = f3(XFMA(a0.x, b0, XFMA(a1.x, b1, XMUL(a2.x, b2))), XFMA(a0.y, b0, XFMA(a1.y, b1, XMUL(a2.y, b2))), XFMA(a0.z, b0, XFMA(a1.z, b1, XMUL(a2.z, b2))));
    ns = xnormalize3(ns);
    // SurfaceInteraction::initialize_sh_frame: Gram-Schmidt of dp_du (= coordinate_system(n).s) against the shading normal
    const float dd = xdot3(ns, g.s);
    float3 ss = f3(XFMA(-ns.x, dd, g.s.x), XFMA(-ns.y, dd, g.s.y), XFMA(-ns.z, dd, g.s.z));
    ss = xnormalize3(ss);
    s.sh.n = ns; s.sh.s = ss; s.sh.t = xcross3(ns, ss);
    return s;
}
// SurfaceInteraction::offset_p (exact: the oracle's offset_p)
__device__ __forceinline__ float3 offset_p(float3 p, float3 n, float3 d) {
    float mag = XMUL(XADD(1.f, fmaxf(fabsf(p.x), fmaxf(fabsf(p.y), fabsf(p.z)))), kRayEps);
    if (xdot3(n, d) < 0.f) mag = -mag;
    return f3(XFMA(mag, n.x, p.x), XFMA(mag, n.y, p.y), XFMA(mag, n.z, p.z));
}
// Scene::sample_emitter_direction(test_visibility) for an envmap: ray_test towards it.p + d * 2 * max(radius, |it.p - centre|).
// Origin, direction and length of the visibility ray, bit-identical to the oracle's shadow_visible (a grazing shadow ray is a
// discrete decision).
struct ShadowRay { float3 o, d; float maxt; };
__device__ __forceinline__ ShadowRay shadow_ray(const MeshView& M, float3 p, float3 n, float3 d) {
    const float3 c = f3(__ldg(M.header), __ldg(M.header + 1), __ldg(M.header + 2));
    const float3 pc = xsub3(p, c);
    const float rad = fmaxf(__ldg(M.header + 3), XSQRT(xdot3(pc, pc)));
    const float3 target = xadd3(p, xscale3(d, XMUL(2.f, rad)));
    ShadowRay r;
    r.o = offset_p(p, n, xsub3(target, p));
    float3 dd = xsub3(target, r.o);
    const float dist = XSQRT(xdot3(dd, dd));
    r.d = xscale3(dd, XDIV(1.f, dist));
    r.maxt = XMUL(dist, XSUB(1.f, kShadowEps));
    return r;
}
__device__ __forceinline__ bool shadow_visible(const MeshView& M, float3 p, float3 n, float3 d) {
    const ShadowRay r = shadow_ray(M, p, n, d);
    Hit h;
    return !mesh_intersect<true>(M, r.o, r.d, r.maxt, h);
}

__device__ __forceinline__ Material fetch_material(const RenderParams& P, float3 p, float3 ng, long long& flat) {
    Material mt; flat = texel_index(P.cam, p);
    mt.a = f3(__ldg(P.a + 3 * flat), __ldg(P.a + 3 * flat + 1), __ldg(P.a + 3 * flat + 2));
    mt.r = __ldg(P.r + flat); mt.m = __ldg(P.m + flat);
    mt.n = (P.use_mesh_normal || !P.n_opt) ? ng : f3(__ldg(P.n_opt + 3 * flat), __ldg(P.n_opt + 3 * flat + 1), __ldg(P.n_opt + 3 * flat + 2));
    return mt;
}

// ---------------------------------------------------------------- out-of-line copies of the big shading functions
// The path kernels inline a BVH traversal loop AND a full shading pass; with every BSDF evaluation inlined (3 per vertex
// forward, 5 in the adjoint) they were 64-130 KB of SASS and 27 % of the issue stalls were instruction fetches
// (stall_no_instruction 3.7 per issue, profiles/r1x).  One shared body per function keeps the kernels inside the
// instruction cache; the calls sit in the shading pass, not in the traversal loop.
__device__ __noinline__ BsdfVal eval_brdf_ool(float3 wi, float3 wo, const Material& mt) { return eval_brdf(wi, wo, mt); }
__device__ __noinline__ BsdfSample sample_brdf_ool(float s1, float s2x, float s2y, float3 wo, const Material& mt) { return sample_brdf(s1, s2x, s2y, wo, mt, make_frame(mt.n)); }
template <bool WANT_N>
__device__ __noinline__ BsdfGrad eval_brdf_grad_ool(float3 wi, float3 wo, const Material& mt, float3 w) { return eval_brdf_grad<WANT_N>(wi, wo, mt, w); }
__device__ __noinline__ EmSample env_sample_direction_ool(const HierView& h, const EnvView& e, float u0, float u1) { return env_sample_direction(h, e, u0, u1); }
// Le(d) * MIS weight of an escaped ray (and its bilinear footprint for the adjoint)
__device__ __noinline__ float3 env_miss_ool(const HierView& h, const EnvView& e, float3 d, float prev_pdf, bool prev_delta, Bilerp& bb, float& mis) {
    float u, v; dir_to_uv(d, u, v);
    mis = mis_weight(prev_pdf, prev_delta ? 0.f : env_pdf_direction(h, e, d, u, v));
    bb = env_lookup<false>(e, u, v);
    return env_value(e, bb);
}

// ---------------------------------------------------------------- resumable traversal (one node / leaf per step)
// The same traversal as mesh_intersect, cut into steps so that a warp can stop it when too few of its lanes still hold
// a ray, let the finished lanes shade and fetch new rays, and resume — the "persistent lanes" scheme of the kernels below.
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {      // p: 32-byte aligned
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
struct Trav {
    float3 o, d, inv, oi; float best; uint32_t cur; int sp; bool active, any, found; Hit h;
};
// Traversal stack: the first kSmStack entries of a thread live in shared memory (entry e of thread t at sh[e * kThreads], bank
// conflict free), deeper ones in local memory.  Local stores are written through to L2: with the whole stack in local memory
// half of the L2 traffic of the adjoint kernel was stack + spill traffic and its L2 hit rate 68 % (profiles/r1x).
#ifndef MB200_MESH_SM_STACK_BWD
#define MB200_MESH_SM_STACK_BWD 16
#endif
#ifndef MB200_MESH_SM_STACK_FWD
#define MB200_MESH_SM_STACK_FWD 0          // forward: measured slower with a shared stack, even one that aliases the idle film staging slice (119 vs 116 ms)
#endif
template <int SM, int STRIDE = kThreads>      // STRIDE: threads that interleave in the shared area (the CTA, or one warp)
struct TStack {
    uint2* loc; uint2* sh;
    __device__ __forceinline__ void push(int& sp, uint2 e) { if (SM > 0 && sp < SM) sh[sp * STRIDE] = e; else loc[sp] = e; ++sp; }
    __device__ __forceinline__ uint2 pop(int& sp) { --sp; return (SM > 0 && sp < SM) ? sh[sp * STRIDE] : loc[sp]; }
};
__device__ __forceinline__ void trav_begin(const MeshView& M, Trav& T, float3 o, float3 d, float maxt, bool any) {
    T.o = o; T.d = d;
    T.inv = f3(__fdiv_rn(1.f, d.x), __fdiv_rn(1.f, d.y), __fdiv_rn(1.f, d.z));
    T.oi = f3(o.x * T.inv.x, o.y * T.inv.y, o.z * T.inv.z);
    T.best = maxt; T.found = false; T.any = any; T.sp = 0; T.active = true;
    T.h.slot = -1; T.h.tri = -1; T.h.t = maxt; T.h.u = T.h.v = 0.f;
    T.cur = (uint32_t)M.n_levels << 27;
}
// Called by ALL lanes of the warp (lanes without a ray do nothing): the three phases — box tests, triangle test, stack pop —
// each start converged, so e.g. the pop loop runs once per step for every lane that needs it instead of once per divergent
// path that reaches it (it ran at 3.6 of 32 lanes: profiles/r1x).
template <int LEAF_MIN = 1, int SM, int STRIDE>
__device__ __forceinline__ void trav_step(const MeshView& M, Trav& T, TStack<SM, STRIDE> stack) {
    const bool act = T.active;
    const uint32_t level = T.cur >> 27, idx = T.cur & 0x7ffffffu;
    const bool leaf = act && level == 0, node = act && level != 0;
    // the loads of BOTH kinds of step are issued before the warp splits into its leaf lanes and its node lanes, so a mixed
    // warp pays one memory round trip per step instead of two (the kernel is latency bound: profiles/r1u)
    float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = d0, d2 = d0, d3 = d0, d4 = d0, d5 = d0;
    if (act) {
        const float4* src = leaf ? M.tv + 3 * (size_t)idx * kLeaf : M.nodes + (size_t)(M.lvl_off[leaf ? 0 : level - 1] + (int)idx) * 6;
        if (leaf) { d0 = __ldg(src); d1 = __ldg(src + 1); d2 = __ldg(src + 2); }
        else {
            // a 96-byte box group is three 32-byte sectors: three 256-bit loads (sm_100 LDG.256) instead of six 128-bit ones halve the
            // L1 tag lookups of a step whose 32 lanes all read different lines
            ldg256(src, d0, d1); ldg256(src + 2, d2, d3); ldg256(src + 4, d4, d5);
        }
    }
    bool need_pop = false;
    if (node) {
        const float lox[4] = {d0.x, d0.y, d0.z, d0.w}, loy[4] = {d1.x, d1.y, d1.z, d1.w}, loz[4] = {d2.x, d2.y, d2.z, d2.w};
        const float hix[4] = {d3.x, d3.y, d3.z, d3.w}, hiy[4] = {d4.x, d4.y, d4.z, d4.w}, hiz[4] = {d5.x, d5.y, d5.z, d5.w};
        // one 32-bit key per child: entry distance with its two low mantissa bits replaced by the child number (t0 >= 0, so
        // the keys order like the distances; clearing mantissa bits only lowers t0 -> culling stays conservative); a miss is
        // +inf.  The 5-comparator network then costs two integer min/max per comparator instead of a compare and four selects.
        uint32_t key[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float ax = fmaf(lox[k], T.inv.x, -T.oi.x), bx = fmaf(hix[k], T.inv.x, -T.oi.x);
            const float ay = fmaf(loy[k], T.inv.y, -T.oi.y), by = fmaf(hiy[k], T.inv.y, -T.oi.y);
            const float az = fmaf(loz[k], T.inv.z, -T.oi.z), bz = fmaf(hiz[k], T.inv.z, -T.oi.z);
            const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
            const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) * 1.0000004f;
            const bool hit = t0 <= fminf(t1, T.best) && lox[k] <= hix[k];
            key[k] = ((hit ? __float_as_uint(t0) : 0x7f800000u) & ~3u) | (uint32_t)k;
        }
#define MB_KSWAP(i, j) { const uint32_t lo_ = min(key[i], key[j]), hi_ = max(key[i], key[j]); key[i] = lo_; key[j] = hi_; }
        MB_KSWAP(0, 1) MB_KSWAP(2, 3) MB_KSWAP(0, 2) MB_KSWAP(1, 3) MB_KSWAP(1, 2)
#undef MB_KSWAP
        const uint32_t cbase = ((level - 1) << 27) | (idx * 4u);
        if (key[0] < 0x7f800000u) {
            if (key[1] < 0x7f800000u) {
                if (key[2] < 0x7f800000u) {
                    if (key[3] < 0x7f800000u) stack.push(T.sp, make_uint2(cbase | (key[3] & 3u), key[3] & ~3u));
                    stack.push(T.sp, make_uint2(cbase | (key[2] & 3u), key[2] & ~3u));
                }
                stack.push(T.sp, make_uint2(cbase | (key[1] & 3u), key[1] & ~3u));
            }
            T.cur = cbase | (key[0] & 3u);
        } else need_pop = true;
    }
    // LEAF_MIN > 1: leaf lanes wait until at least that many of them hold a triangle (or no lane has box work left), so that the
    // ~170-instruction exact triangle test runs at more than its usual 2-4 of 32 lanes.  Measured on the traversal-only kernels
    // (profiles/r3l): 1 / 4 / 8 / 12 / 16 -> C2m 158 / 161 / 169 / 180 / 195 ms — waiting costs more than the divergence; default 1.
    bool run_leaf = true;
    if (LEAF_MIN > 1) {
        const unsigned leaf_mask = __ballot_sync(0xffffffffu, leaf), node_mask = __ballot_sync(0xffffffffu, node);
        run_leaf = __popc(leaf_mask) >= LEAF_MIN || node_mask == 0u;
    }
    if (leaf && run_leaf) {
        need_pop = true;
#pragma unroll
        for (int k = 0; k < kLeaf; ++k) {
            const int slot = (int)idx * kLeaf + k;
            const float4 q0 = k == 0 ? d0 : __ldg(M.tv + 3 * (size_t)slot);
            const int tri = __float_as_int(q0.w);
            if (tri < 0) continue;
            const float4 q1 = k == 0 ? d1 : __ldg(M.tv + 3 * (size_t)slot + 1), q2 = k == 0 ? d2 : __ldg(M.tv + 3 * (size_t)slot + 2);
            float tt, uu, vv;
            if (!tri_intersect(f3(q0.x, q0.y, q0.z), f3(q1.x, q1.y, q1.z), f3(q2.x, q2.y, q2.z), T.o, T.d, T.best, tt, uu, vv)) continue;
            if (T.any) { T.found = true; T.active = false; need_pop = false; break; }
            if (!T.found || tt < T.h.t || (tt == T.h.t && tri < T.h.tri)) { T.h.slot = slot; T.h.tri = tri; T.h.t = tt; T.h.u = uu; T.h.v = vv; T.best = tt; T.found = true; }
        }
    }
    __syncwarp();
    if (need_pop) {
        bool got = false;
        while (T.sp > 0) {
            const uint2 e = stack.pop(T.sp);
            if (__uint_as_float(e.y) <= T.best) { T.cur = e.x; got = true; break; }
        }
        if (!got) T.active = false;
    }
}
// ---------------------------------------------------------------- traversal step with DEFERRED triangle tests (wavefront kernels)
// In trav_step a lane that reaches a leaf runs the ~170-instruction exact triangle test on the spot.  Only ~7 % of a ray's steps are
// leaf steps, so in 9 of 10 warp steps SOME lane sits on a leaf and the whole warp pays for the test at 2-3 active lanes of 32 — more
// issue slots than the box tests themselves.  Here a leaf is only NOTED (its slot goes on a short per-lane list in
// shared memory) and the lane keeps walking its stack; the warp runs the triangle test when enough lanes hold one
// (MB200_WF_LEAF_THRESH), when a lane's list is full, or when no lane has box work left.  The set of triangles a ray tests can only
// grow (`best` shrinks a little later), every test is the same exact test with the same tie rule: hits are bit-identical.
// MEASURED (profiles/r6d_defer_sweep.log, all mesh parity tests green): threshold 1 / 4 / 8 / 12 / 16 / 24 lanes -> C1 30.7 / 30.0 /
// 29.8 / 29.8 / 29.8 / 29.8 ms, C2m 118.8 / 115.5 / 113.3 / 113.3 / 112.9 / 113.3 ms against 27.4 / 104.5 ms of trav_step: batching the
// triangle tests is worth 3-5 %, the second dependent load per step (leaf data after the box data) and the list traffic cost 12 %.
// The traversal is bound by its memory pipeline (divergent 32-lane loads), not by issue slots -> left OFF.
#ifndef MB200_WF_DEFER
#define MB200_WF_DEFER 0
#endif
#ifndef MB200_WF_LEAF_THRESH
#define MB200_WF_LEAF_THRESH 8
#endif
#ifndef MB200_WF_PEND
#define MB200_WF_PEND 3
#endif
constexpr uint32_t kNoNode = 0xffffffffu;
static int wf_leaf_thresh() {       // lanes that must hold an untested leaf before the warp runs the triangle test (tuning: MB200_WF_LEAF_THRESH in the environment)
    static int v = -1;
    if (v < 0) { const char* e = getenv("MB200_WF_LEAF_THRESH"); v = e ? atoi(e) : MB200_WF_LEAF_THRESH; if (v < 1) v = 1; if (v > 32) v = 32; }
    return v;
}
template <int SM, int STRIDE>
__device__ __forceinline__ void trav_step_deferred(const MeshView& M, Trav& T, TStack<SM, STRIDE> stack, uint32_t* __restrict__ pend, int& npend, int thresh) {
    static_assert(kLeaf == 1, "deferred leaf tests assume one triangle per leaf");
    const bool act = T.active;
    const bool node = act && T.cur != kNoNode && (T.cur >> 27) != 0;
    if (node) {
        const uint32_t level = T.cur >> 27, idx = T.cur & 0x7ffffffu;
        float4 d0, d1, d2, d3, d4, d5;
        const float4* src = M.nodes + (size_t)(M.lvl_off[level - 1] + (int)idx) * 6;
        ldg256(src, d0, d1); ldg256(src + 2, d2, d3); ldg256(src + 4, d4, d5);
        const float lox[4] = {d0.x, d0.y, d0.z, d0.w}, loy[4] = {d1.x, d1.y, d1.z, d1.w}, loz[4] = {d2.x, d2.y, d2.z, d2.w};
        const float hix[4] = {d3.x, d3.y, d3.z, d3.w}, hiy[4] = {d4.x, d4.y, d4.z, d4.w}, hiz[4] = {d5.x, d5.y, d5.z, d5.w};
        uint32_t key[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float ax = fmaf(lox[k], T.inv.x, -T.oi.x), bx = fmaf(hix[k], T.inv.x, -T.oi.x);
            const float ay = fmaf(loy[k], T.inv.y, -T.oi.y), by = fmaf(hiy[k], T.inv.y, -T.oi.y);
            const float az = fmaf(loz[k], T.inv.z, -T.oi.z), bz = fmaf(hiz[k], T.inv.z, -T.oi.z);
            const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
            const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) * 1.0000004f;
            const bool hit = t0 <= fminf(t1, T.best) && lox[k] <= hix[k];
            key[k] = ((hit ? __float_as_uint(t0) : 0x7f800000u) & ~3u) | (uint32_t)k;
        }
#define MB_KSWAP(i, j) { const uint32_t lo_ = min(key[i], key[j]), hi_ = max(key[i], key[j]); key[i] = lo_; key[j] = hi_; }
        MB_KSWAP(0, 1) MB_KSWAP(2, 3) MB_KSWAP(0, 2) MB_KSWAP(1, 3) MB_KSWAP(1, 2)
#undef MB_KSWAP
        const uint32_t cbase = ((level - 1) << 27) | (idx * 4u);
        T.cur = kNoNode;
        if (key[0] < 0x7f800000u) {
            if (key[1] < 0x7f800000u) {
                if (key[2] < 0x7f800000u) {
                    if (key[3] < 0x7f800000u) stack.push(T.sp, make_uint2(cbase | (key[3] & 3u), key[3] & ~3u));
                    stack.push(T.sp, make_uint2(cbase | (key[2] & 3u), key[2] & ~3u));
                }
                stack.push(T.sp, make_uint2(cbase | (key[1] & 3u), key[1] & ~3u));
            }
            T.cur = cbase | (key[0] & 3u);
        }
    }
    __syncwarp();
    // advance: every lane whose `cur` is not a box group moves leaves onto its list and pops until it holds a box group (or runs dry,
    // or holds a leaf its full list cannot take: it keeps it in `cur` until the next triangle pass has made room)
    if (act && (T.cur == kNoNode || (T.cur >> 27) == 0)) {
        uint32_t c = T.cur;
        for (;;) {
            if (c == kNoNode) {
                bool got = false;
                while (T.sp > 0) {
                    const uint2 e = stack.pop(T.sp);
                    if (__uint_as_float(e.y) <= T.best) { c = e.x; got = true; break; }
                }
                if (!got) break;
            }
            if ((c >> 27) != 0) break;
            if (npend >= MB200_WF_PEND) break;
            pend[npend * kThreads] = c; ++npend; c = kNoNode;
        }
        T.cur = c;
    }
    __syncwarp();
    const bool has_pend = act && npend > 0;
    const bool held = act && T.cur != kNoNode && (T.cur >> 27) == 0;
    const bool boxwork = act && T.cur != kNoNode && (T.cur >> 27) != 0;
    const unsigned pm = __ballot_sync(0xffffffffu, has_pend);
    const bool flush = pm != 0u && (__popc(pm) >= thresh || __any_sync(0xffffffffu, held) || !__any_sync(0xffffffffu, boxwork));
    if (flush && has_pend) {
        --npend;
        const int slot = (int)pend[npend * kThreads];
        const float4 q0 = __ldg(M.tv + 3 * (size_t)slot);
        const int tri = __float_as_int(q0.w);
        if (tri >= 0) {
            const float4 q1 = __ldg(M.tv + 3 * (size_t)slot + 1), q2 = __ldg(M.tv + 3 * (size_t)slot + 2);
            float tt, uu, vv;
            if (tri_intersect(f3(q0.x, q0.y, q0.z), f3(q1.x, q1.y, q1.z), f3(q2.x, q2.y, q2.z), T.o, T.d, T.best, tt, uu, vv)) {
                if (T.any) { T.found = true; npend = 0; T.sp = 0; T.cur = kNoNode; }
                else if (!T.found || tt < T.h.t || (tt == T.h.t && tri < T.h.tri)) { T.h.slot = slot; T.h.tri = tri; T.h.t = tt; T.h.u = uu; T.h.v = vv; T.best = tt; T.found = true; }
            }
        }
    }
    __syncwarp();
    if (act && T.cur == kNoNode && T.sp == 0 && npend == 0) T.active = false;
}
// Scene::sample_emitter_direction's visibility ray (see shadow_visible) as a resumable any-hit query
__device__ __forceinline__ void trav_begin_shadow(const MeshView& M, Trav& T, float3 p, float3 n, float3 d) {
    const ShadowRay r = shadow_ray(M, p, n, d);
    trav_begin(M, T, r.o, r.d, r.maxt, true);
}

// ---------------------------------------------------------------- forward: PathIntegrator::sample, persistent lanes
// A warp owns a POOL of up to kPool (128) paths (whole pixels: kPool / spp of them, or one kPool-sample chunk of a pixel when
// spp > kPool).  Each lane runs a small state machine — closest-hit ray in flight / shadow ray in flight / idle — and
// fetches the next path of the pool when its own ends; all rays of the warp advance together one BVH step at a time, and
// the traversal loop is left for a shading pass only when fewer than kMinActive (8) lanes still hold a ray.  (One pixel's 32
// samples per warp pass, every lane waiting for the slowest ray of every bounce, ran at 6.2 of 32 threads per
// instruction: profiles/r1q.)  Radiance of finished paths is parked in shared memory and the film taps are reduced per
// pixel afterwards in sample order, so the image is bitwise independent of the order in which lanes finished.
// resident CTAs per SM the path kernels are compiled for (register cap 65536 / (256 N)).  The kernels are latency bound
// (long-scoreboard stalls on BVH / triangle loads, 3.6 warps per scheduler at 128 registers: profiles/r1u), so trading a few
// spills for occupancy pays: C2m forward 175 / 154 / 157 ms and adjoint 182 / 173 / 165 ms at 2 / 3 / 4 CTAs (profiles/r1v).
#ifndef MB200_MESH_MIN_BLOCKS_FWD
#define MB200_MESH_MIN_BLOCKS_FWD 4
#endif
#ifndef MB200_MESH_MIN_BLOCKS_BWD
#define MB200_MESH_MIN_BLOCKS_BWD 4
#endif
#ifndef MB200_MESH_POOL_FWD
#define MB200_MESH_POOL_FWD 128          // paths per warp pool, forward (16 B of shared memory each: the pool competes with L1)
#endif
#ifndef MB200_MESH_POOL_BWD
#define MB200_MESH_POOL_BWD 1024         // adjoint: the pool is only an index range (no film reduction, no shared memory)
#endif
#ifndef MB200_MESH_MIN_ACTIVE_FWD
#define MB200_MESH_MIN_ACTIVE_FWD 8      // leave the traversal loop for a shading pass when fewer lanes than this hold a ray
#endif
#ifndef MB200_MESH_MIN_ACTIVE_BWD
#define MB200_MESH_MIN_ACTIVE_BWD 8
#endif
constexpr int kPool = MB200_MESH_POOL_FWD;
constexpr int kPoolBwd = MB200_MESH_POOL_BWD;
constexpr int kMinActive = MB200_MESH_MIN_ACTIVE_FWD;
constexpr int kMinActiveBwd = MB200_MESH_MIN_ACTIVE_BWD;
enum { ST_IDLE = 0, ST_CLOSEST = 1, ST_SHADOW = 2, ST_DONE = 3 };

template <int FILTER, bool AD_W, bool TRANS = false>
__global__ void __launch_bounds__(kThreads, MB200_MESH_MIN_BLOCKS_FWD) mesh_fwd_kernel(const __grid_constant__ RenderParams P, const __grid_constant__ MeshView M) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* pool = reinterpret_cast<float4*>(smem) + warp * kPool;
    float* rec = smem + kWarpsPerBlock * kPool * 4 + warp * 32 * kRecStride;        // gaussian film only
    uint2 stack_loc[kStack];
    TStack<MB200_MESH_SM_STACK_FWD, 32> stack;
    stack.loc = stack_loc;
    // the first entries of the traversal stack share THIS WARP's film staging slice: a warp's stack is empty whenever it reduces its film
    static_assert(MB200_MESH_SM_STACK_FWD * 8 * 32 <= 32 * kRecStride * 4, "shared stack must fit the warp's staging slice");
    stack.sh = reinterpret_cast<uint2*>(rec) + lane;
    const int npix = P.prows * P.W;
    const int chunks = (P.spp + kPool - 1) / kPool;              // kPool-sample chunks per pixel (1 unless spp > kPool)
    const int ppp = chunks > 1 ? 1 : kPool / P.spp;              // whole pixels per pool
    const int npools = (npix + ppp - 1) / ppp;
    const int ti = lane % 5, tj = lane / 5;
    const int max_verts = min(P.max_depth - 1, kMaxVerts);
    const float3 cam_o = f3(P.cam.c2w[3], P.cam.c2w[7], P.cam.c2w[11]);
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int pool_id = blockIdx.x * kWarpsPerBlock + warp; pool_id < npools; pool_id += gridDim.x * kWarpsPerBlock) {
        const int pix0 = pool_id * ppp, pixn = min(ppp, npix - pix0);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < chunks; ++c) {
            const int s_lo = c * kPool, s_n = min(P.spp - s_lo, kPool);
            const int pool_n = pixn * s_n;
            // ================================================= trace the pool
            int next_q = 0, stage = ST_IDLE, q = 0, nv = 0;
            Trav T; T.active = false;
            Pcg32 rng;
            float3 beta = f3(1.f, 1.f, 1.f), L = f3(0.f, 0.f, 0.f), cem = f3(0.f, 0.f, 0.f), nro = f3(0.f, 0.f, 0.f), nrd = f3(0.f, 0.f, 1.f);
            float prev_pdf = 1.f; bool prev_delta = true, dead = false;
            for (;;) {
                // ---- shading pass: lanes whose ray has finished
                if (!T.active && stage != ST_DONE) {
                    bool path_end = false, start_next = false;
                    if (stage == ST_CLOSEST) {
                        if (!T.found) {                                              // direct emission: the environment
                            if (prev_pdf > 0.f) {
                                Bilerp bb; float mis; const float3 le = env_miss_ool(P.hier, P.env, T.d, prev_pdf, prev_delta, bb, mis);
                                L = L + beta * le * mis;
                            }
                            path_end = true;
                        } else if (nv >= max_verts) {                                // depth + 1 >= max_depth
                            path_end = true;
                        } else {
                            const SurfacePoint sp = hit_point(M, T.h);
                            const float3 view = f3(-T.d.x, -T.d.y, -T.d.z);
                            long long flat; const Material mt = fetch_material(P, sp.p, sp.ng, flat);
                            TransMat tm; if (TRANS) tm = trans_fetch(P.cam, P.trans, flat, view, sp.ng, sp.p);
                            const float uex = rng.next_float(), uey = rng.next_float();
                            const EmSample em = env_sample_direction_ool(P.hier, P.env, uex, uey);
                            const float s1 = rng.next_float();
                            const float s2x = rng.next_float(), s2y = rng.next_float();
                            cem = f3(0.f, 0.f, 0.f);
                            bool need_shadow = false;                                // an emitter sample below the horizon has f = 0 exactly:
                            if (em.pdf != 0.f) {                                     // its visibility cannot change the result, the ray is not traced
                                const BsdfVal fv = TRANS ? trans_eval_brdf(em.d, view, mt, tm, P.trans) : eval_brdf_ool(em.d, view, mt);
                                cem = beta * fv.f * env_value(P.env, em.b) * (mis_weight(em.pdf, fv.pdf) / em.pdf);
                                need_shadow = fmax3(fv.f.x, fv.f.y, fv.f.z) > 0.f;
                            }
                            const BsdfSample bs = TRANS ? trans_sample_brdf(s1, s2x, s2y, view, mt, tm, P.trans, make_frame(mt.n))
                                                        : sample_brdf_ool(s1, s2x, s2y, view, mt);
                            const float3 d_bs = (P.flags & MB200_FLAG_WO_WORLD_QUIRK) ? to_world(sp.sh, bs.wi) : bs.wi;     // mi_plugin.py:1444
                            float3 w = bs.weight;
                            if (AD_W) {
                                const BsdfVal b2 = eval_brdf_ool(d_bs, view, mt);
                                if (b2.pdf > 0.f) w = b2.f * (1.f / b2.pdf);
                            }
                            nro = offset_p(sp.p, sp.ng, d_bs); nrd = d_bs;
                            beta = beta * w; prev_pdf = bs.pdf; prev_delta = false; nv += 1;
                            rng.next_float();                                        // russian-roulette draw (rr_depth 5: never applied)
                            dead = fmax3(beta.x, beta.y, beta.z) == 0.f;
                            if (need_shadow) { trav_begin_shadow(M, T, sp.p, sp.ng, em.d); stage = ST_SHADOW; }
                            else if (dead) path_end = true;
                            else start_next = true;
                        }
                    } else if (stage == ST_SHADOW) {
                        if (!T.found) L = L + cem;
                        if (dead) path_end = true; else start_next = true;
                    }
                    if (start_next) { trav_begin(M, T, nro, nrd, kInf, false); stage = ST_CLOSEST; }
                    if (path_end) { pool[q] = make_float4(L.x, L.y, L.z, 0.f); stage = ST_IDLE; }
                }
                // ---- fetch: idle lanes take the next paths of the pool
                {
                    const bool want = stage == ST_IDLE;
                    const unsigned wm = __ballot_sync(0xffffffffu, want);
                    if (want) {
                        q = next_q + __popc(wm & lt_mask);
                        if (q < pool_n) {
                            const int pix = pix0 + q / s_n, s = s_lo + q % s_n;
                            const int py = P.prow0 + pix / P.W, px = pix % P.W;
                            rng.seed(P.seed, (uint32_t)(py * P.W + px) * (uint32_t)P.spp + (uint32_t)s);
                            const float jx = rng.next_float(), jy = rng.next_float();
                            beta = f3(1.f, 1.f, 1.f); L = f3(0.f, 0.f, 0.f); prev_pdf = 1.f; prev_delta = true; nv = 0;
                            trav_begin(M, T, cam_o, primary_dir(P.cam, XADD((float)px, jx), XADD((float)py, jy)), kInf, false);
                            stage = ST_CLOSEST;
                        } else stage = ST_DONE;
                    }
                    next_q += __popc(wm);
                }
                // ---- traversal: all rays of the warp, one BVH step at a time
                if (!__ballot_sync(0xffffffffu, T.active)) break;                    // every lane is ST_DONE
                for (;;) {
                    trav_step(M, T, stack);
                    const int na = __popc(__ballot_sync(0xffffffffu, T.active));
                    if (na == 0) break;
                    if (na < kMinActive && na < __popc(__ballot_sync(0xffffffffu, stage != ST_DONE))) break;
                }
            }
            __syncwarp();
            // ================================================= film: the pool's pixels, samples in order
            for (int pi = 0; pi < pixn; ++pi) {
                const int pix = pix0 + pi;
                const int py = P.prow0 + pix / P.W, px = pix % P.W;
                const int gpix = py * P.W + px;
                if (chunks == 1) acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int b0 = 0; b0 < s_n; b0 += 32) {
                    const int sl = b0 + lane;                                        // sample index inside the chunk
                    const bool act = sl < s_n;
                    float3 Ls = f3(0.f, 0.f, 0.f);
                    if (act) { const float4 r4 = pool[pi * s_n + sl]; Ls = f3(r4.x, r4.y, r4.z); }
                    if (FILTER == MB200_FILTER_GAUSSIAN) {
                        float wx[5], wy[5];
                        if (act) {
                            Pcg32 r2; r2.seed(P.seed, (uint32_t)gpix * (uint32_t)P.spp + (uint32_t)(s_lo + sl));
                            const float jx = r2.next_float(), jy = r2.next_float();
                            film_taps(jx, wx); film_taps(jy, wy);
                        } else {
#pragma unroll
                            for (int i = 0; i < 5; ++i) { wx[i] = 0.f; wy[i] = 0.f; }
                        }
                        float4* r4 = reinterpret_cast<float4*>(rec + lane * kRecStride);
                        r4[0] = make_float4(wx[0], wx[1], wx[2], wx[3]);
                        r4[1] = make_float4(wx[4], wy[0], wy[1], wy[2]);
                        r4[2] = make_float4(wy[3], wy[4], 0.f, 0.f);
                        r4[3] = make_float4(Ls.x, Ls.y, Ls.z, 1.f);
                        __syncwarp();
                        if (lane < MB200_FILM_TAPS) {
                            const float* rt = rec + ti; const float* ru = rec + 5 + tj;
#pragma unroll 8
                            for (int k = 0; k < 32; ++k) {
                                const float w = rt[k * kRecStride] * ru[k * kRecStride];
                                const float4 l4 = *reinterpret_cast<const float4*>(rec + k * kRecStride + 12);
                                acc.x = fmaf(w, l4.x, acc.x); acc.y = fmaf(w, l4.y, acc.y); acc.z = fmaf(w, l4.z, acc.z); acc.w += w;
                            }
                        }
                        __syncwarp();
                    } else {
                        acc.x += Ls.x; acc.y += Ls.y; acc.z += Ls.z;
                    }
                }
                if (c == chunks - 1) {
                    if (FILTER == MB200_FILTER_GAUSSIAN) {
                        if (lane < MB200_FILM_TAPS)
                            reinterpret_cast<float4*>(P.partials)[(size_t)pix * MB200_FILM_TAPS + lane] = acc;
                    } else {
                        float4 t = acc;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
                            t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
                            t.z += __shfl_xor_sync(0xffffffffu, t.z, o);
                        }
                        if (lane == 0) reinterpret_cast<float4*>(P.partials)[pix] = make_float4(t.x, t.y, t.z, (float)P.spp);
                    }
                }
            }
            __syncwarp();
        }
    }
}
inline size_t mesh_fwd_smem(int filter) {
    // pool records + the film staging area (gaussian film, or whenever the shared stack aliases it)
    return (size_t)kWarpsPerBlock * ((size_t)kPool * 16 + ((filter == MB200_FILTER_GAUSSIAN || MB200_MESH_SM_STACK_FWD > 0) ? 32 * kRecStride * 4 : 0));
}
template <typename K>
inline void launch_mesh_fwd(K kernel, int filter, int grid, cudaStream_t st, const RenderParams& P, const MeshView& M) {
    const size_t bytes = mesh_fwd_smem(filter);
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    kernel<<<grid, kThreads, bytes, st>>>(P, M);
}

// ---------------------------------------------------------------- forward, wavefront formulation
// The same PathIntegrator::sample, cut into kernels at its rays (Laine, Karras, Aila 2013): per batch of <= kWfBatch paths
//   wf_gen      primary rays + path state
//   wf_trace    closest hits of the queued rays      (traversal only: ~48 registers, no spills, every lane always holds a ray —
//   wf_shade    one thread per live path: miss / vertex; queues the shadow ray and the continuation ray      lanes refill from a global counter)
//   wf_trace    any-hit of the shadow rays; adds the parked emitter term of the unoccluded ones
//   ... repeated max_depth-1 times ...     wf_film: per pixel, samples in order -> the same film partials as mesh_fwd_kernel.
// Path state lives in a caller-provided scratch buffer (mb200_mesh_fwd_wf_scratch_bytes), SoA in float4s: 156 bytes per path.
// Measured against the one-kernel persistent-lane formulation above (profiles/r3f): C2m forward 116 -> 79 ms, C1 60 -> 44 ms;
// in the launch list the two traversal kernels are 93 % of the time and all shading 7 %.
#ifndef MB200_WF_BATCH_LOG2
#define MB200_WF_BATCH_LOG2 24          // paths per batch (156 B of scratch each): 2^20 / 2^22 / 2^24 -> 120 / 94 / 87 ms C2m forward (fewer, fuller launches)
#endif
#ifndef MB200_WF_BINS
#define MB200_WF_BINS 1
#endif
#ifndef MB200_WF_LEAF_MIN
#define MB200_WF_LEAF_MIN 1
#endif
#ifndef MB200_WF_TRACE_BLOCKS
#define MB200_WF_TRACE_BLOCKS 4          // CTAs per SM of the traversal kernels: 64 registers, no spills (5: 94 ms, 4: 84 ms, 3: 87 ms, 6: 122 ms)
#endif
constexpr int kWfBatch = 1 << MB200_WF_BATCH_LOG2;
#ifndef MB200_WF_SM_STACK
#define MB200_WF_SM_STACK 16        // measured (profiles/r5w): 0 / 8 / 12 / 16 / 24 entries -> C2m 132.8 / 133.2 / 128.2 / 127.6 / 129.3 ms, C1 37.1 / 37.2 / 36.3 / 36.0 / 36.7 ms
#endif
#ifndef MB200_WF_CHUNK
#define MB200_WF_CHUNK 64               // queue entries a traversal warp draws per global atomic
#endif
constexpr int kWfChunk = MB200_WF_CHUNK;
// Ray queues can be binned by direction octant (MB200_WF_BINS = 8: bin b of queue q at q + b * nb, counts in
// counters[8 * (1 + queue) + b]; the traversal kernel drains bin after bin, so the lanes of a warp hold rays of one octant).
// Measured (profiles/r3p): C2m 155 -> 167 ms, C1 42.9 -> 46.9 ms — queue order = pixel order keeps the ORIGINS of a warp's rays
// together, which is worth more than like directions; default 1 bin.
constexpr int kBins = MB200_WF_BINS;
enum { CNT_A = 8, CNT_B = 16, CNT_S = 24 };
struct WfBuf {
    float4 *ray_o, *ray_d, *hit, *sray_o, *sray_d, *beta, *L, *cem; uint4* rng;
    uint32_t *qa, *qb, *qs; uint32_t* counters;          // counters: [0] n(qa) [1] n(qb) [2] n(qs) [3] fetch cursor
    // adjoint only: film cotangent, pending envmap scatter (cotangent + bilinear footprint), visibility, vertex records
    float4 *dl, *scat, *pbw, *vrec; uint32_t* vis; long long nb; int max_verts;
};
inline size_t wf_scratch_bytes(long long nb) { return (size_t)nb * (9 * 16 + 3 * 4 * kBins) + 256; }
inline WfBuf wf_carve(void* scratch, long long nb) {
    WfBuf B; char* p = (char*)scratch;
    B.counters = (uint32_t*)p; p += 256;
    B.ray_o = (float4*)p; p += nb * 16; B.ray_d = (float4*)p; p += nb * 16; B.hit = (float4*)p; p += nb * 16;
    B.sray_o = (float4*)p; p += nb * 16; B.sray_d = (float4*)p; p += nb * 16; B.beta = (float4*)p; p += nb * 16;
    B.L = (float4*)p; p += nb * 16; B.cem = (float4*)p; p += nb * 16; B.rng = (uint4*)p; p += nb * 16;
    B.qa = (uint32_t*)p; p += nb * 4 * kBins; B.qb = (uint32_t*)p; p += nb * 4 * kBins; B.qs = (uint32_t*)p; p += nb * 4 * kBins;
    B.dl = B.scat = B.pbw = B.vrec = nullptr; B.vis = nullptr; B.nb = nb; B.max_verts = 0;
    return B;
}
constexpr int kVRecF4 = 8;                               // float4s per vertex record
inline size_t wf_bwd_scratch_bytes(long long nb, int max_verts) {
    return wf_scratch_bytes(nb) + (size_t)nb * (3 * 16 + 4 + (size_t)max_verts * kVRecF4 * 16) + 256;
}
inline WfBuf wf_carve_bwd(void* scratch, long long nb, int max_verts) {
    WfBuf B = wf_carve(scratch, nb);
    char* p = (char*)scratch + ((wf_scratch_bytes(nb) + 255) & ~(size_t)255);
    B.dl = (float4*)p; p += nb * 16; B.scat = (float4*)p; p += nb * 16; B.pbw = (float4*)p; p += nb * 16;
    B.vrec = (float4*)p; p += (size_t)nb * max_verts * kVRecF4 * 16; B.vis = (uint32_t*)p;
    B.max_verts = max_verts;
    return B;
}
// vertex record k of path pid: 8 float4s, each array coalesced over pid
__device__ __forceinline__ float4* vrec_at(const WfBuf& B, int k, int j, uint32_t pid) { return B.vrec + ((size_t)(k * kVRecF4 + j) * B.nb + pid); }
__device__ __forceinline__ uint32_t wf_append(uint32_t* counter, bool want) {      // warp-aggregated queue append (returns the slot)
    const unsigned m = __ballot_sync(0xffffffffu, want);
    uint32_t base = 0;
    const int lane = threadIdx.x & 31;
    if (lane == 0 && m) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}
struct QView { uint32_t pre[kBins + 1]; };
__device__ __forceinline__ QView wf_qview(const uint32_t* counts) {
    QView v; v.pre[0] = 0;
#pragma unroll
    for (int b = 0; b < kBins; ++b) v.pre[b + 1] = v.pre[b] + counts[b];
    return v;
}
__device__ __forceinline__ uint32_t wf_qget(const uint32_t* q, const QView& v, uint32_t i, long long nb) {
    int b = 0;
#pragma unroll
    for (int k = 1; k < kBins; ++k) b += i >= v.pre[k] ? 1 : 0;
    return q[(size_t)b * nb + (i - v.pre[b])];
}
__device__ __forceinline__ int wf_octant(float dx, float dy, float dz) {
    return kBins == 8 ? ((dx < 0.f ? 1 : 0) | (dy < 0.f ? 2 : 0) | (dz < 0.f ? 4 : 0)) : 0;
}
__device__ __forceinline__ void wf_append_bin(uint32_t* counts, uint32_t* q, long long nb, bool want, int bin, uint32_t pid) {
#pragma unroll
    for (int b = 0; b < kBins; ++b) {
        const bool w = want && bin == b;
        const uint32_t slot = wf_append(counts + b, w);
        if (w) q[(size_t)b * nb + slot] = pid;
    }
}
__global__ void wf_gen_kernel(const __grid_constant__ RenderParams P, WfBuf B, long long pix0, int nb) {
    const float3 cam_o = f3(P.cam.c2w[3], P.cam.c2w[7], P.cam.c2w[11]);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nb; p += gridDim.x * blockDim.x) {
        const long long pix = pix0 + p / P.spp; const int s = p % P.spp;
        const int py = P.prow0 + (int)(pix / P.W), px = (int)(pix % P.W);
        Pcg32 rng; rng.seed(P.seed, (uint32_t)(py * P.W + px) * (uint32_t)P.spp + (uint32_t)s);
        const float jx = rng.next_float(), jy = rng.next_float();
        const float3 d = primary_dir(P.cam, XADD((float)px, jx), XADD((float)py, jy));
        // (.w of the two ray records: the film position of the primary ray, read by wf_primary_kernel)
        B.ray_o[p] = make_float4(cam_o.x, cam_o.y, cam_o.z, XADD((float)px, jx)); B.ray_d[p] = make_float4(d.x, d.y, d.z, XADD((float)py, jy));
        B.beta[p] = make_float4(1.f, 1.f, 1.f, 1.f);                               // w = prev_pdf
        B.L[p] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x100));               // w = nv | prev_delta << 8
        B.rng[p] = make_uint4((uint32_t)rng.state, (uint32_t)(rng.state >> 32), (uint32_t)rng.inc, (uint32_t)(rng.inc >> 32));
        B.qa[p] = (uint32_t)p;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < 64; ++i) B.counters[i] = 0;
        B.counters[CNT_A] = (uint32_t)nb;                                          // primary rays: one bin (they are coherent anyway)
    }
}
// ---------------------------------------------------------------- primary-visibility index
// Primary rays all leave one point, so which triangles a ray CAN hit is a 2-D question: those whose projection onto the film contains
// the ray's film position.  The index bins every triangle into the pixels its projected bounding box (dilated by kPIdxEps pixels:
// far more than the float rounding of primary_dir / Moeller-Trumbore can move a hit) overlaps; a primary ray then runs the SAME exact
// triangle test on the <= kPIdxK candidates of its pixel whose box contains its film position — the closest hit with the same tie
// rule, hence bit-identical to the BVH traversal (and to brute force), at ~1/5 of its cost: for the reference's depth-map meshes
// (vertex k <-> pixel k, mesh_recon.py:184-258) a pixel sees 8-18 candidates and a ray tests 2-4 of them.  Pixels with more than
// kPIdxK candidates, and meshes with a triangle at or behind the camera plane or one that covers more than 64 pixels, fall back to the
// BVH (per pixel / for the whole image).  Built once per (mesh, camera): mb200_mesh_primary_index_build.
constexpr int kPIdxK = 20;
constexpr float kPIdxEps = 0.01f;
struct PIdxView { int* valid; int* counts; int* lists; float4* bbox; };
__host__ __device__ inline size_t pidx_off_counts() { return 256; }
inline size_t pidx_bytes(int H, int W, int n_slots) {
    return 256 + (((size_t)H * W * (1 + kPIdxK) * 4 + 255) & ~(size_t)255) + (size_t)n_slots * 16;
}
inline PIdxView pidx_view(void* buf, int H, int W) {
    PIdxView v; char* b = (char*)buf;
    v.valid = (int*)b; v.counts = (int*)(b + 256); v.lists = v.counts + (size_t)H * W;
    v.bbox = (float4*)(b + 256 + (((size_t)H * W * (1 + kPIdxK) * 4 + 255) & ~(size_t)255));
    return v;
}
__global__ void pidx_bin_kernel(const __grid_constant__ MeshView M, const __grid_constant__ CamView cam, PIdxView I, int n_slots) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const float4 q0 = __ldg(M.tv + 3 * (size_t)slot);
    float4 bb = make_float4(1.f, 1.f, 0.f, 0.f);                     // empty
    if (__float_as_int(q0.w) >= 0) {
        const float4 q1 = __ldg(M.tv + 3 * (size_t)slot + 1), q2 = __ldg(M.tv + 3 * (size_t)slot + 2);
        const double vx[3] = {q0.x, q1.x, q2.x}, vy[3] = {q0.y, q1.y, q2.y}, vz[3] = {q0.z, q1.z, q2.z};
        const double ox = cam.c2w[3], oy = cam.c2w[7], oz = cam.c2w[11];
        const double t = cam.tan_half_fov_x, aspect = (double)cam.W / (double)cam.H;
        double x0 = 1e30, y0 = 1e30, x1 = -1e30, y1 = -1e30; bool ok = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double dx = vx[k] - ox, dy = vy[k] - oy, dz = vz[k] - oz;
            // camera-local coordinates: columns of the rotation part of cam_to_world (primary_dir maps local (lx, ly, 1) through its rows)
            const double lx = cam.c2w[0] * dx + cam.c2w[4] * dy + cam.c2w[8] * dz;
            const double ly = cam.c2w[1] * dx + cam.c2w[5] * dy + cam.c2w[9] * dz;
            const double lz = cam.c2w[2] * dx + cam.c2w[6] * dy + cam.c2w[10] * dz;
            if (!(lz > 1e-6 * (fabs(lx) + fabs(ly) + fabs(lz)))) { ok = false; break; }
            const double sx = 0.5 * cam.W * (1.0 - lx / (lz * t)), sy = 0.5 * cam.H * (1.0 - ly * aspect / (lz * t));
            x0 = fmin(x0, sx); x1 = fmax(x1, sx); y0 = fmin(y0, sy); y1 = fmax(y1, sy);
        }
        if (!ok) { *I.valid = 0; }
        else {
            x0 -= kPIdxEps; y0 -= kPIdxEps; x1 += kPIdxEps; y1 += kPIdxEps;
            bb = make_float4((float)x0, (float)y0, (float)x1, (float)y1);
            const int px0 = max(0, (int)floor(x0)), px1 = min(cam.W - 1, (int)floor(x1));
            const int py0 = max(0, (int)floor(y0)), py1 = min(cam.H - 1, (int)floor(y1));
            if (px1 >= px0 && py1 >= py0) {
                if ((long long)(px1 - px0 + 1) * (py1 - py0 + 1) > 64) *I.valid = 0;
                else
                    for (int y = py0; y <= py1; ++y)
                        for (int x = px0; x <= px1; ++x) {
                            const int pix = y * cam.W + x;
                            const int pos = atomicAdd(I.counts + pix, 1);
                            if (pos < kPIdxK) I.lists[(size_t)pix * kPIdxK + pos] = slot;
                        }
            }
        }
    }
    I.bbox[slot] = bb;
}
// closest hit of the primary rays of a batch through the index; rays of pixels the index cannot serve go to the fallback queue
__global__ void __launch_bounds__(256) wf_primary_kernel(const __grid_constant__ RenderParams P, const __grid_constant__ MeshView M, WfBuf B, PIdxView I,
                                                         long long pix0, int nb, uint32_t* __restrict__ qfall, uint32_t* fall_count) {
    const int valid = *I.valid;
    const uint32_t nthreads = gridDim.x * blockDim.x;
    for (uint32_t p0 = blockIdx.x * blockDim.x; p0 < (uint32_t)nb; p0 += nthreads) {      // warp-uniform trip count (queue appends are warp-wide)
        const uint32_t p = p0 + threadIdx.x;
        bool fall = false;
        if (p < (uint32_t)nb) {
            const long long pix = pix0 + p / P.spp;
            const int py = P.prow0 + (int)(pix / P.W), px = (int)(pix % P.W);
            const int gpix = py * P.W + px;
            const int cnt = valid ? __ldg(I.counts + gpix) : kPIdxK + 1;
            if (cnt > kPIdxK) fall = true;
            else {
                const float4 o4 = B.ray_o[p], d4 = B.ray_d[p];
                const float3 o = f3(o4.x, o4.y, o4.z), d = f3(d4.x, d4.y, d4.z);
                const float sx = o4.w, sy = d4.w;
                float best = kInf; int bslot = -1, btri = -1; float bt = kInf, bu = 0.f, bv = 0.f;
                const int* lst = I.lists + (size_t)gpix * kPIdxK;
                for (int i = 0; i < cnt; ++i) {
                    const int slot = __ldg(lst + i);
                    const float4 bb = __ldg(I.bbox + slot);
                    if (sx < bb.x || sx > bb.z || sy < bb.y || sy > bb.w) continue;
                    const float4 q0 = __ldg(M.tv + 3 * (size_t)slot), q1 = __ldg(M.tv + 3 * (size_t)slot + 1), q2 = __ldg(M.tv + 3 * (size_t)slot + 2);
                    float tt, uu, vv;
                    if (!tri_intersect(f3(q0.x, q0.y, q0.z), f3(q1.x, q1.y, q1.z), f3(q2.x, q2.y, q2.z), o, d, best, tt, uu, vv)) continue;
                    const int tri = __float_as_int(q0.w);
                    if (bslot < 0 || tt < bt || (tt == bt && tri < btri)) { bslot = slot; btri = tri; bt = tt; bu = uu; bv = vv; best = tt; }
                }
                B.hit[p] = make_float4(__int_as_float(bslot), bt, bu, bv);
            }
        }
        const uint32_t slot_q = wf_append(fall_count, fall);
        if (fall) qfall[slot_q] = p;
    }
}

// rays of queue q[0 .. *count): MODE 0 = closest hit -> hit record; 1 = any hit, L += cem when unoccluded (forward);
// 2 = any hit, visibility flag (adjoint)
template <int MODE>
__global__ void __launch_bounds__(kThreads, MB200_WF_TRACE_BLOCKS) wf_trace_kernel(const __grid_constant__ MeshView M, WfBuf B, const uint32_t* __restrict__ q,
                                                                const uint32_t* __restrict__ count, uint32_t* cursor, int leaf_thresh) {
    constexpr bool ANY = MODE != 0;
    uint2 stack_loc[kStack];
    // the first MB200_WF_SM_STACK entries of a lane's traversal stack in shared memory (entry e of thread t at sh[e * 256 + t]: two
    // wavefronts per push / pop whatever the lanes' stack depths), deeper ones in local memory (one 128-byte line per DISTINCT depth)
    __shared__ uint2 s_stack[MB200_WF_SM_STACK > 0 ? MB200_WF_SM_STACK * kThreads : 1];
    TStack<MB200_WF_SM_STACK> stack; stack.loc = stack_loc; stack.sh = s_stack + threadIdx.x;
#if MB200_WF_DEFER
    __shared__ uint32_t s_pend[MB200_WF_PEND * kThreads];            // leaves a lane has reached and not yet tested (trav_step_deferred)
    int npend = 0;
#endif
    const QView qv = wf_qview(count);
    const uint32_t n = qv.pre[kBins];
    Trav T; T.active = false; uint32_t pid = 0;
    // The warp draws rays from the queue in CHUNKS of kWfChunk consecutive entries (one global atomic per chunk) and hands them to
    // its lanes from registers.  (One atomicAdd on the cursor per loop iteration — some lane finishes a ray in almost every
    // iteration — put a ~600-cycle L2 round trip on the critical path of every traversal step: profiles/r5o.)
    uint32_t pool = 0, pool_end = 0; bool drained = false;          // warp-uniform
    const int lane = threadIdx.x & 31;
#ifdef MB200_TRAV_STATS
    unsigned long long st_node = 0, st_leaf = 0, st_iter = 0;       // tuning build: box steps / triangle tests per lane, loop iterations per warp
    struct StatFlush { unsigned long long &a, &b, &c; uint32_t* cnt; int mode; __device__ ~StatFlush() {
        atomicAdd((unsigned long long*)(cnt + 40 + 8 * 0) + 3 * mode + 0, a); atomicAdd((unsigned long long*)(cnt + 40) + 3 * mode + 1, b);
        atomicAdd((unsigned long long*)(cnt + 40) + 3 * mode + 2, c); } } st_flush{st_node, st_leaf, st_iter, B.counters, MODE};
#endif
    for (;;) {
        // lanes without a ray take the next ones of the warp's chunk
        const unsigned wm = __ballot_sync(0xffffffffu, !T.active);
        if (wm && !drained) {
            if (pool == pool_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, (uint32_t)kWfChunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n) drained = true;
                else { pool = base; pool_end = min(base + (uint32_t)kWfChunk, n); }
            }
            if (!drained) {
                const uint32_t rank = (uint32_t)__popc(wm & ((1u << lane) - 1u));
                const uint32_t i = pool + rank;
                if (!T.active && i < pool_end) {
                    pid = wf_qget(q, qv, i, B.nb);
                    const float4 o = ANY ? B.sray_o[pid] : B.ray_o[pid], d = ANY ? B.sray_d[pid] : B.ray_d[pid];
                    trav_begin(M, T, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), ANY ? d.w : kInf, ANY);
                }
                pool = min(pool + (uint32_t)__popc(wm), pool_end);
            }
        }
        if (!__ballot_sync(0xffffffffu, T.active)) { if (drained) break; else continue; }
        const bool was = T.active;
#ifdef MB200_TRAV_STATS
        if (was) { if ((T.cur >> 27) == 0) ++st_leaf; else ++st_node; }
        if (lane == 0) ++st_iter;
#endif
#if MB200_WF_DEFER
        trav_step_deferred(M, T, stack, s_pend + threadIdx.x, npend, leaf_thresh);
#else
        trav_step<MB200_WF_LEAF_MIN>(M, T, stack);
#endif
        if (was && !T.active) {                                                    // this lane's ray has finished
            if (MODE == 1) {
                if (!T.found) { float4 L = B.L[pid]; const float4 c = B.cem[pid]; L.x += c.x; L.y += c.y; L.z += c.z; B.L[pid] = L; }
            } else if (MODE == 2) {
                B.vis[pid] = T.found ? 0u : 1u;
            } else {
                B.hit[pid] = make_float4(__int_as_float(T.found ? T.h.slot : -1), T.h.t, T.h.u, T.h.v);
            }
        }
    }
}
template <bool AD_W, bool TRANS>
__global__ void __launch_bounds__(kThreads, 2) wf_shade_kernel(const __grid_constant__ RenderParams P, const __grid_constant__ MeshView M, WfBuf B,
                                                                const uint32_t* __restrict__ qin, uint32_t* __restrict__ qout, int cin, int cout) {
    const QView qv = wf_qview(B.counters + cin);
    const uint32_t n = qv.pre[kBins];
    const int max_verts = min(P.max_depth - 1, kMaxVerts);
    const uint32_t nthreads = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += nthreads) {           // warp-uniform trip count (queue appends are warp-wide)
        const uint32_t i = i0 + threadIdx.x;
        const bool have = i < n;
        bool cont = false, shadow = false; uint32_t pid = 0; int obin = 0, sbin = 0;
        if (have) {
            pid = wf_qget(qin, qv, i, B.nb);
            const float4 hr = B.hit[pid], rd4 = B.ray_d[pid];
            const float3 rd = f3(rd4.x, rd4.y, rd4.z);
            float4 b4 = B.beta[pid], L4 = B.L[pid];
            float3 beta = f3(b4.x, b4.y, b4.z), L = f3(L4.x, L4.y, L4.z);
            const float prev_pdf = b4.w; const int fl = __float_as_int(L4.w); int nv = fl & 0xff; const bool prev_delta = (fl >> 8) & 1;
            const int slot = __float_as_int(hr.x);
            if (slot < 0) {
                if (prev_pdf > 0.f) { Bilerp bb; float mis; const float3 le = env_miss_ool(P.hier, P.env, rd, prev_pdf, prev_delta, bb, mis); L = L + beta * le * mis; }
                B.L[pid] = make_float4(L.x, L.y, L.z, L4.w);
            } else if (nv < max_verts) {
                Hit h; h.slot = slot; h.tri = 0; h.t = hr.y; h.u = hr.z; h.v = hr.w;
                const SurfacePoint sp = hit_point(M, h);
                const float3 view = f3(-rd.x, -rd.y, -rd.z);
                long long flat; const Material mt = fetch_material(P, sp.p, sp.ng, flat);
                TransMat tm; if (TRANS) tm = trans_fetch(P.cam, P.trans, flat, view, sp.ng, sp.p);
                const uint4 r4 = B.rng[pid];
                Pcg32 rng; rng.state = (uint64_t)r4.x | ((uint64_t)r4.y << 32); rng.inc = (uint64_t)r4.z | ((uint64_t)r4.w << 32);
                const float uex = rng.next_float(), uey = rng.next_float();
                const EmSample em = env_sample_direction_ool(P.hier, P.env, uex, uey);
                const float s1 = rng.next_float();
                const float s2x = rng.next_float(), s2y = rng.next_float();
                if (em.pdf != 0.f) {
                    const BsdfVal fv = TRANS ? trans_eval_brdf(em.d, view, mt, tm, P.trans) : eval_brdf_ool(em.d, view, mt);
                    if (fmax3(fv.f.x, fv.f.y, fv.f.z) > 0.f) {
                        const float3 cem = beta * fv.f * env_value(P.env, em.b) * (mis_weight(em.pdf, fv.pdf) / em.pdf);
                        B.cem[pid] = make_float4(cem.x, cem.y, cem.z, 0.f);
                        // Scene::sample_emitter_direction's visibility ray (trav_begin_shadow)
                        const ShadowRay sr = shadow_ray(M, sp.p, sp.ng, em.d);
                        B.sray_o[pid] = make_float4(sr.o.x, sr.o.y, sr.o.z, 0.f); B.sray_d[pid] = make_float4(sr.d.x, sr.d.y, sr.d.z, sr.maxt);
                        shadow = true; sbin = wf_octant(sr.d.x, sr.d.y, sr.d.z);
                    }
                }
                const BsdfSample bs = TRANS ? trans_sample_brdf(s1, s2x, s2y, view, mt, tm, P.trans, make_frame(mt.n)) : sample_brdf_ool(s1, s2x, s2y, view, mt);
                const float3 d_bs = (P.flags & MB200_FLAG_WO_WORLD_QUIRK) ? to_world(sp.sh, bs.wi) : bs.wi;
                float3 w = bs.weight;
                if (AD_W) { const BsdfVal b2 = eval_brdf_ool(d_bs, view, mt); if (b2.pdf > 0.f) w = b2.f * (1.f / b2.pdf); }
                const float3 no = offset_p(sp.p, sp.ng, d_bs);
                beta = beta * w; nv += 1;
                rng.next_float();
                B.rng[pid] = make_uint4((uint32_t)rng.state, (uint32_t)(rng.state >> 32), (uint32_t)rng.inc, (uint32_t)(rng.inc >> 32));
                B.beta[pid] = make_float4(beta.x, beta.y, beta.z, bs.pdf);
                B.L[pid] = make_float4(L.x, L.y, L.z, __int_as_float(nv));            // prev_delta = false from now on
                cont = fmax3(beta.x, beta.y, beta.z) != 0.f;
                if (cont) { B.ray_o[pid] = make_float4(no.x, no.y, no.z, 0.f); B.ray_d[pid] = make_float4(d_bs.x, d_bs.y, d_bs.z, 0.f); obin = wf_octant(d_bs.x, d_bs.y, d_bs.z); }
            }
        }
        wf_append_bin(B.counters + cout, qout, B.nb, cont, obin, pid);
        wf_append_bin(B.counters + CNT_S, B.qs, B.nb, shadow, sbin, pid);
    }
}
// film: the pixels of the batch, samples in order (same arithmetic as mesh_fwd_kernel's reduction)
template <int FILTER>
__global__ void wf_film_kernel(const __grid_constant__ RenderParams P, WfBuf B, long long pix0, int npix_batch, int s_lo, int s_n, int first, int last) {
    __shared__ __align__(16) float s_rec[FILTER == MB200_FILTER_GAUSSIAN ? kWarpsPerBlock * 32 * kRecStride : 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* rec = s_rec + (FILTER == MB200_FILTER_GAUSSIAN ? warp * 32 * kRecStride : 0);
    const int ti = lane % 5, tj = lane / 5;
    for (int pi = blockIdx.x * kWarpsPerBlock + warp; pi < npix_batch; pi += gridDim.x * kWarpsPerBlock) {
        const long long pix = pix0 + pi;
        const int py = P.prow0 + (int)(pix / P.W), px = (int)(pix % P.W);
        const int gpix = py * P.W + px;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!first) {                                                                // a pixel spanning several batches (spp > kWfBatch)
            if (FILTER == MB200_FILTER_GAUSSIAN) { if (lane < MB200_FILM_TAPS) acc = reinterpret_cast<float4*>(P.partials)[(size_t)pix * MB200_FILM_TAPS + lane]; }
            else if (lane == 0) acc = reinterpret_cast<float4*>(P.partials)[pix];
        }
        for (int b0 = 0; b0 < s_n; b0 += 32) {
            const int sl = b0 + lane; const bool act = sl < s_n;
            float3 Ls = f3(0.f, 0.f, 0.f);
            if (act) { const float4 r4 = B.L[(size_t)pi * s_n + sl]; Ls = f3(r4.x, r4.y, r4.z); }
            if (FILTER == MB200_FILTER_GAUSSIAN) {
                float wx[5], wy[5];
                if (act) {
                    Pcg32 r2; r2.seed(P.seed, (uint32_t)gpix * (uint32_t)P.spp + (uint32_t)(s_lo + sl));
                    const float jx = r2.next_float(), jy = r2.next_float();
                    film_taps(jx, wx); film_taps(jy, wy);
                } else {
#pragma unroll
                    for (int i = 0; i < 5; ++i) { wx[i] = 0.f; wy[i] = 0.f; }
                }
                float4* r4 = reinterpret_cast<float4*>(rec + lane * kRecStride);
                r4[0] = make_float4(wx[0], wx[1], wx[2], wx[3]);
                r4[1] = make_float4(wx[4], wy[0], wy[1], wy[2]);
                r4[2] = make_float4(wy[3], wy[4], 0.f, 0.f);
                r4[3] = make_float4(Ls.x, Ls.y, Ls.z, 1.f);
                __syncwarp();
                if (lane < MB200_FILM_TAPS) {
                    const float* rt = rec + ti; const float* ru = rec + 5 + tj;
#pragma unroll 8
                    for (int k = 0; k < 32; ++k) {
                        const float w = rt[k * kRecStride] * ru[k * kRecStride];
                        const float4 l4 = *reinterpret_cast<const float4*>(rec + k * kRecStride + 12);
                        acc.x = fmaf(w, l4.x, acc.x); acc.y = fmaf(w, l4.y, acc.y); acc.z = fmaf(w, l4.z, acc.z); acc.w += w;
                    }
                }
                __syncwarp();
            } else {
                acc.x += Ls.x; acc.y += Ls.y; acc.z += Ls.z;
            }
        }
        if (FILTER == MB200_FILTER_GAUSSIAN) {
            if (lane < MB200_FILM_TAPS) reinterpret_cast<float4*>(P.partials)[(size_t)pix * MB200_FILM_TAPS + lane] = acc;
        } else {
            float4 t = acc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                t.x += __shfl_xor_sync(0xffffffffu, t.x, o); t.y += __shfl_xor_sync(0xffffffffu, t.y, o); t.z += __shfl_xor_sync(0xffffffffu, t.z, o);
            }
            (void)last;
            if (lane == 0) reinterpret_cast<float4*>(P.partials)[pix] = make_float4(t.x, t.y, t.z, (float)P.spp);
        }
    }
}

// ---------------------------------------------------------------- adjoint
struct VRec {                       // what the backward walk needs of one scattering vertex (30 words, local memory)
    Material mt; float3 view, em_d, cem, d_bs, cpre, w, E; int flat;
};

// Persistent lanes as in mesh_fwd_kernel: the forward walk of each path is cut at its rays; a lane whose path has ended
// keeps its vertex records until the warp-synchronous backward walk that follows the shading pass (lanes that did not end
// a path in this pass take part with zero vertices), then fetches the next path of the pool.
template <int FILTER, bool WANT_MAT, bool WANT_N, bool WANT_ENV>
__global__ void __launch_bounds__(kThreads, MB200_MESH_MIN_BLOCKS_BWD) mesh_bwd_kernel(const __grid_constant__ RenderParams P, const __grid_constant__ MeshView M) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npix = P.prows * P.W;
    float4* const genv = WANT_ENV ? P.g_env4 + (long long)(blockIdx.x % P.env_slabs) * P.env_slab_stride : nullptr;
    const int max_verts = min(P.max_depth - 1, kMaxVerts);
    const float3 cam_o = f3(P.cam.c2w[3], P.cam.c2w[7], P.cam.c2w[11]);
    const unsigned lt_mask = (1u << lane) - 1u;
    VRec recs[WANT_MAT ? kMaxVerts : 1];
    __shared__ uint2 s_stack[MB200_MESH_SM_STACK_BWD > 0 ? MB200_MESH_SM_STACK_BWD * kThreads : 1];
    uint2 stack_loc[kStack];
    TStack<MB200_MESH_SM_STACK_BWD> stack;
    stack.loc = stack_loc; stack.sh = s_stack + threadIdx.x;
    const int ppp = P.spp >= kPoolBwd ? 1 : kPoolBwd / P.spp;          // whole pixels per pool (no film reduction here: a pool may hold any number of samples)
    const int npools = (npix + ppp - 1) / ppp;
    for (int pool_id = blockIdx.x * kWarpsPerBlock + warp; pool_id < npools; pool_id += gridDim.x * kWarpsPerBlock) {
        const int pix0 = pool_id * ppp, pixn = min(ppp, npix - pix0);
        const int pool_n = pixn * P.spp;
        int next_q = 0, stage = ST_IDLE, nv = 0;
        Trav T; T.active = false;
        Pcg32 rng;
        float3 beta = f3(1.f, 1.f, 1.f), dl = f3(0.f, 0.f, 0.f), R = f3(0.f, 0.f, 0.f), nro = f3(0.f, 0.f, 0.f), nrd = f3(0.f, 0.f, 1.f);
        float3 scat = f3(0.f, 0.f, 0.f), pE = f3(0.f, 0.f, 0.f), pcem = f3(0.f, 0.f, 0.f); Bilerp pb; pb.i00 = 0; pb.w0x = pb.w1x = pb.w0y = pb.w1y = 0.f;
        float prev_pdf = 1.f; bool prev_delta = true, dead = false;
        for (;;) {
            bool path_end = false;
            // ---- shading pass
            if (!T.active && stage != ST_DONE) {
                bool start_next = false;
                if (stage == ST_CLOSEST) {
                    if (!T.found) {
                        if (prev_pdf > 0.f) {
                            Bilerp bb; float mis; const float3 le = env_miss_ool(P.hier, P.env, T.d, prev_pdf, prev_delta, bb, mis);
                            if (WANT_MAT) R = le * mis;
                            if (WANT_ENV) env_scatter(genv, P.env.Wi, bb, dl * beta * mis);
                        }
                        path_end = true;
                    } else if (nv >= max_verts) {
                        path_end = true;
                    } else {
                        const SurfacePoint sp = hit_point(M, T.h);
                        const float3 view = f3(-T.d.x, -T.d.y, -T.d.z);
                        long long flat; const Material mt = fetch_material(P, sp.p, sp.ng, flat);
                        const float uex = rng.next_float(), uey = rng.next_float();
                        const EmSample em = env_sample_direction_ool(P.hier, P.env, uex, uey);
                        const float s1 = rng.next_float();
                        const float s2x = rng.next_float(), s2y = rng.next_float();
                        pE = f3(0.f, 0.f, 0.f); pcem = f3(0.f, 0.f, 0.f); scat = f3(0.f, 0.f, 0.f);
                        bool need_shadow = false;                   // f = 0 exactly (emitter sample below the horizon): value and every gradient
                        if (em.pdf != 0.f) {                        // of this term are 0 whatever the visibility -> the shadow ray is not traced
                            const BsdfVal fv = eval_brdf_ool(em.d, view, mt);
                            need_shadow = dot(mt.n, em.d) > 0.f;       // NoL = 0 multiplies the value and every term of eval_brdf_grad
                            const float k = mis_weight(em.pdf, fv.pdf) / em.pdf;
                            if (WANT_MAT) { const float3 lek = env_value(P.env, em.b) * k; pE = fv.f * lek; pcem = dl * beta * lek; }
                            if (WANT_ENV) { scat = dl * beta * fv.f * k; pb = em.b; }
                        }
                        const BsdfSample bs = sample_brdf_ool(s1, s2x, s2y, view, mt);
                        const float3 d_bs = (P.flags & MB200_FLAG_WO_WORLD_QUIRK) ? to_world(sp.sh, bs.wi) : bs.wi;
                        const BsdfVal b2 = eval_brdf_ool(d_bs, view, mt);
                        const float3 w = b2.pdf > 0.f ? b2.f * (1.f / b2.pdf) : bs.weight;
                        if (WANT_MAT) {
                            VRec& V = recs[nv];
                            V.mt = mt; V.view = view; V.em_d = em.d; V.cem = f3(0.f, 0.f, 0.f); V.d_bs = d_bs; V.w = w; V.E = f3(0.f, 0.f, 0.f); V.flat = (int)flat;
                            V.cpre = b2.pdf > 0.f ? dl * beta * (1.f / b2.pdf) : f3(0.f, 0.f, 0.f);
                        }
                        nro = offset_p(sp.p, sp.ng, d_bs); nrd = d_bs;
                        beta = beta * w; prev_pdf = bs.pdf; prev_delta = false; nv += 1;
                        rng.next_float();
                        dead = fmax3(beta.x, beta.y, beta.z) == 0.f;
                        if (need_shadow) { trav_begin_shadow(M, T, sp.p, sp.ng, em.d); stage = ST_SHADOW; }
                        else if (dead) path_end = true;
                        else start_next = true;
                    }
                } else if (stage == ST_SHADOW) {
                    if (!T.found) {                                 // visible
                        if (WANT_MAT) { recs[nv - 1].E = pE; recs[nv - 1].cem = pcem; }
                        if (WANT_ENV) env_scatter(genv, P.env.Wi, pb, scat);
                    }
                    if (dead) path_end = true; else start_next = true;
                }
                if (start_next) { trav_begin(M, T, nro, nrd, kInf, false); stage = ST_CLOSEST; }
                if (path_end) stage = ST_IDLE;
            }
            // ---- backward walk of the paths that ended in this pass, warp-synchronous: R = radiance leaving vertex k+1 towards vertex k
            if (WANT_MAT) {
                const int wnv = path_end ? nv : 0;
                const int max_nv = __reduce_max_sync(0xffffffffu, wnv);
                for (int k = max_nv - 1; k >= 0; --k) {
                    float g[WANT_N ? 8 : 5];
#pragma unroll
                    for (int i = 0; i < (WANT_N ? 8 : 5); ++i) g[i] = 0.f;
                    int flat = -1 - lane;                       // unique: lanes without vertex k form singleton groups
                    if (k < wnv) {
                        const VRec& V = recs[k];
                        flat = V.flat;
                        if (V.cem.x != 0.f || V.cem.y != 0.f || V.cem.z != 0.f) {
                            const BsdfGrad bg = eval_brdf_grad_ool<WANT_N>(V.em_d, V.view, V.mt, V.cem);
                            g[0] += bg.ga.x; g[1] += bg.ga.y; g[2] += bg.ga.z; g[3] += bg.gr; g[4] += bg.gm;
                            if (WANT_N) { g[5] += bg.gn.x; g[6] += bg.gn.y; g[7] += bg.gn.z; }
                        }
                        const float3 cw = V.cpre * R;
                        if (cw.x != 0.f || cw.y != 0.f || cw.z != 0.f) {
                            const BsdfGrad bg = eval_brdf_grad_ool<WANT_N>(V.d_bs, V.view, V.mt, cw);
                            g[0] += bg.ga.x; g[1] += bg.ga.y; g[2] += bg.ga.z; g[3] += bg.gr; g[4] += bg.gm;
                            if (WANT_N) { g[5] += bg.gn.x; g[6] += bg.gn.y; g[7] += bg.gn.z; }
                        }
                        R = V.E + V.w * R;
                    }
                    const unsigned peers = __match_any_sync(0xffffffffu, flat);
                    reduce_peers(peers, g);
                    if (flat >= 0 && lane == __ffs(peers) - 1) {
                        if (P.g_a) { atomicAdd(P.g_a + 3 * (size_t)flat, g[0]); atomicAdd(P.g_a + 3 * (size_t)flat + 1, g[1]); atomicAdd(P.g_a + 3 * (size_t)flat + 2, g[2]); }
                        if (P.g_r) atomicAdd(P.g_r + flat, g[3]);
                        if (P.g_m) atomicAdd(P.g_m + flat, g[4]);
                        if (WANT_N && P.g_n) { atomicAdd(P.g_n + 3 * (size_t)flat, g[5]); atomicAdd(P.g_n + 3 * (size_t)flat + 1, g[6]); atomicAdd(P.g_n + 3 * (size_t)flat + 2, g[7]); }
                    }
                }
            }
            // ---- fetch
            {
                const bool want = stage == ST_IDLE;
                const unsigned wm = __ballot_sync(0xffffffffu, want);
                if (want) {
                    const int q = next_q + __popc(wm & lt_mask);
                    if (q < pool_n) {
                        const int pix = pix0 + q / P.spp, s = q % P.spp;
                        const int py = P.prow0 + pix / P.W, px = pix % P.W;
                        rng.seed(P.seed, (uint32_t)(py * P.W + px) * (uint32_t)P.spp + (uint32_t)s);
                        const float jx = rng.next_float(), jy = rng.next_float();
                        if (FILTER == MB200_FILTER_GAUSSIAN) {      // film adjoint: 5x5 gather of G = grad / W around the pixel
                            float wx[5], wy[5]; film_taps(jx, wx); film_taps(jy, wy);
                            dl = f3(0.f, 0.f, 0.f);
#pragma unroll
                            for (int j = 0; j < 5; ++j) {
                                float3 row = f3(0.f, 0.f, 0.f);
                                const int qy = py + (j - 2);
                                const bool rowok = qy >= P.grow0 && qy < P.grow0 + P.grows && qy >= 0 && qy < P.H;
#pragma unroll
                                for (int i = 0; i < 5; ++i) {
                                    const int qx = px + (i - 2);
                                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                                    if (rowok && qx >= 0 && qx < P.W) g = __ldg(P.gadj + (size_t)(qy - P.grow0) * P.W + qx);
                                    row.x = fmaf(wx[i], g.x, row.x); row.y = fmaf(wx[i], g.y, row.y); row.z = fmaf(wx[i], g.z, row.z);
                                }
                                dl.x = fmaf(wy[j], row.x, dl.x); dl.y = fmaf(wy[j], row.y, dl.y); dl.z = fmaf(wy[j], row.z, dl.z);
                            }
                        } else {
                            const float4 g = __ldg(P.gadj + (size_t)(py - P.grow0) * P.W + px);
                            dl = f3(g.x, g.y, g.z);
                        }
                        beta = f3(1.f, 1.f, 1.f); R = f3(0.f, 0.f, 0.f); prev_pdf = 1.f; prev_delta = true; nv = 0;
                        trav_begin(M, T, cam_o, primary_dir(P.cam, XADD((float)px, jx), XADD((float)py, jy)), kInf, false);
                        stage = ST_CLOSEST;
                    } else stage = ST_DONE;
                }
                next_q += __popc(wm);
            }
            // ---- traversal
            if (!__ballot_sync(0xffffffffu, T.active)) break;
            for (;;) {
                trav_step(M, T, stack);
                const int na = __popc(__ballot_sync(0xffffffffu, T.active));
                if (na == 0) break;
                if (na < kMinActiveBwd && na < __popc(__ballot_sync(0xffffffffu, stage != ST_DONE))) break;
            }
        }
    }
}

// ---------------------------------------------------------------- debug / G-buffer extraction kernels
// ---------------------------------------------------------------- adjoint, wavefront formulation
// As the forward: wf_gen_bwd (primary rays + the film cotangent dl of every path) -> per bounce wf_trace<closest> ->
// wf_shade_bwd (miss term; vertex: record for the backward walk, parked emitter term, shadow + continuation rays) ->
// wf_trace<visibility> -> wf_apply_bwd (unoccluded: scatter the emitter term's envmap gradient; occluded: drop the term from the
// record) -> finally wf_walk: one thread per path, the backward walk over its records with the same per-texel peer reduction.
template <int FILTER>
__global__ void wf_gen_bwd_kernel(const __grid_constant__ RenderParams P, WfBuf B, long long pix0, int nb) {
    const float3 cam_o = f3(P.cam.c2w[3], P.cam.c2w[7], P.cam.c2w[11]);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nb; p += gridDim.x * blockDim.x) {
        const long long pix = pix0 + p / P.spp; const int s = p % P.spp;
        const int py = P.prow0 + (int)(pix / P.W), px = (int)(pix % P.W);
        Pcg32 rng; rng.seed(P.seed, (uint32_t)(py * P.W + px) * (uint32_t)P.spp + (uint32_t)s);
        const float jx = rng.next_float(), jy = rng.next_float();
        float3 dl;
        if (FILTER == MB200_FILTER_GAUSSIAN) {          // film adjoint: 5x5 gather of G = grad / W around the pixel
            float wx[5], wy[5]; film_taps(jx, wx); film_taps(jy, wy);
            dl = f3(0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                float3 row = f3(0.f, 0.f, 0.f);
                const int qy = py + (j - 2);
                const bool rowok = qy >= P.grow0 && qy < P.grow0 + P.grows && qy >= 0 && qy < P.H;
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const int qx = px + (i - 2);
                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rowok && qx >= 0 && qx < P.W) g = __ldg(P.gadj + (size_t)(qy - P.grow0) * P.W + qx);
                    row.x = fmaf(wx[i], g.x, row.x); row.y = fmaf(wx[i], g.y, row.y); row.z = fmaf(wx[i], g.z, row.z);
                }
                dl.x = fmaf(wy[j], row.x, dl.x); dl.y = fmaf(wy[j], row.y, dl.y); dl.z = fmaf(wy[j], row.z, dl.z);
            }
        } else {
            const float4 g = __ldg(P.gadj + (size_t)(py - P.grow0) * P.W + px);
            dl = f3(g.x, g.y, g.z);
        }
        const float3 d = primary_dir(P.cam, XADD((float)px, jx), XADD((float)py, jy));
        B.ray_o[p] = make_float4(cam_o.x, cam_o.y, cam_o.z, XADD((float)px, jx)); B.ray_d[p] = make_float4(d.x, d.y, d.z, XADD((float)py, jy));
        B.beta[p] = make_float4(1.f, 1.f, 1.f, 1.f);
        B.L[p] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x100));               // (R.xyz, nv | prev_delta << 8)
        B.dl[p] = make_float4(dl.x, dl.y, dl.z, 0.f);
        B.rng[p] = make_uint4((uint32_t)rng.state, (uint32_t)(rng.state >> 32), (uint32_t)rng.inc, (uint32_t)(rng.inc >> 32));
        B.qa[p] = (uint32_t)p;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < 64; ++i) B.counters[i] = 0;
        B.counters[CNT_A] = (uint32_t)nb;                                          // primary rays: one bin (they are coherent anyway)
    }
}
template <bool WANT_MAT, bool WANT_ENV>
__global__ void __launch_bounds__(kThreads, 2) wf_shade_bwd_kernel(const __grid_constant__ RenderParams P, const __grid_constant__ MeshView M, WfBuf B,
                                                                    const uint32_t* __restrict__ qin, uint32_t* __restrict__ qout, int cin, int cout) {
    const QView qv = wf_qview(B.counters + cin);
    const uint32_t n = qv.pre[kBins];
    const int max_verts = min(P.max_depth - 1, kMaxVerts);
    float4* const genv = WANT_ENV ? P.g_env4 + (long long)(blockIdx.x % P.env_slabs) * P.env_slab_stride : nullptr;
    const uint32_t nthreads = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += nthreads) {
        const uint32_t i = i0 + threadIdx.x;
        const bool have = i < n;
        bool cont = false, shadow = false; uint32_t pid = 0; int obin = 0, sbin = 0;
        if (have) {
            pid = wf_qget(qin, qv, i, B.nb);
            const float4 hr = B.hit[pid], rd4 = B.ray_d[pid], b4 = B.beta[pid], L4 = B.L[pid], dl4 = B.dl[pid];
            const float3 rd = f3(rd4.x, rd4.y, rd4.z), dl = f3(dl4.x, dl4.y, dl4.z);
            float3 beta = f3(b4.x, b4.y, b4.z);
            const float prev_pdf = b4.w; const int fl = __float_as_int(L4.w); int nv = fl & 0xff; const bool prev_delta = (fl >> 8) & 1;
            const int slot = __float_as_int(hr.x);
            if (slot < 0) {
                if (prev_pdf > 0.f) {
                    Bilerp bb; float mis; const float3 le = env_miss_ool(P.hier, P.env, rd, prev_pdf, prev_delta, bb, mis);
                    if (WANT_MAT) B.L[pid] = make_float4(le.x * mis, le.y * mis, le.z * mis, L4.w);
                    if (WANT_ENV) env_scatter(genv, P.env.Wi, bb, dl * beta * mis);
                }
            } else if (nv < max_verts) {
                Hit h; h.slot = slot; h.tri = 0; h.t = hr.y; h.u = hr.z; h.v = hr.w;
                const SurfacePoint sp = hit_point(M, h);
                const float3 view = f3(-rd.x, -rd.y, -rd.z);
                long long flat; const Material mt = fetch_material(P, sp.p, sp.ng, flat);
                const uint4 r4 = B.rng[pid];
                Pcg32 rng; rng.state = (uint64_t)r4.x | ((uint64_t)r4.y << 32); rng.inc = (uint64_t)r4.z | ((uint64_t)r4.w << 32);
                const float uex = rng.next_float(), uey = rng.next_float();
                const EmSample em = env_sample_direction_ool(P.hier, P.env, uex, uey);
                const float s1 = rng.next_float();
                const float s2x = rng.next_float(), s2y = rng.next_float();
                float3 E = f3(0.f, 0.f, 0.f), cem = f3(0.f, 0.f, 0.f);
                if (em.pdf != 0.f && dot(mt.n, em.d) > 0.f) {      // NoL = 0 multiplies the value and every gradient term: no shadow ray
                    const BsdfVal fv = eval_brdf_ool(em.d, view, mt);
                    const float k = mis_weight(em.pdf, fv.pdf) / em.pdf;
                    if (WANT_MAT) { const float3 lek = env_value(P.env, em.b) * k; E = fv.f * lek; cem = dl * beta * lek; }
                    if (WANT_ENV) {
                        const float3 sc = dl * beta * fv.f * k;
                        B.scat[pid] = make_float4(sc.x, sc.y, sc.z, __uint_as_float(em.b.i00));
                        B.pbw[pid] = make_float4(em.b.w0x, em.b.w1x, em.b.w0y, em.b.w1y);
                    }
                    const ShadowRay sr = shadow_ray(M, sp.p, sp.ng, em.d);
                    B.sray_o[pid] = make_float4(sr.o.x, sr.o.y, sr.o.z, 0.f); B.sray_d[pid] = make_float4(sr.d.x, sr.d.y, sr.d.z, sr.maxt);
                    shadow = true; sbin = wf_octant(sr.d.x, sr.d.y, sr.d.z);
                }
                const BsdfSample bs = sample_brdf_ool(s1, s2x, s2y, view, mt);
                const float3 d_bs = (P.flags & MB200_FLAG_WO_WORLD_QUIRK) ? to_world(sp.sh, bs.wi) : bs.wi;
                const BsdfVal b2 = eval_brdf_ool(d_bs, view, mt);
                const float3 w = b2.pdf > 0.f ? b2.f * (1.f / b2.pdf) : bs.weight;
                if (WANT_MAT) {
                    const float3 cpre = b2.pdf > 0.f ? dl * beta * (1.f / b2.pdf) : f3(0.f, 0.f, 0.f);
                    *vrec_at(B, nv, 0, pid) = make_float4(mt.a.x, mt.a.y, mt.a.z, mt.r);
                    *vrec_at(B, nv, 1, pid) = make_float4(mt.m, mt.n.x, mt.n.y, mt.n.z);
                    *vrec_at(B, nv, 2, pid) = make_float4(view.x, view.y, view.z, __int_as_float((int)flat));
                    *vrec_at(B, nv, 3, pid) = make_float4(em.d.x, em.d.y, em.d.z, cem.x);
                    *vrec_at(B, nv, 4, pid) = make_float4(cem.y, cem.z, d_bs.x, d_bs.y);
                    *vrec_at(B, nv, 5, pid) = make_float4(d_bs.z, cpre.x, cpre.y, cpre.z);
                    *vrec_at(B, nv, 6, pid) = make_float4(w.x, w.y, w.z, E.x);
                    *vrec_at(B, nv, 7, pid) = make_float4(E.y, E.z, 0.f, 0.f);
                }
                const float3 no = offset_p(sp.p, sp.ng, d_bs);
                beta = beta * w; nv += 1;
                rng.next_float();
                B.rng[pid] = make_uint4((uint32_t)rng.state, (uint32_t)(rng.state >> 32), (uint32_t)rng.inc, (uint32_t)(rng.inc >> 32));
                B.beta[pid] = make_float4(beta.x, beta.y, beta.z, bs.pdf);
                B.L[pid] = make_float4(0.f, 0.f, 0.f, __int_as_float(nv));
                cont = fmax3(beta.x, beta.y, beta.z) != 0.f;
                if (cont) { B.ray_o[pid] = make_float4(no.x, no.y, no.z, 0.f); B.ray_d[pid] = make_float4(d_bs.x, d_bs.y, d_bs.z, 0.f); obin = wf_octant(d_bs.x, d_bs.y, d_bs.z); }
            }
        }
        wf_append_bin(B.counters + cout, qout, B.nb, cont, obin, pid);
        wf_append_bin(B.counters + CNT_S, B.qs, B.nb, shadow, sbin, pid);
    }
}
template <bool WANT_MAT, bool WANT_ENV>
__global__ void wf_apply_bwd_kernel(const __grid_constant__ RenderParams P, WfBuf B) {
    const QView qv = wf_qview(B.counters + CNT_S);
    const uint32_t n = qv.pre[kBins];
    float4* const genv = WANT_ENV ? P.g_env4 + (long long)(blockIdx.x % P.env_slabs) * P.env_slab_stride : nullptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t pid = wf_qget(B.qs, qv, i, B.nb);
        if (B.vis[pid]) {
            if (WANT_ENV) {
                const float4 sc = B.scat[pid], w = B.pbw[pid];
                Bilerp b; b.i00 = __float_as_uint(sc.w); b.w0x = w.x; b.w1x = w.y; b.w0y = w.z; b.w1y = w.w;
                env_scatter(genv, P.env.Wi, b, f3(sc.x, sc.y, sc.z));
            }
        } else if (WANT_MAT) {                                                     // occluded: the emitter term of this vertex vanishes
            const int k = (__float_as_int(B.L[pid].w) & 0xff) - 1;
            float4* r3 = vrec_at(B, k, 3, pid); float4* r4 = vrec_at(B, k, 4, pid); float4* r6 = vrec_at(B, k, 6, pid); float4* r7 = vrec_at(B, k, 7, pid);
            float4 v3 = *r3, v4 = *r4, v6 = *r6, v7 = *r7;
            v3.w = 0.f; v4.x = 0.f; v4.y = 0.f; v6.w = 0.f; v7.x = 0.f; v7.y = 0.f;
            *r3 = v3; *r4 = v4; *r6 = v6; *r7 = v7;
        }
    }
}
template <bool WANT_N>
__global__ void __launch_bounds__(kThreads, 3) wf_walk_kernel(const __grid_constant__ RenderParams P, WfBuf B, int nb) {
    const int lane = threadIdx.x & 31;
    const int nb_pad = (nb + 31) & ~31;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nb_pad; p += gridDim.x * blockDim.x) {     // warp-uniform trip count
        const bool have = p < nb;
        float3 R = f3(0.f, 0.f, 0.f); int nv = 0;
        if (have) { const float4 L4 = B.L[p]; R = f3(L4.x, L4.y, L4.z); nv = __float_as_int(L4.w) & 0xff; }
        const int max_nv = __reduce_max_sync(0xffffffffu, nv);
        for (int k = max_nv - 1; k >= 0; --k) {
            float g[WANT_N ? 8 : 5];
#pragma unroll
            for (int i = 0; i < (WANT_N ? 8 : 5); ++i) g[i] = 0.f;
            int flat = -1 - lane;
            if (k < nv) {
                const float4 v0 = *vrec_at(B, k, 0, p), v1 = *vrec_at(B, k, 1, p), v2 = *vrec_at(B, k, 2, p), v3 = *vrec_at(B, k, 3, p);
                const float4 v4 = *vrec_at(B, k, 4, p), v5 = *vrec_at(B, k, 5, p), v6 = *vrec_at(B, k, 6, p), v7 = *vrec_at(B, k, 7, p);
                Material mt; mt.a = f3(v0.x, v0.y, v0.z); mt.r = v0.w; mt.m = v1.x; mt.n = f3(v1.y, v1.z, v1.w);
                const float3 view = f3(v2.x, v2.y, v2.z), em_d = f3(v3.x, v3.y, v3.z), cem = f3(v3.w, v4.x, v4.y), d_bs = f3(v4.z, v4.w, v5.x);
                const float3 cpre = f3(v5.y, v5.z, v5.w), w = f3(v6.x, v6.y, v6.z), E = f3(v6.w, v7.x, v7.y);
                flat = __float_as_int(v2.w);
                if (cem.x != 0.f || cem.y != 0.f || cem.z != 0.f) {
                    const BsdfGrad bg = eval_brdf_grad_ool<WANT_N>(em_d, view, mt, cem);
                    g[0] += bg.ga.x; g[1] += bg.ga.y; g[2] += bg.ga.z; g[3] += bg.gr; g[4] += bg.gm;
                    if (WANT_N) { g[5] += bg.gn.x; g[6] += bg.gn.y; g[7] += bg.gn.z; }
                }
                const float3 cw = cpre * R;
                if (cw.x != 0.f || cw.y != 0.f || cw.z != 0.f) {
                    const BsdfGrad bg = eval_brdf_grad_ool<WANT_N>(d_bs, view, mt, cw);
                    g[0] += bg.ga.x; g[1] += bg.ga.y; g[2] += bg.ga.z; g[3] += bg.gr; g[4] += bg.gm;
                    if (WANT_N) { g[5] += bg.gn.x; g[6] += bg.gn.y; g[7] += bg.gn.z; }
                }
                R = E + w * R;
            }
            const unsigned peers = __match_any_sync(0xffffffffu, flat);
            reduce_peers(peers, g);
            if (flat >= 0 && lane == __ffs(peers) - 1) {
                if (P.g_a) { atomicAdd(P.g_a + 3 * (size_t)flat, g[0]); atomicAdd(P.g_a + 3 * (size_t)flat + 1, g[1]); atomicAdd(P.g_a + 3 * (size_t)flat + 2, g[2]); }
                if (P.g_r) atomicAdd(P.g_r + flat, g[3]);
                if (P.g_m) atomicAdd(P.g_m + flat, g[4]);
                if (WANT_N && P.g_n) { atomicAdd(P.g_n + 3 * (size_t)flat, g[5]); atomicAdd(P.g_n + 3 * (size_t)flat + 1, g[6]); atomicAdd(P.g_n + 3 * (size_t)flat + 2, g[7]); }
            }
        }
    }
}

__global__ void mesh_intersect_kernel(const __grid_constant__ MeshView M, const float* __restrict__ o, const float* __restrict__ d,
                                      const float* __restrict__ maxt, int n, int any_hit, int32_t* __restrict__ out_tri, float* __restrict__ out_tuv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 ro = f3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), rd = f3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
    const float mt = maxt ? maxt[i] : kInf;
    Hit h; bool found;
    if (any_hit) found = mesh_intersect<true>(M, ro, rd, mt, h); else found = mesh_intersect<false>(M, ro, rd, mt, h);
    out_tri[i] = any_hit ? (found ? 1 : 0) : (found ? h.tri : -1);
    if (out_tuv) { out_tuv[3 * i] = found && !any_hit ? h.t : 0.f; out_tuv[3 * i + 1] = found && !any_hit ? h.u : 0.f; out_tuv[3 * i + 2] = found && !any_hit ? h.v : 0.f; }
}
// primary visibility at film offset (jx, jy) inside every pixel: gpos (H,W,4) = (p, hit?1:0), gnrm (H,W,4), tri (H,W), flat (H,W)
__global__ void mesh_primary_kernel(const __grid_constant__ RenderParams P, const __grid_constant__ MeshView M, float jx, float jy,
                                    float4* __restrict__ gpos, float4* __restrict__ gnrm, int32_t* __restrict__ tri, int32_t* __restrict__ flat_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.H * P.W) return;
    const int py = i / P.W, px = i % P.W;
    const float3 ro = f3(P.cam.c2w[3], P.cam.c2w[7], P.cam.c2w[11]);
    const float3 rd = primary_dir(P.cam, XADD((float)px, jx), XADD((float)py, jy));
    Hit h;
    if (mesh_intersect<false>(M, ro, rd, kInf, h)) {
        const SurfacePoint sp = hit_point(M, h);
        gpos[i] = make_float4(sp.p.x, sp.p.y, sp.p.z, 1.f); gnrm[i] = make_float4(sp.ng.x, sp.ng.y, sp.ng.z, 0.f);
        if (tri) tri[i] = h.tri;
        if (flat_out) flat_out[i] = (int32_t)texel_index(P.cam, sp.p);
    } else {
        gpos[i] = make_float4(0.f, 0.f, 0.f, 0.f); gnrm[i] = make_float4(0.f, 0.f, 1.f, 0.f);
        if (tri) tri[i] = -1;
        if (flat_out) flat_out[i] = -1;
    }
}

// ================================================================ BVH build (GPU)
__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void mesh_init_kernel(int* bbox_ord) {
    if (threadIdx.x < 3) bbox_ord[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) bbox_ord[threadIdx.x] = (int)0x80000000;
}
__device__ __forceinline__ void tri_bounds(const float* __restrict__ verts, const int32_t* __restrict__ tris, int t, float lo[3], float hi[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = 1e30f; hi[k] = -1e30f; }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int vi = tris[3 * (size_t)t + j];
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float x = verts[3 * (size_t)vi + k]; lo[k] = fminf(lo[k], x); hi[k] = fmaxf(hi[k], x); }
    }
}
__global__ void mesh_scene_bounds_kernel(const float* __restrict__ verts, const int32_t* __restrict__ tris, int nt, int* bbox_ord) {
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
        float l[3], h[3]; tri_bounds(verts, tris, t, l, h);
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], l[k]); hi[k] = fmaxf(hi[k], h[k]); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(bbox_ord + k, f2ord(lo[k])); atomicMax(bbox_ord + 3 + k, f2ord(hi[k])); }
    }
}
// header: centre.xyz, radius = |bbox.max - centre|, then bbox lo.xyz, hi.xyz  (operation order of the oracle's mbo_mesh_create)
__global__ void mesh_header_kernel(const int* bbox_ord, float* header) {
    float r2 = 0.f;
    for (int k = 0; k < 3; ++k) {
        const float lo = ord2f(bbox_ord[k]), hi = ord2f(bbox_ord[3 + k]);
        const float c = XMUL(0.5f, XADD(lo, hi)), e = XSUB(hi, c);
        header[k] = c; r2 = XADD(r2, XMUL(e, e));
        header[4 + k] = lo; header[8 + k] = hi;
    }
    header[3] = XSQRT(r2); header[7] = 0.f; header[11] = 0.f;
}
__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t expand15(uint32_t v) {        // 15 bits -> every second bit of 30
    v &= 0x7fffu;
    v = (v | (v << 8)) & 0x00FF00FFu; v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u; v = (v | (v << 1)) & 0x55555555u;
    return v;
}
// ORDER 0: 30-bit 3-D Morton code of the centroid in the scene box (any mesh).
// ORDER 1: 30-bit 2-D Morton code of the DIRECTION of the centroid as seen from the origin (octahedral map, 15 bits per axis).
//   The reference's meshes are depth maps unprojected from a camera at the origin (mesh_recon.py:184-258: vertex k <-> pixel k), i.e.
//   sheets that a 3-D curve threads badly — depth bits interleave with the two screen axes, neighbouring quads end up far apart and
//   the implicit tree's mid-level boxes overlap (measured, profiles/r5o_trav_stats.log: 63 box steps per closest-hit ray).  Ordered
//   by viewing direction the implicit 4-ary tree IS the screen-space quadtree of the depth map.
template <int ORDER>
__global__ void mesh_morton_kernel(const float* __restrict__ verts, const int32_t* __restrict__ tris, int nt, const float* __restrict__ header,
                                   uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    float lo[3], hi[3]; tri_bounds(verts, tris, t, lo, hi);
    if (ORDER == 1) {
        const float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
        const float l1 = fabsf(cx) + fabsf(cy) + fabsf(cz);
        float u = 0.f, v = 0.f;
        if (l1 > 0.f) {
            u = cx / l1; v = cy / l1;
            if (cz > 0.f) {          // the hemisphere behind a camera that looks down -z: folded over the diagonals
                const float fu = (1.f - fabsf(v)) * (u >= 0.f ? 1.f : -1.f), fv = (1.f - fabsf(u)) * (v >= 0.f ? 1.f : -1.f);
                u = fu; v = fv;
            }
        }
        const uint32_t qu = (uint32_t)fminf(fmaxf((u * 0.5f + 0.5f) * 32768.f, 0.f), 32767.f);
        const uint32_t qv = (uint32_t)fminf(fmaxf((v * 0.5f + 0.5f) * 32768.f, 0.f), 32767.f);
        keys[t] = (expand15(qv) << 1) | expand15(qu);
        vals[t] = t;
        return;
    }
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float blo = header[4 + k], ext = header[8 + k] - blo;
        const float x = ext > 0.f ? (0.5f * (lo[k] + hi[k]) - blo) / ext : 0.f;
        q[k] = (uint32_t)fminf(fmaxf(x * 1024.f, 0.f), 1023.f);
    }
    keys[t] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    vals[t] = t;
}
// Mesh::recompute_vertex_normals: face normals weighted by the corner angle, accumulated in double
// (exact chain: the oracle's unit_angle with the shared mbx_asin01, so that the interpolated shading normals are bit-identical)
__device__ __forceinline__ float unit_angle(float3 u, float3 v) {
    const float3 dm = xsub3(v, u), dp = xadd3(v, u);
    const float t = XMUL(2.f, mbx_asin01(fminf(XMUL(0.5f, XSQRT(xdot3(dm, dm))), 1.f)));
    return xdot3(u, v) >= 0.f ? t : XSUB(MB_PI, XMUL(2.f, mbx_asin01(fminf(XMUL(0.5f, XSQRT(xdot3(dp, dp))), 1.f))));
}
__global__ void mesh_vn_accum_kernel(const float* __restrict__ verts, const int32_t* __restrict__ tris, int nt, double* __restrict__ acc) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    int vi[3]; float3 p[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { vi[j] = tris[3 * (size_t)t + j]; p[j] = f3(verts[3 * (size_t)vi[j]], verts[3 * (size_t)vi[j] + 1], verts[3 * (size_t)vi[j] + 2]); }
    float3 n = xcross3(xsub3(p[1], p[0]), xsub3(p[2], p[0]));
    const float l2 = xdot3(n, n);
    if (l2 == 0.f) return;
    const float il = XDIV(1.f, XSQRT(l2));
    n = f3(XMUL(n.x, il), XMUL(n.y, il), XMUL(n.z, il));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float3 d0 = xnormalize3(xsub3(p[(i + 1) % 3], p[i])), d1 = xnormalize3(xsub3(p[(i + 2) % 3], p[i]));
        const float w = unit_angle(d0, d1);
        atomicAdd(acc + 3 * (size_t)vi[i], (double)XMUL(n.x, w));
        atomicAdd(acc + 3 * (size_t)vi[i] + 1, (double)XMUL(n.y, w));
        atomicAdd(acc + 3 * (size_t)vi[i] + 2, (double)XMUL(n.z, w));
    }
}
__global__ void mesh_vn_finish_kernel(const double* __restrict__ acc, int nv, float4* __restrict__ vn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const double x = acc[3 * (size_t)i], y = acc[3 * (size_t)i + 1], z = acc[3 * (size_t)i + 2];
    const double l = sqrt(x * x + y * y + z * z);
    vn[i] = l > 0.0 ? make_float4((float)(x / l), (float)(y / l), (float)(z / l), 0.f) : make_float4(1.f, 0.f, 0.f, 0.f);
}
__global__ void mesh_gather_kernel(const float* __restrict__ verts, const int32_t* __restrict__ tris, int nt, int n_slots,
                                   const int32_t* __restrict__ order, const float4* __restrict__ vn, float4* __restrict__ tv, float4* __restrict__ tn) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    if (s >= nt) {
        tv[3 * (size_t)s] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)); tv[3 * (size_t)s + 1] = tv[3 * (size_t)s + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tn) tn[3 * (size_t)s] = tn[3 * (size_t)s + 1] = tn[3 * (size_t)s + 2] = make_float4(0.f, 0.f, 1.f, 0.f);
        return;
    }
    const int t = order[s];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int vi = tris[3 * (size_t)t + j];
        tv[3 * (size_t)s + j] = make_float4(verts[3 * (size_t)vi], verts[3 * (size_t)vi + 1], verts[3 * (size_t)vi + 2], j == 0 ? __int_as_float(t) : 0.f);
        if (tn) tn[3 * (size_t)s + j] = vn[vi];
    }
}
// box of node i at some level is component (i & 3) of the six float4 of group (i >> 2)
__device__ __forceinline__ void store_box(float4* nodes, int group_off, int i, const float lo[3], const float hi[3]) {
    float* g = reinterpret_cast<float*>(nodes + (size_t)(group_off + (i >> 2)) * 6);
#pragma unroll
    for (int k = 0; k < 3; ++k) { g[4 * k + (i & 3)] = lo[k]; g[12 + 4 * k + (i & 3)] = hi[k]; }
}
__global__ void mesh_leaf_bounds_kernel(const float4* __restrict__ tv, int n_leaves, int n_padded, float4* __restrict__ nodes, int group_off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_padded) return;
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    if (i < n_leaves) {
        for (int k = 0; k < kLeaf; ++k) {
            const float4 q0 = tv[3 * ((size_t)i * kLeaf + k)];
            if (__float_as_int(q0.w) < 0) continue;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float4 q = tv[3 * ((size_t)i * kLeaf + k) + j];
                lo[0] = fminf(lo[0], q.x); lo[1] = fminf(lo[1], q.y); lo[2] = fminf(lo[2], q.z);
                hi[0] = fmaxf(hi[0], q.x); hi[1] = fmaxf(hi[1], q.y); hi[2] = fmaxf(hi[2], q.z);
            }
        }
        // conservative padding: Moeller-Trumbore hits are not exactly on the triangle's plane in float
        const float pad = 3.814697265625e-06f * (1.f + fmaxf(fmaxf(fmaxf(fabsf(lo[0]), fabsf(hi[0])), fmaxf(fabsf(lo[1]), fabsf(hi[1]))), fmaxf(fabsf(lo[2]), fabsf(hi[2]))));
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] -= pad; hi[k] += pad; }
    }
    store_box(nodes, group_off, i, lo, hi);
}
__global__ void mesh_level_kernel(float4* __restrict__ nodes, int child_group_off, int n_nodes, int n_padded, int group_off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_padded) return;
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    if (i < n_nodes) {
        const float4* g = nodes + (size_t)(child_group_off + i) * 6;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float4 l = g[k], h = g[3 + k];
            lo[k] = fminf(fminf(l.x, l.y), fminf(l.z, l.w)); hi[k] = fmaxf(fmaxf(h.x, h.y), fmaxf(h.z, h.w));
        }
    }
    store_box(nodes, group_off, i, lo, hi);
}

// ---------------------------------------------------------------- host side
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int make_view(const mb200_mesh_desc* md, const void* mesh_buf, MeshView& M) {
    if (!md || !mesh_buf || md->nt <= 0 || md->n_levels < 1 || md->n_levels > MB200_MESH_MAX_LEVELS) return MB200_EINVAL;
    const char* b = reinterpret_cast<const char*>(mesh_buf);
    M.header = reinterpret_cast<const float*>(b + md->off_header);
    M.tv = reinterpret_cast<const float4*>(b + md->off_tv);
    M.tn = md->face_normals ? nullptr : reinterpret_cast<const float4*>(b + md->off_tn);
    M.nodes = reinterpret_cast<const float4*>(b + md->off_nodes);
    M.n_levels = md->n_levels;
    for (int l = 0; l < MB200_MESH_MAX_LEVELS; ++l) M.lvl_off[l] = l < md->n_levels ? md->lvl_group_off[l] : 0;
    return MB200_OK;
}

struct ScratchLayout { size_t bbox, keys0, keys1, vals0, vals1, vnacc, vn, cub, cub_bytes, total; };
int scratch_layout(int nv, int nt, int face_normals, ScratchLayout& S) {
    size_t off = 0;
    S.bbox = off; off += 256;
    S.keys0 = off; off += align_up(sizeof(uint32_t) * (size_t)nt, 256);
    S.keys1 = off; off += align_up(sizeof(uint32_t) * (size_t)nt, 256);
    S.vals0 = off; off += align_up(sizeof(int32_t) * (size_t)nt, 256);
    S.vals1 = off; off += align_up(sizeof(int32_t) * (size_t)nt, 256);
    S.vnacc = off; off += face_normals ? 0 : align_up(sizeof(double) * 3 * (size_t)nv, 256);
    S.vn = off; off += face_normals ? 0 : align_up(sizeof(float4) * (size_t)nv, 256);
    size_t cub_bytes = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr); cub::DoubleBuffer<int32_t> v(nullptr, nullptr);
    if (cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, k, v, nt, 0, 30) != cudaSuccess) { cudaGetLastError(); return MB200_ELAUNCH; }
    S.cub = off; S.cub_bytes = cub_bytes; off += align_up(cub_bytes, 256);
    S.total = off;
    return MB200_OK;
}

int mesh_render_params(const mb200_cfg* c, const float* a, const float* r, const float* m, const float* n_opt, const float* env4,
                       const float* hier, const mb200_hier_desc* d, RenderParams& P) {
    int rc = fill_params(c, nullptr, nullptr, a, r, m, n_opt, env4, hier, d, P, false);
    if (rc) return rc;
    if (c->max_depth - 1 > kMaxVerts) return MB200_ERANGE;
    return MB200_OK;
}

template <int FILTER>
int launch_mesh_bwd(const RenderParams& P, const MeshView& M, bool want_mat, bool want_n, bool want_env, cudaStream_t st) {
    const int grid = grid_for(P.prows * P.W);
#define MB_BWD(MT, N, E) mesh_bwd_kernel<FILTER, MT, N, E><<<grid, kThreads, 0, st>>>(P, M)
    if (want_mat && want_n && want_env) MB_BWD(true, true, true);
    else if (want_mat && want_n) MB_BWD(true, true, false);
    else if (want_mat && want_env) MB_BWD(true, false, true);
    else if (want_mat) MB_BWD(true, false, false);
    else if (want_env) MB_BWD(false, false, true);
#undef MB_BWD
    return mb200_check_launch();
}

}  // namespace

extern "C" {

int mb200_mesh_describe(int nv, int nt, int face_normals, mb200_mesh_desc* out) {
    if (!out || nv <= 0 || nt <= 0) return MB200_EINVAL;
    memset(out, 0, sizeof(*out));
    out->nv = nv; out->nt = nt; out->face_normals = face_normals ? 1 : 0;
    long long n = ((long long)nt + kLeaf - 1) / kLeaf;          // leaves
    if (n >= (1ll << 27)) return MB200_ERANGE;
    out->n_slots = (int32_t)(n * kLeaf);
    int L = 0; int goff = 0;
    for (;;) {
        if (L >= MB200_MESH_MAX_LEVELS) return MB200_ERANGE;
        out->lvl_nodes[L] = (int32_t)n; out->lvl_group_off[L] = goff;
        const long long groups = (n + 3) / 4;
        goff += (int)groups; ++L;
        if (n <= 4) break;
        n = groups;
    }
    out->n_levels = L; out->n_groups = goff;
    size_t off = 0;
    out->off_header = (int64_t)off; off += 256;
    out->off_tv = (int64_t)off; off += align_up(sizeof(float4) * 3 * (size_t)out->n_slots, 256);
    out->off_tn = (int64_t)off; off += face_normals ? 0 : align_up(sizeof(float4) * 3 * (size_t)out->n_slots, 256);
    out->off_nodes = (int64_t)off; off += align_up(sizeof(float4) * 6 * (size_t)goff, 256);
    out->total_bytes = (int64_t)off;
    return MB200_OK;
}

size_t mb200_mesh_scratch_bytes(int nv, int nt, int face_normals) {
    ScratchLayout S;
    if (nv <= 0 || nt <= 0 || scratch_layout(nv, nt, face_normals, S) != MB200_OK) return 0;
    return S.total;
}

int mb200_mesh_build(const float* verts, const int32_t* tris, const mb200_mesh_desc* md, void* mesh_buf, void* scratch, void* stream) {
    if (!verts || !tris || !md || !mesh_buf || !scratch) return MB200_EINVAL;
    mb200_mesh_desc chk;
    int rc = mb200_mesh_describe(md->nv, md->nt, md->face_normals, &chk);
    if (rc) return rc;
    if (chk.total_bytes != md->total_bytes || chk.n_levels != md->n_levels) return MB200_EINVAL;
    ScratchLayout S; rc = scratch_layout(md->nv, md->nt, md->face_normals, S);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    char* sb = reinterpret_cast<char*>(scratch); char* mb = reinterpret_cast<char*>(mesh_buf);
    const int nt = md->nt, nv = md->nv, tb = 256;
    int* bbox = reinterpret_cast<int*>(sb + S.bbox);
    float* header = reinterpret_cast<float*>(mb + md->off_header);
    float4* tv = reinterpret_cast<float4*>(mb + md->off_tv);
    float4* tn = md->face_normals ? nullptr : reinterpret_cast<float4*>(mb + md->off_tn);
    float4* nodes = reinterpret_cast<float4*>(mb + md->off_nodes);
    mesh_init_kernel<<<1, 32, 0, st>>>(bbox);
    int grid = (nt + tb - 1) / tb; const int cap = mb200_sm_count() * 8;
    mesh_scene_bounds_kernel<<<grid < cap ? grid : cap, tb, 0, st>>>(verts, tris, nt, bbox);
    mesh_header_kernel<<<1, 1, 0, st>>>(bbox, header);
    uint32_t* k0 = reinterpret_cast<uint32_t*>(sb + S.keys0); uint32_t* k1 = reinterpret_cast<uint32_t*>(sb + S.keys1);
    int32_t* v0 = reinterpret_cast<int32_t*>(sb + S.vals0); int32_t* v1 = reinterpret_cast<int32_t*>(sb + S.vals1);
    static int order = -1;
    if (order < 0) { const char* e = getenv("MB200_MESH_ORDER"); order = e ? atoi(e) : 0; }
    if (order == 1) mesh_morton_kernel<1><<<grid, tb, 0, st>>>(verts, tris, nt, header, k0, v0);
    else            mesh_morton_kernel<0><<<grid, tb, 0, st>>>(verts, tris, nt, header, k0, v0);
    cub::DoubleBuffer<uint32_t> kb(k0, k1); cub::DoubleBuffer<int32_t> vb(v0, v1);
    size_t cub_bytes = S.cub_bytes;
    if (cub::DeviceRadixSort::SortPairs(sb + S.cub, cub_bytes, kb, vb, nt, 0, 30, st) != cudaSuccess) return mb200_check_launch();
    float4* vn = nullptr;
    if (!md->face_normals) {
        double* acc = reinterpret_cast<double*>(sb + S.vnacc); vn = reinterpret_cast<float4*>(sb + S.vn);
        if (mb200_check(cudaMemsetAsync(acc, 0, sizeof(double) * 3 * (size_t)nv, st)) != MB200_OK) return MB200_ELAUNCH;
        mesh_vn_accum_kernel<<<grid, tb, 0, st>>>(verts, tris, nt, acc);
        mesh_vn_finish_kernel<<<(nv + tb - 1) / tb, tb, 0, st>>>(acc, nv, vn);
    }
    mesh_gather_kernel<<<(md->n_slots + tb - 1) / tb, tb, 0, st>>>(verts, tris, nt, md->n_slots, vb.Current(), vn, tv, tn);
    for (int l = 0; l < md->n_levels; ++l) {
        const int n = md->lvl_nodes[l], padded = (n + 3) / 4 * 4;
        if (l == 0) mesh_leaf_bounds_kernel<<<(padded + tb - 1) / tb, tb, 0, st>>>(tv, n, padded, nodes, md->lvl_group_off[0]);
        else mesh_level_kernel<<<(padded + tb - 1) / tb, tb, 0, st>>>(nodes, md->lvl_group_off[l - 1], n, padded, md->lvl_group_off[l]);
    }
    return mb200_check_launch();
}

int mb200_mesh_shade_fwd(const mb200_cfg* c, const mb200_mesh_desc* md, const void* mesh_buf,
                         const float* a, const float* r, const float* m, const float* n_opt,
                         const float* env4, const float* hier, const mb200_hier_desc* d, float* partials, void* stream) {
    RenderParams P; int rc = mesh_render_params(c, a, r, m, n_opt, env4, hier, d, P);
    if (rc) return rc;
    MeshView M; rc = make_view(md, mesh_buf, M);
    if (rc) return rc;
    if (!partials) return MB200_EINVAL;
    P.prows = mb200_fwd_partial_rows(c, &P.prow0); P.partials = partials;
    const int grid = grid_for(P.prows * P.W);
    cudaStream_t st = (cudaStream_t)stream;
    const bool ad = (c->flags & MB200_FLAG_AD_WEIGHTS) != 0;
    if (c->filter == MB200_FILTER_GAUSSIAN) {
        if (ad) launch_mesh_fwd(mesh_fwd_kernel<MB200_FILTER_GAUSSIAN, true, false>, c->filter, grid, st, P, M);
        else    launch_mesh_fwd(mesh_fwd_kernel<MB200_FILTER_GAUSSIAN, false, false>, c->filter, grid, st, P, M);
    } else {
        if (ad) launch_mesh_fwd(mesh_fwd_kernel<MB200_FILTER_BOX, true, false>, c->filter, grid, st, P, M);
        else    launch_mesh_fwd(mesh_fwd_kernel<MB200_FILTER_BOX, false, false>, c->filter, grid, st, P, M);
    }
    return mb200_check_launch();
}

// The shadow-ray traversal of bounce k and the closest-hit traversal of bounce k+1 are independent: the former runs on a side
// stream (one per device, created on first use) so that the nearly empty grids of the late bounces overlap.
struct WfSide { cudaStream_t s = nullptr; cudaEvent_t shaded = nullptr, any_done = nullptr; };
static WfSide* wf_side() {
    static WfSide side[64];
    int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    WfSide& w = side[dev];
    if (!w.s) {
        if (cudaStreamCreateWithFlags(&w.s, cudaStreamNonBlocking) != cudaSuccess) { w.s = nullptr; return nullptr; }
        cudaEventCreateWithFlags(&w.shaded, cudaEventDisableTiming); cudaEventCreateWithFlags(&w.any_done, cudaEventDisableTiming);
    }
    return &w;
}

size_t mb200_mesh_fwd_wf_scratch_bytes(const mb200_cfg* c) {
    if (!c || c->spp <= 0 || c->spp > kWfBatch) return 0;
    int r0; const int prows = mb200_fwd_partial_rows(c, &r0);
    const long long npix = (long long)prows * c->W, bp = kWfBatch / c->spp;
    const long long nb = (npix < bp ? npix : bp) * c->spp;
    return wf_scratch_bytes(nb);
}

static void fill_cam(const mb200_cfg* c, CamView& cam) {
    for (int i = 0; i < 16; ++i) { cam.view[i] = c->view[i]; cam.proj[i] = c->proj[i]; cam.c2w[i] = c->cam_to_world[i]; }
    cam.tan_half_fov_x = c->tan_half_fov_x; cam.H = c->H; cam.W = c->W;
    cam.stride = (c->flags & MB200_FLAG_ROW_STRIDE_H) ? c->H : c->W;
}
size_t mb200_mesh_primary_index_bytes(const mb200_cfg* c, const mb200_mesh_desc* md) {
    if (!c || !md || c->H <= 0 || c->W <= 0 || md->n_slots <= 0) return 0;
    return pidx_bytes(c->H, c->W, md->n_slots);
}
int mb200_mesh_primary_index_build(const mb200_cfg* c, const mb200_mesh_desc* md, const void* mesh_buf, void* index, void* stream) {
    if (!c || !md || !mesh_buf || !index || c->H <= 0 || c->W <= 0) return MB200_EINVAL;
    MeshView M; int rc = make_view(md, mesh_buf, M);
    if (rc) return rc;
    CamView cam; fill_cam(c, cam);
    cudaStream_t st = (cudaStream_t)stream;
    PIdxView I = pidx_view(index, c->H, c->W);
    if (mb200_check(cudaMemsetAsync(index, 0, 256 + (size_t)c->H * c->W * 4, st)) != MB200_OK) return MB200_ELAUNCH;     // header + counts
    const int one = 1;
    if (mb200_check(cudaMemcpyAsync(I.valid, &one, 4, cudaMemcpyHostToDevice, st)) != MB200_OK) return MB200_ELAUNCH;
    pidx_bin_kernel<<<(md->n_slots + 255) / 256, 256, 0, st>>>(M, cam, I, md->n_slots);
    return mb200_check_launch();
}

int mb200_mesh_shade_fwd_wf(const mb200_cfg* c, const mb200_trans* t, const mb200_mesh_desc* md, const void* mesh_buf,
                            const float* a, const float* r, const float* m, const float* n_opt,
                            const float* env4, const float* hier, const mb200_hier_desc* d, float* partials,
                            void* scratch, size_t scratch_bytes, const void* primary_index, void* stream) {
    RenderParams P; int rc = mesh_render_params(c, a, r, m, n_opt, env4, hier, d, P);
    if (rc) return rc;
    if (t && (rc = fill_trans(t, P)) != MB200_OK) return rc;
    MeshView M; rc = make_view(md, mesh_buf, M);
    if (rc) return rc;
    if (!partials || !scratch) return MB200_EINVAL;
    if (c->spp > kWfBatch) return MB200_EUNSUPPORTED;
    const bool ad = (c->flags & MB200_FLAG_AD_WEIGHTS) != 0;
    if (t && ad) return MB200_EUNSUPPORTED;
    P.prows = mb200_fwd_partial_rows(c, &P.prow0); P.partials = partials;
    const long long npix = (long long)P.prows * P.W, bp = kWfBatch / c->spp;
    const long long nb_max = (npix < bp ? npix : bp) * c->spp;
    if (scratch_bytes < wf_scratch_bytes(nb_max)) return MB200_EINVAL;
    WfBuf B = wf_carve(scratch, nb_max);
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = mb200_sm_count();
    const int max_verts = (c->max_depth - 1 < kMaxVerts ? c->max_depth - 1 : kMaxVerts);
    WfSide* side = wf_side();
    for (long long pix0 = 0; pix0 < npix; pix0 += bp) {
        const int npb = (int)((npix - pix0) < bp ? (npix - pix0) : bp);
        const int nb = npb * c->spp;
        wf_gen_kernel<<<sms * 8, 256, 0, st>>>(P, B, pix0, nb);
        uint32_t* qin = B.qa; uint32_t* qout = B.qb; int cin = CNT_A, cout = CNT_B;
        bool any_pending = false;
        for (int it = 0; it <= (max_verts < 0 ? 0 : max_verts); ++it) {
            cudaMemsetAsync(B.counters + 3, 0, 4, st);
            if (it == 0 && primary_index) {
                // primary rays: closest hit among the candidates of the ray's pixel; what the index cannot serve goes through the BVH
                const PIdxView I = pidx_view(const_cast<void*>(primary_index), c->H, c->W);
                cudaMemsetAsync(B.counters + CNT_B, 0, 4 * kBins, st);                 // fallback queue (rays the index cannot serve)
                wf_primary_kernel<<<sms * 8, 256, 0, st>>>(P, M, B, I, pix0, nb, B.qb, B.counters + CNT_B);
                wf_trace_kernel<0><<<sms * MB200_WF_TRACE_BLOCKS, kThreads, 0, st>>>(M, B, B.qb, B.counters + CNT_B, B.counters + 3, wf_leaf_thresh());
            } else
            wf_trace_kernel<0><<<sms * MB200_WF_TRACE_BLOCKS, kThreads, 0, st>>>(M, B, qin, B.counters + cin, B.counters + 3, wf_leaf_thresh());
            if (any_pending) { cudaStreamWaitEvent(st, side->any_done, 0); any_pending = false; }   // the shadow rays of the previous bounce
            cudaMemsetAsync(B.counters + cout, 0, 4 * kBins, st);
            cudaMemsetAsync(B.counters + CNT_S, 0, 4 * kBins, st);
            if (t)       wf_shade_kernel<false, true><<<sms * 4, kThreads, 0, st>>>(P, M, B, qin, qout, cin, cout);
            else if (ad) wf_shade_kernel<true, false><<<sms * 4, kThreads, 0, st>>>(P, M, B, qin, qout, cin, cout);
            else         wf_shade_kernel<false, false><<<sms * 4, kThreads, 0, st>>>(P, M, B, qin, qout, cin, cout);
            cudaStream_t sa = side ? side->s : st;
            if (side) { cudaEventRecord(side->shaded, st); cudaStreamWaitEvent(sa, side->shaded, 0); }
            cudaMemsetAsync(B.counters + 4, 0, 4, sa);
            wf_trace_kernel<1><<<sms * MB200_WF_TRACE_BLOCKS, kThreads, 0, sa>>>(M, B, B.qs, B.counters + CNT_S, B.counters + 4, wf_leaf_thresh());
            if (side) { cudaEventRecord(side->any_done, sa); any_pending = true; }
            uint32_t* tq = qin; qin = qout; qout = tq; const int tc = cin; cin = cout; cout = tc;
        }
        if (any_pending) cudaStreamWaitEvent(st, side->any_done, 0);
        if (c->filter == MB200_FILTER_GAUSSIAN) wf_film_kernel<MB200_FILTER_GAUSSIAN><<<sms * 8, kThreads, 0, st>>>(P, B, pix0, npb, 0, c->spp, 1, 1);
        else                                    wf_film_kernel<MB200_FILTER_BOX><<<sms * 8, kThreads, 0, st>>>(P, B, pix0, npb, 0, c->spp, 1, 1);
    }
    return mb200_check_launch();
}

int mb200_trans_mesh_shade_fwd(const mb200_cfg* c, const mb200_trans* t, const mb200_mesh_desc* md, const void* mesh_buf,
                               const float* a, const float* r, const float* m, const float* n_opt,
                               const float* env4, const float* hier, const mb200_hier_desc* d, float* partials, void* stream) {
    RenderParams P; int rc = mesh_render_params(c, a, r, m, n_opt, env4, hier, d, P);
    if (rc) return rc;
    if ((rc = fill_trans(t, P)) != MB200_OK) return rc;
    MeshView M; rc = make_view(md, mesh_buf, M);
    if (rc) return rc;
    if (!partials) return MB200_EINVAL;
    if (c->flags & MB200_FLAG_AD_WEIGHTS) return MB200_EUNSUPPORTED;      // the reference never differentiates TransBSDF
    P.prows = mb200_fwd_partial_rows(c, &P.prow0); P.partials = partials;
    const int grid = grid_for(P.prows * P.W);
    cudaStream_t st = (cudaStream_t)stream;
    if (c->filter == MB200_FILTER_GAUSSIAN) launch_mesh_fwd(mesh_fwd_kernel<MB200_FILTER_GAUSSIAN, false, true>, c->filter, grid, st, P, M);
    else                                    launch_mesh_fwd(mesh_fwd_kernel<MB200_FILTER_BOX, false, true>, c->filter, grid, st, P, M);
    return mb200_check_launch();
}

int mb200_mesh_shade_bwd(const mb200_cfg* c, const mb200_mesh_desc* md, const void* mesh_buf,
                         const float* a, const float* r, const float* m, const float* n_opt,
                         const float* env4, const float* hier, const mb200_hier_desc* d, const float* gadj,
                         float* g_a, float* g_r, float* g_m, float* g_n, float* g_env4, int n_env_slabs, void* stream) {
    RenderParams P; int rc = mesh_render_params(c, a, r, m, n_opt, env4, hier, d, P);
    if (rc) return rc;
    MeshView M; rc = make_view(md, mesh_buf, M);
    if (rc) return rc;
    if (!gadj || (g_env4 && n_env_slabs < 1)) return MB200_EINVAL;
    P.env_slabs = g_env4 ? n_env_slabs : 1; P.env_slab_stride = (long long)d->res_x * d->res_y;
    P.prow0 = c->row0; P.prows = c->rows;
    P.gadj = reinterpret_cast<const float4*>(gadj); P.grows = mb200_bwd_gadj_rows(c, &P.grow0);
    P.g_a = g_a; P.g_r = g_r; P.g_m = g_m; P.g_n = g_n; P.g_env4 = reinterpret_cast<float4*>(g_env4);
    const bool want_n = g_n != nullptr && !c->use_mesh_normal && n_opt != nullptr;
    const bool want_mat = g_a || g_r || g_m || want_n;
    const bool want_env = g_env4 != nullptr;
    if (!want_mat && !want_env) return MB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return c->filter == MB200_FILTER_GAUSSIAN ? launch_mesh_bwd<MB200_FILTER_GAUSSIAN>(P, M, want_mat, want_n, want_env, st)
                                              : launch_mesh_bwd<MB200_FILTER_BOX>(P, M, want_mat, want_n, want_env, st);
}

#ifndef MB200_WF_BWD_BATCH_LOG2
#define MB200_WF_BWD_BATCH_LOG2 23
#endif
static long long wf_bwd_batch_pixels(const mb200_cfg* c) { const long long bp = (1ll << MB200_WF_BWD_BATCH_LOG2) / c->spp; return bp < 1 ? 1 : bp; }
size_t mb200_mesh_bwd_wf_scratch_bytes(const mb200_cfg* c) {
    if (!c || c->spp <= 0 || c->spp > (1 << MB200_WF_BWD_BATCH_LOG2)) return 0;
    const long long npix = (long long)c->rows * c->W, bp = wf_bwd_batch_pixels(c);
    const int mv = (c->max_depth - 1 < kMaxVerts ? c->max_depth - 1 : kMaxVerts);
    return wf_bwd_scratch_bytes((npix < bp ? npix : bp) * c->spp, mv < 1 ? 1 : mv);
}

int mb200_mesh_shade_bwd_wf(const mb200_cfg* c, const mb200_mesh_desc* md, const void* mesh_buf,
                            const float* a, const float* r, const float* m, const float* n_opt,
                            const float* env4, const float* hier, const mb200_hier_desc* d, const float* gadj,
                            float* g_a, float* g_r, float* g_m, float* g_n, float* g_env4, int n_env_slabs,
                            void* scratch, size_t scratch_bytes, const void* primary_index, void* stream) {
    RenderParams P; int rc = mesh_render_params(c, a, r, m, n_opt, env4, hier, d, P);
    if (rc) return rc;
    MeshView M; rc = make_view(md, mesh_buf, M);
    if (rc) return rc;
    if (!gadj || !scratch || (g_env4 && n_env_slabs < 1)) return MB200_EINVAL;
    if (c->spp > (1 << MB200_WF_BWD_BATCH_LOG2)) return MB200_EUNSUPPORTED;
    P.env_slabs = g_env4 ? n_env_slabs : 1; P.env_slab_stride = (long long)d->res_x * d->res_y;
    P.prow0 = c->row0; P.prows = c->rows;
    P.gadj = reinterpret_cast<const float4*>(gadj); P.grows = mb200_bwd_gadj_rows(c, &P.grow0);
    P.g_a = g_a; P.g_r = g_r; P.g_m = g_m; P.g_n = g_n; P.g_env4 = reinterpret_cast<float4*>(g_env4);
    const bool want_n = g_n != nullptr && !c->use_mesh_normal && n_opt != nullptr;
    const bool want_mat = g_a || g_r || g_m || want_n;
    const bool want_env = g_env4 != nullptr;
    if (!want_mat && !want_env) return MB200_OK;
    const long long npix = (long long)P.prows * P.W, bp = wf_bwd_batch_pixels(c);
    int max_verts = (c->max_depth - 1 < kMaxVerts ? c->max_depth - 1 : kMaxVerts); if (max_verts < 0) max_verts = 0;
    const long long nb_max = (npix < bp ? npix : bp) * c->spp;
    if (scratch_bytes < wf_bwd_scratch_bytes(nb_max, max_verts < 1 ? 1 : max_verts)) return MB200_EINVAL;
    WfBuf B = wf_carve_bwd(scratch, nb_max, max_verts < 1 ? 1 : max_verts);
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = mb200_sm_count();
    WfSide* side = wf_side();
    for (long long pix0 = 0; pix0 < npix; pix0 += bp) {
        const int npb = (int)((npix - pix0) < bp ? (npix - pix0) : bp);
        const int nb = npb * c->spp;
        if (c->filter == MB200_FILTER_GAUSSIAN) wf_gen_bwd_kernel<MB200_FILTER_GAUSSIAN><<<sms * 8, 256, 0, st>>>(P, B, pix0, nb);
        else                                    wf_gen_bwd_kernel<MB200_FILTER_BOX><<<sms * 8, 256, 0, st>>>(P, B, pix0, nb);
        uint32_t* qin = B.qa; uint32_t* qout = B.qb; int cin = CNT_A, cout = CNT_B;
        bool any_pending = false;
        for (int it = 0; it <= max_verts; ++it) {
            cudaMemsetAsync(B.counters + 3, 0, 4, st);
            if (it == 0 && primary_index) {
                const PIdxView I = pidx_view(const_cast<void*>(primary_index), c->H, c->W);
                cudaMemsetAsync(B.counters + CNT_B, 0, 4 * kBins, st);                 // fallback queue (rays the index cannot serve)
                wf_primary_kernel<<<sms * 8, 256, 0, st>>>(P, M, B, I, pix0, nb, B.qb, B.counters + CNT_B);
                wf_trace_kernel<0><<<sms * MB200_WF_TRACE_BLOCKS, kThreads, 0, st>>>(M, B, B.qb, B.counters + CNT_B, B.counters + 3, wf_leaf_thresh());
            } else
            wf_trace_kernel<0><<<sms * MB200_WF_TRACE_BLOCKS, kThreads, 0, st>>>(M, B, qin, B.counters + cin, B.counters + 3, wf_leaf_thresh());
            if (any_pending) { cudaStreamWaitEvent(st, side->any_done, 0); any_pending = false; }
            cudaMemsetAsync(B.counters + cout, 0, 4 * kBins, st);
            cudaMemsetAsync(B.counters + CNT_S, 0, 4 * kBins, st);
            if (want_mat && want_env)  wf_shade_bwd_kernel<true, true><<<sms * 4, kThreads, 0, st>>>(P, M, B, qin, qout, cin, cout);
            else if (want_mat)         wf_shade_bwd_kernel<true, false><<<sms * 4, kThreads, 0, st>>>(P, M, B, qin, qout, cin, cout);
            else                       wf_shade_bwd_kernel<false, true><<<sms * 4, kThreads, 0, st>>>(P, M, B, qin, qout, cin, cout);
            cudaStream_t sa = side ? side->s : st;
            if (side) { cudaEventRecord(side->shaded, st); cudaStreamWaitEvent(sa, side->shaded, 0); }
            cudaMemsetAsync(B.counters + 4, 0, 4, sa);
            wf_trace_kernel<2><<<sms * MB200_WF_TRACE_BLOCKS, kThreads, 0, sa>>>(M, B, B.qs, B.counters + CNT_S, B.counters + 4, wf_leaf_thresh());
            if (want_mat && want_env)  wf_apply_bwd_kernel<true, true><<<sms * 4, 256, 0, sa>>>(P, B);
            else if (want_mat)         wf_apply_bwd_kernel<true, false><<<sms * 4, 256, 0, sa>>>(P, B);
            else                       wf_apply_bwd_kernel<false, true><<<sms * 4, 256, 0, sa>>>(P, B);
            if (side) { cudaEventRecord(side->any_done, sa); any_pending = true; }
            uint32_t* tq = qin; qin = qout; qout = tq; const int tc = cin; cin = cout; cout = tc;
        }
        if (any_pending) cudaStreamWaitEvent(st, side->any_done, 0);
        if (want_mat) {
            if (want_n) wf_walk_kernel<true><<<sms * 6, kThreads, 0, st>>>(P, B, nb);
            else        wf_walk_kernel<false><<<sms * 6, kThreads, 0, st>>>(P, B, nb);
        }
    }
    return mb200_check_launch();
}

int mb200_mesh_intersect(const mb200_mesh_desc* md, const void* mesh_buf, const float* o, const float* d, const float* maxt,
                         int n, int any_hit, int32_t* out_tri, float* out_tuv, void* stream) {
    MeshView M; int rc = make_view(md, mesh_buf, M);
    if (rc) return rc;
    if (!o || !d || !out_tri || n < 0) return MB200_EINVAL;
    if (n == 0) return MB200_OK;
    mesh_intersect_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(M, o, d, maxt, n, any_hit, out_tri, out_tuv);
    return mb200_check_launch();
}

int mb200_mesh_primary(const mb200_cfg* c, const mb200_mesh_desc* md, const void* mesh_buf, float jx, float jy,
                       float* gpos, float* gnrm, int32_t* tri, int32_t* flat, void* stream) {
    if (!c || !gpos || !gnrm || c->H <= 0 || c->W <= 0) return MB200_EINVAL;
    MeshView M; int rc = make_view(md, mesh_buf, M);
    if (rc) return rc;
    RenderParams P; memset(&P, 0, sizeof(P));
    for (int i = 0; i < 16; ++i) { P.cam.view[i] = c->view[i]; P.cam.proj[i] = c->proj[i]; P.cam.c2w[i] = c->cam_to_world[i]; }
    P.cam.tan_half_fov_x = c->tan_half_fov_x; P.cam.H = c->H; P.cam.W = c->W;
    P.cam.stride = (c->flags & MB200_FLAG_ROW_STRIDE_H) ? c->H : c->W;
    P.H = c->H; P.W = c->W;
    const int n = c->H * c->W;
    mesh_primary_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, M, jx, jy, reinterpret_cast<float4*>(gpos), reinterpret_cast<float4*>(gnrm), tri, flat);
    return mb200_check_launch();
}

}  // extern "C"
