// mb200_shade.cuh — per-pixel context and the per-sample forward shading function of the G-buffer kernels
// (device functions only: included by mb200_render.cu and by the host-emulation build of tests/host_emul, which
// compiles this very code with g++ to check its decisions against the oracle without a GPU).
#pragma once
#include "mb200_render_common.cuh"

namespace {

// what one sample decided (debug / parity instrumentation; compiled out of the production kernels)
struct SampleDbg {
    uint32_t ox, oy; long long flat; int lobe; int em_i00, bs_i00; float3 d_em, d_bs;
};

// Shared-memory staging of the emitter's sampling pyramid (levels >= P.hier.smem_from; ALL levels and the float4 texels for the
// small envmaps the reference really optimises: 16x32 learned, envmaps/0.hdr 32x16) — north_star: "the envmap marginal/conditional
// CDFs and mip levels are staged in shared memory".  Called by every thread of the CTA before its pixel loop.
// The copy itself is the bulk-copy engine's (cp.async.bulk global -> shared, completion counted on an mbarrier — the TMA path for
// 1-D data): ONE thread issues two copies per CTA and everybody waits on the barrier, instead of ~11 load/store iterations by all 256
// threads of each of the grid's 4 736 CTAs (the plain loop was 2.5 % of the adjoint kernel's executed instructions, profiles/r5l).
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t mb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void stage_bulk(uint32_t bar, float4* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(mb_smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
#endif
__device__ __forceinline__ StagedEnv stage_env(const RenderParams& P, float4* dyn) {
    StagedEnv S{nullptr, nullptr};
    const int nf = P.hier.smem_floats, nh4 = (nf + 3) >> 2;
    if (nf <= 0 && P.env_smem_texels <= 0) return S;
    const float* hsrc = P.hier.data + P.hier.smem_off0;                  // level offsets are 16-byte aligned
    float4* t = dyn + nh4;
#if defined(__CUDACC__)
    __shared__ __align__(8) unsigned long long s_bar;
    const uint32_t bar = mb_smem_u32(&s_bar);
    const uint32_t hbytes = (uint32_t)(nf >> 2) * 16u, tbytes = (uint32_t)P.env_smem_texels * 16u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(hbytes + tbytes) : "memory");
        if (hbytes) stage_bulk(bar, dyn, hsrc, hbytes);
        if (tbytes) stage_bulk(bar, t, P.env.tex, tbytes);
    }
    if (threadIdx.x < (nf & 3)) reinterpret_cast<float*>(dyn)[(nf & ~3) + threadIdx.x] = __ldg(hsrc + (nf & ~3) + threadIdx.x);   // < 16-byte tail
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(0) : "memory");
    __syncthreads();                                                      // the tail stores
#else       // host emulation build: plain copies
    for (int i = threadIdx.x; i < nf; i += blockDim.x) reinterpret_cast<float*>(dyn)[i] = hsrc[i];
    for (int i = threadIdx.x; i < P.env_smem_texels; i += blockDim.x) t[i] = P.env.tex[i];
    __syncthreads();
#endif
    if (nf > 0) S.hier = reinterpret_cast<const float*>(dyn);
    if (P.env_smem_texels > 0) S.tex = t;
    return S;
}

struct PixelCtx {
    bool valid; long long flat; Material mt; float3 view; Frame fgeo, fshade; TransMat tm;
};

// NMAP = false: the shading normal IS the G-buffer normal (use_mesh_normal, what the reference runs by default) — known at compile
// time, so mt.n, the second frame and everything derived from them share registers with the geometric ones (12 fewer live values in
// kernels that sit at their register caps).
template <bool TRANS = false, bool NMAP = true>
__device__ __forceinline__ PixelCtx load_pixel(const RenderParams& P, int gpix) {
    PixelCtx c;
    const float4 gp = __ldg(P.gpos + gpix), gn = __ldg(P.gnrm + gpix);
    c.valid = gp.w != 0.f;
    const float3 p = f3(gp.x, gp.y, gp.z), ng = f3(gn.x, gn.y, gn.z);
    c.flat = 0; c.mt.a = f3(0, 0, 0); c.mt.r = 1.f; c.mt.m = 0.f; c.mt.n = ng; c.view = f3(0, 0, 1);
    if (c.valid) {
        c.flat = texel_index(P.cam, p);
        c.mt.a = f3(__ldg(P.a + 3 * c.flat), __ldg(P.a + 3 * c.flat + 1), __ldg(P.a + 3 * c.flat + 2));
        c.mt.r = __ldg(P.r + c.flat); c.mt.m = __ldg(P.m + c.flat);
        if (NMAP && !P.use_mesh_normal && P.n_opt)
            c.mt.n = f3(__ldg(P.n_opt + 3 * c.flat), __ldg(P.n_opt + 3 * c.flat + 1), __ldg(P.n_opt + 3 * c.flat + 2));
        c.view = xnormalize3(xsub3(f3(P.cam.c2w[3], P.cam.c2w[7], P.cam.c2w[11]), p));
        if (TRANS) c.tm = trans_fetch(P.cam, P.trans, c.flat, c.view, ng, p);
    }
    c.fgeo = make_frame(ng); c.fshade = NMAP ? make_frame(c.mt.n) : c.fgeo;
    return c;
}

// ---------------------------------------------------------------- one forward sample
template <bool AD_W, bool TRANS = false, bool DBG = false>
__device__ __forceinline__ float3 shade_sample(const RenderParams& P, const PixelCtx& c, int px, int py, uint32_t lane_id,
                                               float& jx, float& jy, SampleDbg* dbg = nullptr, const StagedEnv S = StagedEnv{nullptr, nullptr}) {
    Pcg32 rng; rng.seed(P.seed, lane_id);
    jx = rng.next_float(); jy = rng.next_float();
    if (DBG) { dbg->ox = dbg->oy = 0; dbg->flat = -1; dbg->lobe = -1; dbg->em_i00 = dbg->bs_i00 = -1; dbg->d_em = dbg->d_bs = f3(0.f, 0.f, 0.f); }
    if (!c.valid) {
        const float3 d = primary_dir(P.cam, XADD((float)px, jx), XADD((float)py, jy));
        float u, v; dir_to_uv(d, u, v);
        const Bilerp bm = env_lookup(P.env, u, v);
        if (DBG) { dbg->d_bs = d; dbg->bs_i00 = (int)bm.i00; }
        return env_value(P.env, bm, S.tex);
    }
    float3 L = f3(0.f, 0.f, 0.f);
    if (P.max_depth < 2) return L;
    const float uex = rng.next_float(), uey = rng.next_float();
    const float s1 = rng.next_float();
    const float s2x = rng.next_float(), s2y = rng.next_float();
    // (the russian-roulette draw that follows is never consumed: rr_depth 5 > max_depth)
    // ---- emitter sampling
    const EmSample em = env_sample_direction(P.hier, P.env, uex, uey, S.hier);
    if (DBG) { dbg->ox = em.ox; dbg->oy = em.oy; dbg->flat = c.flat; dbg->em_i00 = (int)em.b.i00; dbg->d_em = em.d; }
    if (em.pdf != 0.f) {
        const float3 le = env_value(P.env, em.b, S.tex);
        const BsdfVal fv = TRANS ? trans_eval_brdf(em.d, c.view, c.mt, c.tm, P.trans) : eval_brdf(em.d, c.view, c.mt);
        const float k = mis_weight(em.pdf, fv.pdf) / em.pdf;
        L = fv.f * le * k;
    }
    // ---- BSDF sampling
    const BsdfSample bs = TRANS ? trans_sample_brdf(s1, s2x, s2y, c.view, c.mt, c.tm, P.trans, c.fshade) : sample_brdf(s1, s2x, s2y, c.view, c.mt, c.fshade);
    const float3 d_bs = (P.flags & MB200_FLAG_WO_WORLD_QUIRK) ? to_world(c.fgeo, bs.wi) : bs.wi;
    if (DBG) {      // the envmap cell of the BSDF-sampled direction, whether or not the sample carries weight
        float u, v; dir_to_uv(d_bs, u, v);
        dbg->lobe = bs.lobe; dbg->d_bs = d_bs; dbg->bs_i00 = (int)env_lookup(P.env, u, v).i00;
    }
    float3 w_bs = bs.weight;
    if (AD_W) {
        const BsdfVal b2 = eval_brdf(d_bs, c.view, c.mt);
        if (b2.pdf > 0.f) w_bs = b2.f * (1.f / b2.pdf);
    }
    if (fmax3(w_bs.x, w_bs.y, w_bs.z) != 0.f && bs.pdf > 0.f) {
        float u, v; dir_to_uv(d_bs, u, v);
        const float em_pdf = env_pdf_direction(P.hier, P.env, d_bs, u, v, S.hier);
        const float3 le = env_value(P.env, env_lookup(P.env, u, v), S.tex);
        L = L + w_bs * le * mis_weight(bs.pdf, em_pdf);
    }
    return L;
}


}  // namespace
