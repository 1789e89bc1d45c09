// mb200_render.cu — fused forward / adjoint envmap shading kernels (sm_100a) + their C-ABI entry points.
//
// Mapping: ONE WARP PER PIXEL, lanes stride over the pixel's samples (lane l takes s = l, l+32, ...).
// All per-pixel data (G-buffer, texel fetch, frames) is warp-uniform, so the samples of a pixel never
// diverge on geometry; per-sample state (PCG32 stream, hierarchy descent, BSDF evals) lives in registers.
//
// Forward (gaussian film): every sample contributes to a 5x5 pixel footprint.  Instead of 100 float
// atomics per sample the warp stages (wx[5], wy[5], L.rgb) of its 32 in-flight samples in shared memory
// and lanes 0..24 each own one tap: the pixel's 25 (rgb,w) tap sums are written once, coalesced, to
// `partials` and mb200_film_develop gathers them in a fixed order -> bitwise deterministic and
// shard-invariant.  Forward (box film): 3 warp-shuffle reductions per pixel.
//
// Adjoint: second render with seed_grad.  Film adjoint is a 5x5 gather of
// G = grad/W staged per warp in shared memory; material gradients are reduced per pixel in registers
// with warp shuffles and leave as one RED per channel; envmap gradients leave as 16-byte vector
// reductions (red.global.add.v4.f32) into a float4 texel grid.
#include "mb200_render_common.cuh"
#include "mb200_shade.cuh"

namespace {

// dynamic shared memory the env staging may use per CTA: static + dynamic stays under the 48 KB that needs no opt-in
// (forward: 20 KB of tap records; adjoint: 3.2 KB of film cotangents)
constexpr size_t kStageBudgetFwd = 27 * 1024, kStageBudgetBwd = 44 * 1024;


// ---------------------------------------------------------------- forward kernel
// LPP = lanes per pixel (32 or 8): a warp shades 32 / LPP pixels at a time, lane l = (pixel group l / LPP, sample slot l % LPP), and a
// lane walks the samples slot, slot + LPP, ... of ITS pixel.  LPP = 8 amortises everything that is per pixel — G-buffer and map
// loads, the texel projection, the view vector, two frames, index arithmetic, the final stores: ~170 instructions a lane, 6-8 % of
// the kernel at 64 spp with a whole warp per pixel — over 4x as many samples per lane, and keeps all lanes busy from 8 spp up
// (a warp per pixel idles lanes below 32 spp).  The film-tap reduction is unchanged in cost: still 32 staged records per batch,
// lanes 0..24 each own one tap, now with one accumulator per pixel group.
template <int FILTER, bool AD_W, bool TRANS = false, int LPP = 32, bool NMAP = true>
__global__ void __launch_bounds__(kThreads, MB_MIN_BLOCKS_FWD) shade_fwd_kernel(const __grid_constant__ RenderParams P) {
    constexpr int PPW = 32 / LPP;
    // Staged tap records of the warp's 32 in-flight samples, STRUCTURE OF ARRAYS: rows 0..4 = wx[i][record], rows 5..9 = wy[j][record]
    // (row stride 36 floats: the five rows a 128-bit load touches start 4 banks apart), then one float4 (L, 1) per record.  A tap
    // lane fetches the weights of FOUR records with two 128-bit loads (26 instructions per four records instead of 32 with one
    // 16-float record per sample and two scalar loads per record and tap).
    constexpr int kTapRow = 36, kTapWarp = 10 * kTapRow + 32 * 4;
    __shared__ __align__(16) float s_rec[FILTER == MB200_FILTER_GAUSSIAN ? kWarpsPerBlock * kTapWarp : 4];
    extern __shared__ float4 s_dyn[];
    const StagedEnv S = stage_env(P, s_dyn);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / LPP, sl = lane % LPP;
    float* rec = s_rec + (FILTER == MB200_FILTER_GAUSSIAN ? warp * kTapWarp : 0);
    float4* recL = reinterpret_cast<float4*>(rec + 10 * kTapRow);
    const int npix = P.prows * P.W;
    const int ti = lane % 5, tj = lane / 5;              // tap owned by this lane (lanes 0..24)
    for (int pix0 = (blockIdx.x * kWarpsPerBlock + warp) * PPW; pix0 < npix; pix0 += gridDim.x * kWarpsPerBlock * PPW) {
        const bool pix_ok = pix0 + grp < npix;           // (a ragged last warp: the extra groups shade nothing)
        const int pix = pix_ok ? pix0 + grp : npix - 1;
        const int py = P.prow0 + pix / P.W, px = pix % P.W;
        const int gpix = py * P.W + px;
        const PixelCtx c = load_pixel<TRANS, NMAP>(P, gpix);
        float4 acc[PPW];
#pragma unroll
        for (int g = 0; g < PPW; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < P.spp; s0 += LPP) {
            const int s = s0 + sl;
            float3 L = f3(0.f, 0.f, 0.f); float jx = 0.f, jy = 0.f;
            const bool act = pix_ok && s < P.spp;
            if (act) L = shade_sample<AD_W, TRANS>(P, c, px, py, (uint32_t)gpix * (uint32_t)P.spp + (uint32_t)s, jx, jy, nullptr, S);
            if (FILTER == MB200_FILTER_GAUSSIAN) {
                float wx[5], wy[5]; film_taps(jx, wx); film_taps(jy, wy);
                if (!act) { wx[0] = wx[1] = wx[2] = wx[3] = wx[4] = 0.f; }
#pragma unroll
                for (int i = 0; i < 5; ++i) { rec[i * kTapRow + lane] = wx[i]; rec[(5 + i) * kTapRow + lane] = wy[i]; }
                recL[lane] = make_float4(L.x, L.y, L.z, 1.f);
                __syncwarp();
                if (lane < MB200_FILM_TAPS) {
                    // inactive lanes staged wx = 0 -> w = 0: always 32 records (LPP per pixel group), fully unrollable; records are
                    // added in their order, as before
                    const float* rt = rec + ti * kTapRow; const float* ru = rec + (5 + tj) * kTapRow;
#pragma unroll
                    for (int g = 0; g < PPW; ++g) {
#pragma unroll 2
                        for (int kk = 0; kk < LPP; kk += 4) {
                            const int k = g * LPP + kk;
                            const float4 a4 = *reinterpret_cast<const float4*>(rt + k), b4 = *reinterpret_cast<const float4*>(ru + k);
                            const float w4[4] = {a4.x * b4.x, a4.y * b4.y, a4.z * b4.z, a4.w * b4.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 l4 = recL[k + q];
                                acc[g].x = fmaf(w4[q], l4.x, acc[g].x); acc[g].y = fmaf(w4[q], l4.y, acc[g].y); acc[g].z = fmaf(w4[q], l4.z, acc[g].z);
                                acc[g].w += w4[q];
                            }
                        }
                    }
                }
                __syncwarp();
            } else {
                acc[0].x += L.x; acc[0].y += L.y; acc[0].z += L.z;
            }
        }
        if (FILTER == MB200_FILTER_GAUSSIAN) {
            if (lane < MB200_FILM_TAPS) {
#pragma unroll
                for (int g = 0; g < PPW; ++g)
                    if (pix0 + g < npix) reinterpret_cast<float4*>(P.partials)[(size_t)(pix0 + g) * MB200_FILM_TAPS + lane] = acc[g];
            }
        } else {
#pragma unroll
            for (int o = LPP / 2; o > 0; o >>= 1) {
                acc[0].x += __shfl_xor_sync(0xffffffffu, acc[0].x, o);
                acc[0].y += __shfl_xor_sync(0xffffffffu, acc[0].y, o);
                acc[0].z += __shfl_xor_sync(0xffffffffu, acc[0].z, o);
            }
            if (sl == 0 && pix_ok) reinterpret_cast<float4*>(P.partials)[pix] = make_float4(acc[0].x, acc[0].y, acc[0].z, (float)P.spp);
        }
    }
}

// ---------------------------------------------------------------- develop: gather taps in a fixed order, rgb / w
template <int FILTER>
__global__ void film_develop_kernel(const float4* __restrict__ partials, int H, int W, int prow0, int prows,
                                    int row0, int rows, float* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * W) return;
    const int qy = row0 + i / W, qx = i % W;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (FILTER == MB200_FILTER_GAUSSIAN) {
#pragma unroll
        for (int t = 0; t < MB200_FILM_TAPS; ++t) {
            const int sy = qy - (t / 5 - 2), sx = qx - (t % 5 - 2);
            if (sx < 0 || sx >= W || sy < prow0 || sy >= prow0 + prows) continue;
            const float4 q = __ldg(partials + ((size_t)(sy - prow0) * W + sx) * MB200_FILM_TAPS + t);
            acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
        }
    } else {
        acc = __ldg(partials + (size_t)(qy - prow0) * W + qx);
    }
    const float ws = acc.w == 0.f ? 1.f : acc.w;
    img[3 * (size_t)i] = __fdiv_rn(acc.x, ws); img[3 * (size_t)i + 1] = __fdiv_rn(acc.y, ws); img[3 * (size_t)i + 2] = __fdiv_rn(acc.z, ws);
}

// ---------------------------------------------------------------- film weights of a render (RNG only)
__global__ void __launch_bounds__(kThreads) film_weights_kernel(int W, int spp, uint32_t seed, int wrow0, int wrows, float* __restrict__ wpart) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npix = wrows * W;
    for (int pix = blockIdx.x * kWarpsPerBlock + warp; pix < npix; pix += gridDim.x * kWarpsPerBlock) {
        const int gpix = (wrow0 + pix / W) * W + pix % W;
        float acc[MB200_FILM_TAPS];
#pragma unroll
        for (int t = 0; t < MB200_FILM_TAPS; ++t) acc[t] = 0.f;
        for (int s = lane; s < spp; s += 32) {
            Pcg32 rng; rng.seed(seed, (uint32_t)gpix * (uint32_t)spp + (uint32_t)s);
            const float jx = rng.next_float(), jy = rng.next_float();
            float wx[5], wy[5]; film_taps(jx, wx); film_taps(jy, wy);
#pragma unroll
            for (int t = 0; t < MB200_FILM_TAPS; ++t) acc[t] = fmaf(wx[t % 5], wy[t / 5], acc[t]);
        }
        // reduce-scatter butterfly over the warp (the 25 tap sums padded to 32): at the step with partner distance h a lane keeps the
        // half of its values whose index bit equals its own lane bit and hands the other half to its partner, so after 5 steps
        // lane t holds the warp total of tap t: 31 shuffles instead of 25 x 5, and the store is coalesced
        float v[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] = t < MB200_FILM_TAPS ? acc[t] : 0.f;
#define MB_RS_STEP(HALF)                                                                              \
        {   const bool up = (lane & HALF) != 0;                                                       \
            _Pragma("unroll") for (int i = 0; i < HALF; ++i) {                                         \
                const float send = up ? v[i] : v[i + HALF], keep = up ? v[i + HALF] : v[i];            \
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, HALF);                                \
            } }
        MB_RS_STEP(16) MB_RS_STEP(8) MB_RS_STEP(4) MB_RS_STEP(2) MB_RS_STEP(1)
#undef MB_RS_STEP
        if (lane < MB200_FILM_TAPS) wpart[(size_t)pix * MB200_FILM_TAPS + lane] = v[0];
    }
}
// The same sums with T LANES PER PIXEL (T = 1: one thread per pixel) for images with enough pixels to fill the GPU that way: no
// 31-shuffle butterfly (which costs about as many instructions per sample as the RNG and the taps together at 64 spp), and the 25
// sums of a warp's 32 / T pixels leave through shared memory as coalesced rows (stride 25 is odd: conflict free).
// Measured at C2 (profiles/r6g_fw_sweep.log): one warp per pixel 159 us; T = 1 / 2 / 4 / 8 -> 105 / 108 / 125 / 138 us; capping the
// residency (4 CTAs per SM, persistent) so that the loss kernels beside it always find a free CTA slot: no gain (step 2.527 vs 2.514 ms).
template <int T>
__global__ void __launch_bounds__(kThreads) film_weights_px_kernel(int W, int spp, uint32_t seed, int wrow0, int wrows, float* __restrict__ wpart) {
    constexpr int kPixPerWarp = 32 / T, kPixPerCta = kThreads / T;
    __shared__ float s_out[kPixPerCta * MB200_FILM_TAPS];
    const int npix = wrows * W, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & (T - 1);
    float* sw = s_out + warp * kPixPerWarp * MB200_FILM_TAPS;                  // this warp's pixels x 25
    for (int base = blockIdx.x * kPixPerCta; base < npix; base += gridDim.x * kPixPerCta) {
        const int wbase = base + warp * kPixPerWarp, pix = wbase + lane / T;
        float acc[MB200_FILM_TAPS];
#pragma unroll
        for (int t = 0; t < MB200_FILM_TAPS; ++t) acc[t] = 0.f;
        if (pix < npix) {
            const uint32_t gpix = (uint32_t)((wrow0 + pix / W) * W + pix % W);
            for (int s = sub; s < spp; s += T) {
                Pcg32 rng; rng.seed(seed, gpix * (uint32_t)spp + (uint32_t)s);
                const float jx = rng.next_float(), jy = rng.next_float();
                float wx[5], wy[5]; film_taps(jx, wx); film_taps(jy, wy);
#pragma unroll
                for (int t = 0; t < MB200_FILM_TAPS; ++t) acc[t] = fmaf(wx[t % 5], wy[t / 5], acc[t]);
            }
        }
#pragma unroll
        for (int h = 1; h < T; h <<= 1) {
#pragma unroll
            for (int t = 0; t < MB200_FILM_TAPS; ++t) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], h);
        }
        if (sub == 0) {
#pragma unroll
            for (int t = 0; t < MB200_FILM_TAPS; ++t) sw[(lane / T) * MB200_FILM_TAPS + t] = acc[t];
        }
        __syncwarp();
        const long long w0 = (long long)wbase * MB200_FILM_TAPS, wend = (long long)npix * MB200_FILM_TAPS;
#pragma unroll
        for (int i = 0; i < (kPixPerWarp * MB200_FILM_TAPS + 31) / 32; ++i) {
            const int k = i * 32 + lane;
            if (k < kPixPerWarp * MB200_FILM_TAPS && w0 + k < wend) wpart[w0 + k] = sw[k];
        }
        __syncwarp();
    }
}
// G[q] = grad[q] / W_q
template <int FILTER>
__global__ void film_adjoint_kernel(const float* __restrict__ wpart, int H, int W, int wrow0, int wrows, int grow0, int grows,
                                    float inv_spp, const float* __restrict__ grad, float4* __restrict__ gadj) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= grows * W) return;
    const int qy = grow0 + i / W, qx = i % W;
    float inv = inv_spp;
    if (FILTER == MB200_FILTER_GAUSSIAN) {
        float ws = 0.f;
#pragma unroll
        for (int t = 0; t < MB200_FILM_TAPS; ++t) {
            const int sy = qy - (t / 5 - 2), sx = qx - (t % 5 - 2);
            if (sx < 0 || sx >= W || sy < wrow0 || sy >= wrow0 + wrows) continue;
            ws += __ldg(wpart + ((size_t)(sy - wrow0) * W + sx) * MB200_FILM_TAPS + t);
        }
        inv = 1.f / (ws == 0.f ? 1.f : ws);
    }
    gadj[i] = make_float4(grad[3 * (size_t)i] * inv, grad[3 * (size_t)i + 1] * inv, grad[3 * (size_t)i + 2] * inv, 0.f);
}

// ---------------------------------------------------------------- adjoint kernel

// ENVAGG (envmap-gradient scatter): false = every lane issues its four 16-byte reductions into the CTA's L2-resident slab as it goes
// (the default: the kernel is instruction-issue bound and the L2 atomic units absorb the updates, profiles/r4b_envphase*.log);
// true = warp-aggregated (match.any + register peer reduction) and, for small maps, block-privatised in shared memory — what
// north_star prescribes; measured SLOWER on B200 (16x32 + sun: 1.55 / 1.64 ms against 1.40 ms) and kept selectable (MB200_ENV_SCATTER=agg).
// LPP: lanes per pixel, as in shade_fwd_kernel (each pixel group of a warp has its own 5x5 film cotangent in shared memory; the
// material gradients are reduced over the LPP lanes of a group).
template <int FILTER, bool WANT_MAT, bool WANT_N, bool WANT_ENV, bool ENVAGG = false, int LPP = 32, bool NMAP = true>
__global__ void __launch_bounds__(kThreads, MB_MIN_BLOCKS_BWD) shade_bwd_kernel(const __grid_constant__ RenderParams P) {
    constexpr int PPW = 32 / LPP;
    __shared__ float4 s_g[FILTER == MB200_FILTER_GAUSSIAN ? kWarpsPerBlock * PPW * MB200_FILM_TAPS : 1];
    extern __shared__ float4 s_dyn[];
    const StagedEnv S = stage_env(P, s_dyn);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / LPP, sl = lane % LPP;
    float4* gt = s_g + (FILTER == MB200_FILTER_GAUSSIAN ? (warp * PPW + grp) * MB200_FILM_TAPS : 0);
    const int npix = P.prows * P.W;
    // this CTA's privatised copy of the envmap-gradient grid (mb200_env_grad_slabs) and, for small maps, its shared-memory accumulator
    float4* const genv = WANT_ENV ? P.g_env4 + (long long)(blockIdx.x % P.env_slabs) * P.env_slab_stride : nullptr;
    float* senv = nullptr;
    if (WANT_ENV && ENVAGG && P.env_grad_smem) {
        senv = reinterpret_cast<float*>(s_dyn + ((P.hier.smem_floats + 3) >> 2) + P.env_smem_texels);
        for (int i = threadIdx.x; i < 3 * (int)P.env_slab_stride; i += blockDim.x) senv[i] = 0.f;
        __syncthreads();
    }
    for (int pix0 = (blockIdx.x * kWarpsPerBlock + warp) * PPW; pix0 < npix; pix0 += gridDim.x * kWarpsPerBlock * PPW) {
        const bool pix_ok = pix0 + grp < npix;
        const int pix = pix_ok ? pix0 + grp : npix - 1;
        const int py = P.prow0 + pix / P.W, px = pix % P.W;
        const int gpix = py * P.W + px;
        const PixelCtx c = load_pixel<false, NMAP>(P, gpix);
        float3 gbox = f3(0, 0, 0);
        if (FILTER == MB200_FILTER_GAUSSIAN) {
            __syncwarp();
            for (int t = sl; t < MB200_FILM_TAPS; t += LPP) {
                const int qy = py + (t / 5 - 2), qx = px + (t % 5 - 2);
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qx >= 0 && qx < P.W && qy >= P.grow0 && qy < P.grow0 + P.grows && qy >= 0 && qy < P.H)
                    g = __ldg(P.gadj + (size_t)(qy - P.grow0) * P.W + qx);
                gt[t] = g;
            }
            __syncwarp();
        } else {
            const float4 g = __ldg(P.gadj + (size_t)(py - P.grow0) * P.W + px);
            gbox = f3(g.x, g.y, g.z);
        }
        float3 ga = f3(0, 0, 0), gn = f3(0, 0, 0); float gr = 0.f, gm = 0.f;
        const BrdfPix bp = brdf_pixel_terms(c.view, c.mt);      // view / material terms of the BSDF: once per pixel, not per evaluation
        for (int s0 = 0; s0 < P.spp; s0 += LPP) {         // uniform trip count: the aggregated envmap scatter below is warp-collective
            const int s = s0 + sl;
            Bilerp bA, bB; float3 cA = f3(0, 0, 0), cB = f3(0, 0, 0); bool aA = false, aB = false;   // this lane's envmap-gradient updates
            bA.i00 = bB.i00 = 0; bA.w0x = bA.w1x = bA.w0y = bA.w1y = bB.w0x = bB.w1x = bB.w0y = bB.w1y = 0.f;
            do {
            if (s >= P.spp || !pix_ok) break;
            Pcg32 rng; rng.seed(P.seed, (uint32_t)gpix * (uint32_t)P.spp + (uint32_t)s);
            const float jx = rng.next_float(), jy = rng.next_float();
            // film adjoint: dl = sum_taps wx_i wy_j G[p + (i,j)]
            float3 dl = gbox;
            if (FILTER == MB200_FILTER_GAUSSIAN) {
                float wx[5], wy[5]; film_taps(jx, wx); film_taps(jy, wy);
                // One of the two outer taps of an axis is EXACTLY zero (film_taps: w[0] = 0 for j > 0.5, w[4] = 0 otherwise), and an
                // fmaf with a zero weight leaves its accumulator unchanged: the 4 x 4 taps that can be non-zero give the same sum, term
                // for term in the same order, with 16 cotangent loads and 60 FMAs instead of 25 and 90.
                const int ox = jx > 0.5f ? 1 : 0, oy = jy > 0.5f ? 1 : 0;
                float vx[4], vy[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { vx[k] = ox ? wx[k + 1] : wx[k]; vy[k] = oy ? wy[k + 1] : wy[k]; }
                const float4* g0 = gt + (oy * 5 + ox);
                dl = f3(0, 0, 0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float3 row = f3(0, 0, 0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 g = g0[j * 5 + i];
                        row.x = fmaf(vx[i], g.x, row.x); row.y = fmaf(vx[i], g.y, row.y); row.z = fmaf(vx[i], g.z, row.z);
                    }
                    dl.x = fmaf(vy[j], row.x, dl.x); dl.y = fmaf(vy[j], row.y, dl.y); dl.z = fmaf(vy[j], row.z, dl.z);
                }
            }
            if (!c.valid) {
                if (WANT_ENV) {
                    const float3 d = primary_dir(P.cam, XADD((float)px, jx), XADD((float)py, jy));
                    float u, v; dir_to_uv(d, u, v);
                    bA = env_lookup(P.env, u, v); cA = dl; aA = true;
                    if (!ENVAGG) env_scatter(genv, P.env.Wi, bA, cA);
                }
                break;
            }
            if (P.max_depth < 2) break;
            const float uex = rng.next_float(), uey = rng.next_float();
            const float s1 = rng.next_float();
            const float s2x = rng.next_float(), s2y = rng.next_float();
            // ---- emitter term: L1 = f(d_em) * Le/pdf * mis
            const EmSample em = env_sample_direction(P.hier, P.env, uex, uey, S.hier);
            if (em.pdf != 0.f) {
                BrdfGradCtx gc;
                const BsdfVal fv = WANT_MAT ? eval_brdf_ctx<WANT_N>(em.d, c.view, c.mt, bp, gc) : eval_brdf(em.d, c.view, c.mt);
                const float k = mis_weight(em.pdf, fv.pdf) / em.pdf;
                if (WANT_MAT) {
                    const float3 le = env_value(P.env, em.b, S.tex);
                    const BsdfGrad bg = brdf_grad_apply<WANT_N>(gc, em.d, c.view, c.mt, bp, dl * le * k);
                    ga = ga + bg.ga; gr += bg.gr; gm += bg.gm; if (WANT_N) gn = gn + bg.gn;
                }
                if (WANT_ENV) { bA = em.b; cA = dl * fv.f * k; aA = true; if (!ENVAGG) env_scatter(genv, P.env.Wi, bA, cA); }
            }
            // ---- BSDF term: L2 = f(d_bs)/detach(p2) * Le(d_bs) * mis.  Of the lobe sample itself only the direction and its pdf
            // are needed (the primal weight f/(pdf+eps) only where the re-evaluated pdf is 0: a rare fallback, evaluated lazily)
            int lobe;
            const float3 wi_bs = sample_lobe_direction(s1, s2x, s2y, c.view, c.mt.r, c.fshade, lobe);
            const float pdf_s = eval_brdf_pdf(wi_bs, c.view, c.mt, bp);
            const float bs_pdf = pdf_s > 0.f ? pdf_s : 0.f;
            const float3 d_bs = (P.flags & MB200_FLAG_WO_WORLD_QUIRK) ? to_world(c.fgeo, wi_bs) : wi_bs;
            BrdfGradCtx gc2;
            const BsdfVal b2 = WANT_MAT ? eval_brdf_ctx<WANT_N>(d_bs, c.view, c.mt, bp, gc2) : eval_brdf(d_bs, c.view, c.mt);
            float3 w_bs;
            if (b2.pdf > 0.f) w_bs = b2.f * (1.f / b2.pdf);
            else {
                const BsdfVal bv = eval_brdf(wi_bs, c.view, c.mt);
                w_bs = bv.pdf > 1e-6f ? bv.f * frcp(bv.pdf + 1e-6f) : f3(0.f, 0.f, 0.f);
            }
            if (fmax3(w_bs.x, w_bs.y, w_bs.z) != 0.f && bs_pdf > 0.f) {
                float u, v; dir_to_uv(d_bs, u, v);
                const float mis = mis_weight(bs_pdf, env_pdf_direction(P.hier, P.env, d_bs, u, v, S.hier));
                const Bilerp bb = env_lookup(P.env, u, v);
                if (WANT_MAT && b2.pdf > 0.f) {
                    const float3 le = env_value(P.env, bb, S.tex);
                    const BsdfGrad bg = brdf_grad_apply<WANT_N>(gc2, d_bs, c.view, c.mt, bp, dl * le * (mis / b2.pdf));
                    ga = ga + bg.ga; gr += bg.gr; gm += bg.gm; if (WANT_N) gn = gn + bg.gn;
                }
                if (WANT_ENV) { bB = bb; cB = dl * w_bs * mis; aB = true; if (!ENVAGG) env_scatter(genv, P.env.Wi, bB, cB); }
            }
            } while (false);
            if (WANT_ENV && ENVAGG) {
                env_scatter_agg(genv, senv, P.env.Wi, bA, cA, aA);
                if (LPP < 32 || c.valid) env_scatter_agg(genv, senv, P.env.Wi, bB, cB, aB);   // (one pixel per warp: validity is warp-uniform)
            }
        }
        if (WANT_MAT) {
#pragma unroll
            for (int o = LPP / 2; o > 0; o >>= 1) {
                ga.x += __shfl_xor_sync(0xffffffffu, ga.x, o); ga.y += __shfl_xor_sync(0xffffffffu, ga.y, o);
                ga.z += __shfl_xor_sync(0xffffffffu, ga.z, o);
                gr += __shfl_xor_sync(0xffffffffu, gr, o); gm += __shfl_xor_sync(0xffffffffu, gm, o);
                if (WANT_N) {
                    gn.x += __shfl_xor_sync(0xffffffffu, gn.x, o); gn.y += __shfl_xor_sync(0xffffffffu, gn.y, o);
                    gn.z += __shfl_xor_sync(0xffffffffu, gn.z, o);
                }
            }
            if (sl == 0 && pix_ok && c.valid && P.max_depth >= 2) {
                if (P.g_a) { atomicAdd(P.g_a + 3 * c.flat, ga.x); atomicAdd(P.g_a + 3 * c.flat + 1, ga.y); atomicAdd(P.g_a + 3 * c.flat + 2, ga.z); }
                if (P.g_r) atomicAdd(P.g_r + c.flat, gr);
                if (P.g_m) atomicAdd(P.g_m + c.flat, gm);
                if (WANT_N && P.g_n) { atomicAdd(P.g_n + 3 * c.flat, gn.x); atomicAdd(P.g_n + 3 * c.flat + 1, gn.y); atomicAdd(P.g_n + 3 * c.flat + 2, gn.z); }
            }
        }
    }
    if (WANT_ENV && ENVAGG && senv) {  // flush the CTA's shared-memory gradient map into its slab: one 16-byte reduction per touched texel
        __syncthreads();
        for (int i = threadIdx.x; i < (int)P.env_slab_stride; i += blockDim.x) {
            const float x = senv[3 * i], y = senv[3 * i + 1], z = senv[3 * i + 2];
            if (x != 0.f || y != 0.f || z != 0.f) atomicAdd(genv + i, make_float4(x, y, z, 0.f));
        }
    }
}

// ---------------------------------------------------------------- debug: integer decisions per lane
__global__ void sample_indices_kernel(const __grid_constant__ RenderParams P, int32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)P.prows * P.W * P.spp;
    if (i >= n) return;
    const int s = (int)(i % P.spp); const long long pix = i / P.spp;
    const int py = P.prow0 + (int)(pix / P.W), px = (int)(pix % P.W);
    const int gpix = py * P.W + px;
    const float4 gp = __ldg(P.gpos + gpix);
    int32_t o0 = 0, o1 = 0, o2 = -1, o3 = -1;
    if (gp.w != 0.f && P.max_depth >= 2) {
        Pcg32 rng; rng.seed(P.seed, (uint32_t)gpix * (uint32_t)P.spp + (uint32_t)s);
        rng.next_float(); rng.next_float();
        const float uex = rng.next_float(), uey = rng.next_float();
        const float s1 = rng.next_float();
        const HSample hs = hier_sample(P.hier, uex, uey);
        o0 = (int32_t)hs.ox; o1 = (int32_t)hs.oy;
        o2 = (int32_t)texel_index(P.cam, f3(gp.x, gp.y, gp.z));
        o3 = s1 > 0.5f ? 1 : 0;
    }
    reinterpret_cast<int4*>(out)[i] = make_int4(o0, o1, o2, o3);
}


// per-sample decision record through the production shade_sample (DBG instantiation): 12 int32 words per lane =
// hier off.x, off.y, texel index, lobe, envmap cell of the emitter sample, envmap cell of the BSDF-sampled direction,
// float bits of the emitter direction (3), float bits of the BSDF-sampled direction (3)
__global__ void sample_record_kernel(const __grid_constant__ RenderParams P, int32_t* __restrict__ out, float* __restrict__ out_L) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)P.prows * P.W * P.spp;
    if (i >= n) return;
    const int s = (int)(i % P.spp); const long long pix = i / P.spp;
    const int py = P.prow0 + (int)(pix / P.W), px = (int)(pix % P.W);
    const int gpix = py * P.W + px;
    const PixelCtx c = load_pixel(P, gpix);
    SampleDbg d; float jx, jy;
    const float3 L = (P.flags & MB200_FLAG_AD_WEIGHTS)
        ? shade_sample<true, false, true>(P, c, px, py, (uint32_t)gpix * (uint32_t)P.spp + (uint32_t)s, jx, jy, &d)
        : shade_sample<false, false, true>(P, c, px, py, (uint32_t)gpix * (uint32_t)P.spp + (uint32_t)s, jx, jy, &d);
    int32_t* o = out + 12 * i;
    o[0] = (int32_t)d.ox; o[1] = (int32_t)d.oy; o[2] = (int32_t)d.flat; o[3] = d.lobe; o[4] = d.em_i00; o[5] = d.bs_i00;
    o[6] = __float_as_int(d.d_em.x); o[7] = __float_as_int(d.d_em.y); o[8] = __float_as_int(d.d_em.z);
    o[9] = __float_as_int(d.d_bs.x); o[10] = __float_as_int(d.d_bs.y); o[11] = __float_as_int(d.d_bs.z);
    if (out_L) { out_L[3 * i] = L.x; out_L[3 * i + 1] = L.y; out_L[3 * i + 2] = L.z; }
}

// Lanes per pixel of the G-buffer kernels.  Measured at C2 (64 spp; profiles/r5i_lpp_ab.log, r5j, r5k):
//   adjoint  warp per pixel 1.291 ms -> 8 lanes per pixel 1.207 ms (-6.5 %; 4 lanes: 1.209): default 8 up to 128 spp (above, the
//            per-pixel work no longer shows and a whole warp per pixel keeps the film cotangents of ONE pixel in shared memory);
//   forward  8 lanes per pixel needs four tap accumulators: 1.232 ms at the kernel's 64-register cap (spills) against 1.189 ms for a
//            warp per pixel, 1.195 ms at 80 registers / 3 CTAs per SM; 16 lanes per pixel (two accumulators) 1.165 ms against 1.173
//            ms -> default 16 for 16..128 spp, 8 below (so that no lane idles from 8 spp up), a warp per pixel above.
// MB200_LPP=8|32 forces both kernels, 16 the forward (measurements).
inline int lanes_per_pixel(int spp, bool adjoint) {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("MB200_LPP"); forced = e ? atoi(e) : 0; }
    if (forced == 8 || forced == 32 || (forced == 16 && !adjoint)) return forced;
    if (adjoint) return spp <= 128 ? 8 : 32;
    return spp < 16 ? 8 : (spp <= 128 ? 16 : 32);
}

template <int FILTER, int LPP>
int launch_bwd(const RenderParams& P, bool want_mat, bool want_n, bool want_env, size_t dyn, cudaStream_t st) {
    const int npix = (P.prows * P.W + (32 / LPP) - 1) / (32 / LPP);        // warps' worth of pixel groups
    const bool nmap = !P.use_mesh_normal && P.n_opt != nullptr;
#define MB_BWD_(M, N, E, A, NM) shade_bwd_kernel<FILTER, M, N, E, A, LPP, NM><<<persistent_grid(shade_bwd_kernel<FILTER, M, N, E, A, LPP, NM>, npix, dyn), kThreads, dyn, st>>>(P)
#define MB_BWD(M, N, E, A) do { if ((N) || nmap) MB_BWD_(M, N, E, A, true); else MB_BWD_(M, false, E, A, false); } while (0)
    const bool agg = want_env && env_scatter_aggregated();
    if (agg) {
        if (want_mat && want_n) MB_BWD_(true, true, true, true, true);
        else if (want_mat) MB_BWD_(true, false, true, true, true);
        else MB_BWD_(false, false, true, true, true);
    }
    else if (want_mat && want_n && want_env) MB_BWD_(true, true, true, false, true);
    else if (want_mat && want_n) MB_BWD_(true, true, false, false, true);
    else if (want_mat && want_env) MB_BWD(true, false, true, false);
    else if (want_mat) MB_BWD(true, false, false, false);
    else if (want_env) MB_BWD(false, false, true, false);
#undef MB_BWD_
#undef MB_BWD
    return mb200_check_launch();
}

}  // namespace

extern "C" {

int mb200_partial_stride(int filter) { return filter == MB200_FILTER_GAUSSIAN ? MB200_FILM_TAPS * 4 : 4; }

int mb200_fwd_partial_rows(const mb200_cfg* c, int* first_row) {
    int f, n; halo_rows(c, c->filter == MB200_FILTER_GAUSSIAN ? 2 : 0, &f, &n);
    if (first_row) *first_row = f; return n;
}
int mb200_bwd_wpart_rows(const mb200_cfg* c, int* first_row) {
    int f, n; halo_rows(c, c->filter == MB200_FILTER_GAUSSIAN ? 4 : 0, &f, &n);
    if (first_row) *first_row = f; return n;
}
int mb200_bwd_gadj_rows(const mb200_cfg* c, int* first_row) {
    int f, n; halo_rows(c, c->filter == MB200_FILTER_GAUSSIAN ? 2 : 0, &f, &n);
    if (first_row) *first_row = f; return n;
}

int mb200_shade_fwd(const mb200_cfg* c, const float* gpos, const float* gnrm, const float* a, const float* r, const float* m,
                    const float* n_opt, const float* env4, const float* hier, const mb200_hier_desc* d,
                    float* partials, void* stream) {
    RenderParams P; int rc = fill_params(c, gpos, gnrm, a, r, m, n_opt, env4, hier, d, P);
    if (rc) return rc;
    if (!partials) return MB200_EINVAL;
    P.prows = mb200_fwd_partial_rows(c, &P.prow0); P.partials = partials;
    const int npix = P.prows * P.W;
    cudaStream_t st = (cudaStream_t)stream;
    const bool ad = (c->flags & MB200_FLAG_AD_WEIGHTS) != 0;
    const size_t dyn = plan_env_staging(P, d, env_staging_budget(kStageBudgetFwd));
    const bool nmap = !P.use_mesh_normal && P.n_opt != nullptr;
#define MB_FWD_(F, A, L, NM) shade_fwd_kernel<F, A, false, L, NM><<<persistent_grid(shade_fwd_kernel<F, A, false, L, NM>, (npix + 32 / L - 1) / (32 / L), dyn), kThreads, dyn, st>>>(P)
#define MB_FWD(F, A, L) do { if (nmap) MB_FWD_(F, A, L, true); else MB_FWD_(F, A, L, false); } while (0)
    if (lanes_per_pixel(c->spp, false) == 16) {
        if (c->filter == MB200_FILTER_GAUSSIAN) { if (ad) MB_FWD(MB200_FILTER_GAUSSIAN, true, 16); else MB_FWD(MB200_FILTER_GAUSSIAN, false, 16); }
        else                                    { if (ad) MB_FWD(MB200_FILTER_BOX, true, 16);      else MB_FWD(MB200_FILTER_BOX, false, 16); }
    } else if (lanes_per_pixel(c->spp, false) == 8) {
        if (c->filter == MB200_FILTER_GAUSSIAN) { if (ad) MB_FWD(MB200_FILTER_GAUSSIAN, true, 8); else MB_FWD(MB200_FILTER_GAUSSIAN, false, 8); }
        else                                    { if (ad) MB_FWD(MB200_FILTER_BOX, true, 8);      else MB_FWD(MB200_FILTER_BOX, false, 8); }
    } else {
        if (c->filter == MB200_FILTER_GAUSSIAN) { if (ad) MB_FWD(MB200_FILTER_GAUSSIAN, true, 32); else MB_FWD(MB200_FILTER_GAUSSIAN, false, 32); }
        else                                    { if (ad) MB_FWD(MB200_FILTER_BOX, true, 32);      else MB_FWD(MB200_FILTER_BOX, false, 32); }
    }
#undef MB_FWD
#undef MB_FWD_
    return mb200_check_launch();
}

int mb200_trans_shade_fwd(const mb200_cfg* c, const mb200_trans* t, const float* gpos, const float* gnrm, const float* a, const float* r,
                          const float* m, const float* n_opt, const float* env4, const float* hier, const mb200_hier_desc* d,
                          float* partials, void* stream) {
    RenderParams P; int rc = fill_params(c, gpos, gnrm, a, r, m, n_opt, env4, hier, d, P);
    if (rc) return rc;
    if ((rc = fill_trans(t, P)) != MB200_OK) return rc;
    if (!partials) return MB200_EINVAL;
    if (c->flags & MB200_FLAG_AD_WEIGHTS) return MB200_EUNSUPPORTED;      // the reference never differentiates TransBSDF
    P.prows = mb200_fwd_partial_rows(c, &P.prow0); P.partials = partials;
    const int npix = P.prows * P.W;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t dyn = plan_env_staging(P, d, env_staging_budget(kStageBudgetFwd));
    if (c->filter == MB200_FILTER_GAUSSIAN) shade_fwd_kernel<MB200_FILTER_GAUSSIAN, false, true><<<persistent_grid(shade_fwd_kernel<MB200_FILTER_GAUSSIAN, false, true>, npix, dyn), kThreads, dyn, st>>>(P);
    else                                    shade_fwd_kernel<MB200_FILTER_BOX, false, true><<<persistent_grid(shade_fwd_kernel<MB200_FILTER_BOX, false, true>, npix, dyn), kThreads, dyn, st>>>(P);
    return mb200_check_launch();
}

int mb200_film_develop(const mb200_cfg* c, const float* partials, float* img, void* stream) {
    if (!c || !partials || !img) return MB200_EINVAL;
    int prow0; const int prows = mb200_fwd_partial_rows(c, &prow0);
    const int n = c->rows * c->W, tb = 256;
    cudaStream_t st = (cudaStream_t)stream;
    if (c->filter == MB200_FILTER_GAUSSIAN)
        film_develop_kernel<MB200_FILTER_GAUSSIAN><<<(n + tb - 1) / tb, tb, 0, st>>>(reinterpret_cast<const float4*>(partials), c->H, c->W, prow0, prows, c->row0, c->rows, img);
    else
        film_develop_kernel<MB200_FILTER_BOX><<<(n + tb - 1) / tb, tb, 0, st>>>(reinterpret_cast<const float4*>(partials), c->H, c->W, prow0, prows, c->row0, c->rows, img);
    return mb200_check_launch();
}

int mb200_film_weights(const mb200_cfg* c, float* wpart, void* stream) {
    if (!c) return MB200_EINVAL;
    if (c->filter != MB200_FILTER_GAUSSIAN) return MB200_OK;
    if (!wpart) return MB200_EINVAL;
    if ((double)c->H * (double)c->W * (double)c->spp >= 4294967296.0) return MB200_ERANGE;
    int wrow0; const int wrows = mb200_bwd_wpart_rows(c, &wrow0);
    const int npix = wrows * c->W, sms = mb200_sm_count();
    // one thread per pixel once that alone gives every SM two full CTAs; one warp per pixel below (small images / shards, high spp)
    static int px_mode = -1;
    if (px_mode < 0) { const char* e = getenv("MB200_FILM_WEIGHTS_PX"); px_mode = e ? atoi(e) : 1; }
    if (px_mode && npix >= sms * 2 * kThreads) {
        const int tiles = (npix + kThreads - 1) / kThreads, cap = sms * 8;
        film_weights_px_kernel<1><<<tiles < cap ? tiles : cap, kThreads, 0, (cudaStream_t)stream>>>(c->W, c->spp, c->seed, wrow0, wrows, wpart);
    } else
        film_weights_kernel<<<persistent_grid(film_weights_kernel, npix), kThreads, 0, (cudaStream_t)stream>>>(c->W, c->spp, c->seed, wrow0, wrows, wpart);
    return mb200_check_launch();
}

int mb200_film_adjoint(const mb200_cfg* c, const float* wpart, const float* grad_img_halo, float* gadj, void* stream) {
    if (!c || !grad_img_halo || !gadj) return MB200_EINVAL;
    int grow0; const int grows = mb200_bwd_gadj_rows(c, &grow0);
    const int n = grows * c->W, tb = 256;
    cudaStream_t st = (cudaStream_t)stream;
    if (c->filter == MB200_FILTER_GAUSSIAN) {
        if (!wpart) return MB200_EINVAL;
        int wrow0; const int wrows = mb200_bwd_wpart_rows(c, &wrow0);
        film_adjoint_kernel<MB200_FILTER_GAUSSIAN><<<(n + tb - 1) / tb, tb, 0, st>>>(wpart, c->H, c->W, wrow0, wrows, grow0, grows, 0.f, grad_img_halo, reinterpret_cast<float4*>(gadj));
    } else {
        film_adjoint_kernel<MB200_FILTER_BOX><<<(n + tb - 1) / tb, tb, 0, st>>>(nullptr, c->H, c->W, 0, 0, grow0, grows, 1.f / (float)c->spp, grad_img_halo, reinterpret_cast<float4*>(gadj));
    }
    return mb200_check_launch();
}

int mb200_shade_bwd(const mb200_cfg* c, const float* gpos, const float* gnrm, const float* a, const float* r, const float* m,
                    const float* n_opt, const float* env4, const float* hier, const mb200_hier_desc* d, const float* gadj,
                    float* g_a, float* g_r, float* g_m, float* g_n, float* g_env4, int n_env_slabs, void* stream) {
    RenderParams P; int rc = fill_params(c, gpos, gnrm, a, r, m, n_opt, env4, hier, d, P);
    if (rc) return rc;
    if (!gadj || (g_env4 && n_env_slabs < 1)) return MB200_EINVAL;
    P.env_slabs = g_env4 ? n_env_slabs : 1; P.env_slab_stride = (long long)d->res_x * d->res_y;
    P.prow0 = c->row0; P.prows = c->rows;
    P.gadj = reinterpret_cast<const float4*>(gadj); P.grows = mb200_bwd_gadj_rows(c, &P.grow0);
    P.g_a = g_a; P.g_r = g_r; P.g_m = g_m; P.g_n = g_n; P.g_env4 = reinterpret_cast<float4*>(g_env4);
    const bool want_n = g_n != nullptr && !c->use_mesh_normal;
    const bool want_mat = g_a || g_r || g_m || want_n;
    const bool want_env = g_env4 != nullptr;
    if (!want_mat && !want_env) return MB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t budget = env_staging_budget(kStageBudgetBwd);
    size_t dyn = plan_env_staging(P, d, budget);
    const size_t gbytes = ((size_t)3 * sizeof(float) * (size_t)P.env_slab_stride + 15) & ~(size_t)15;
    P.env_grad_smem = (want_env && env_scatter_aggregated() == 2 && dyn + gbytes <= budget) ? 1 : 0;   // block-privatised gradient map (small envmaps)
    if (P.env_grad_smem) dyn += gbytes;
    if (lanes_per_pixel(c->spp, true) == 8)
        return c->filter == MB200_FILTER_GAUSSIAN ? launch_bwd<MB200_FILTER_GAUSSIAN, 8>(P, want_mat, want_n, want_env, dyn, st)
                                                  : launch_bwd<MB200_FILTER_BOX, 8>(P, want_mat, want_n, want_env, dyn, st);
    return c->filter == MB200_FILTER_GAUSSIAN ? launch_bwd<MB200_FILTER_GAUSSIAN, 32>(P, want_mat, want_n, want_env, dyn, st)
                                              : launch_bwd<MB200_FILTER_BOX, 32>(P, want_mat, want_n, want_env, dyn, st);
}

int mb200_debug_sample_indices(const mb200_cfg* c, const float* gpos, const float* r, const float* hier,
                               const mb200_hier_desc* d, int32_t* out, void* stream) {
    if (!out) return MB200_EINVAL;
    RenderParams P; int rc = fill_params(c, gpos, gpos, r, r, r, r, hier, hier, d, P);   // only gpos/hier/cam are read
    if (rc) return rc;
    P.prow0 = c->row0; P.prows = c->rows;
    const long long n = (long long)P.prows * P.W * P.spp; const int tb = 256;
    sample_indices_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, (cudaStream_t)stream>>>(P, out);
    return mb200_check_launch();
}

int mb200_debug_sample_record(const mb200_cfg* c, const float* gpos, const float* gnrm, const float* a, const float* r, const float* m,
                              const float* n_opt, const float* env4, const float* hier, const mb200_hier_desc* d,
                              int32_t* out, float* out_radiance, void* stream) {
    if (!out) return MB200_EINVAL;
    RenderParams P; int rc = fill_params(c, gpos, gnrm, a, r, m, n_opt, env4, hier, d, P);
    if (rc) return rc;
    P.prow0 = c->row0; P.prows = c->rows;
    const long long n = (long long)P.prows * P.W * P.spp; const int tb = 128;
    sample_record_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, (cudaStream_t)stream>>>(P, out, out_radiance);
    return mb200_check_launch();
}

}  // extern "C"
