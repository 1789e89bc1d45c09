// mb200_posmlp_tc.cu — PosMLP forward (mymodels/mlps.py:211-251) on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// All five layers run as   D^T[feature, pixel] = W[feature, k] . X^T[k, pixel]   with the WEIGHTS as the M-side operand
// (two M = 128 halves) and a tile of 128 PIXELS as the N side, so one thread of the epilogue owns one output feature
// (= one TMEM lane) and walks the pixels (TMEM columns): its bias and its two coordinate weights are per-thread constants,
// and a warp writes 32 consecutive features of one pixel — coalesced 128-byte rows of the pre-activation cache.
//
// Precision: the reference runs these GEMMs in FP32 (TF32 off) and the parity bar is 1e-5, so every operand is split in
// two FP16 terms (x = x1 + x2, w = w1 + w2; 22 mantissa bits) and each K-step issues three kind::f16 MMAs
// (w1 x1 + w1 x2 + w2 x1) accumulating in FP32 in TMEM: activations are sin() outputs, the embedding's sin / cos features
// and colours are in [-1, 1] and weights are O(0.1) — all comfortably inside FP16 range.  What is NOT in that range, the
// two raw pixel coordinates (row, col in 0..3839) that enter lin0 and skip-connect into lin1 / lin3, have zero weight in
// the tensor-core images and are added in exact FP32 by the epilogue: z = acc + b + w_row * row + w_col * col.
//
// Roles (576 threads): warps 0-15 = epilogue (TMEM lane quarter = warp % 4, feature half = (warp / 4) % 2, pixel half =
// warp / 8: two threads per output feature, 64 pixels each — 4 warps per scheduler hide the LDS / MUFU / STG latencies),
// warp 16 = producer (one thread streams pre-split weight chunks global -> smem with cp.async.bulk + mbarrier tx-count),
// warp 17 = TMEM owner + MMA issuer (one elected thread issues tcgen05.mma, tcgen05.commit signals the barriers).
//
// Shared memory operand images are the no-swizzle K-major canonical layout of the UMMA smem descriptor
// (8 rows x 16 bytes core matrices): element (row r, k) at  (k/8)*LBO + (r/8)*128 + (r%8)*16 + (k%8)*2  bytes.
#include <cuda_fp16.h>
#include "mb200_posmlp.h"

namespace posmlp {
namespace {

constexpr int NPIX = 128;                 // pixels per tile (UMMA N)
constexpr int KCH = 32;                   // k per weight stage
constexpr int NCHUNK = HID / KCH;         // 8 chunks per layer
constexpr int NSTAGE = 2;
constexpr int NLAYER = 5;                 // lin0 .. lin4, all on the tensor cores
constexpr int W_BLK = 128 * 16;           // one k8 block of a 128-row weight half (bytes) = LBO of the A operand
constexpr int W_STAGE = 2 * 2 * (KCH / 8) * W_BLK;          // [term][half][k8][128 rows][8 halves] = 32768 B
constexpr int X_BLK = (NPIX + 1) * 16;    // one k8 block of the activation image, padded by one row: conflict-free epilogue stores
constexpr int X_SPLIT = (HID / 8) * X_BLK;                   // 66048 B per split term
constexpr int SM_X = 0;
constexpr int SM_W = SM_X + 2 * X_SPLIT;                     // 132096
constexpr int SM_PT = SM_W + NSTAGE * W_STAGE;               // 197632
constexpr int SM_OB = SM_PT + NPIX * 16 * 4;                 // 205824
constexpr int SM_BAR = SM_OB + OSTRIDE * NPIX * 4;           // 209920
constexpr int SM_TOTAL = SM_BAR + 64;
constexpr int NEPI = 512;                 // epilogue threads
constexpr int NTHREADS = NEPI + 64;
constexpr int TMEM_COLS = 256;            // two 128-feature halves x 128 pixel columns of FP32 accumulators

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, kind::f16, issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor): start >> 4 in [0,14), leading byte
// offset >> 4 in [16,30) (between the two 16-byte K chunks of one MMA), stride byte offset >> 4 in [32,46) (between 8-row
// groups), descriptor version 1 in [46,48), layout type 0 in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = F16 (0), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(NPIX >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---------------------------------------------------------------- weight pre-split
// wprep[L][chunk][term][half][k8][row 0..127][8 halves]  (L = lin0..lin4), w1 = fp16(w), w2 = fp16(w - w1).
// Tensor-core k index: lin0 uses k = embedding index (one K = 16 step); lin1 / lin3 see [h (h0) | embedding (d0)].
// Rows beyond the layer's outputs and the two coordinate columns (embedding index 0, 1: FP32 in the epilogue) are zero.
__global__ void __launch_bounds__(256) posmlp_prep_kernel(const Dims D, const float* __restrict__ params, __half* __restrict__ wprep) {
    const int total = NLAYER * HID * HID;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int L = i / (HID * HID), f = (i / HID) % HID, k = i % HID;
        const int n_out = (L == 0 || L == 2) ? D.h0 : (L == 4 ? D.n_out : HID);
        const int n_in = (L == 0) ? D.d0 : HID, ld = n_in;
        const int emb0 = (L == 0) ? 0 : ((L == 1 || L == 3) ? D.h0 : HID);        // first embedding column of this layer's input
        float w = 0.f;
        if (f < n_out && k < n_in && k != emb0 && k != emb0 + 1) w = params[D.oW[L] + f * ld + k];
        const __half w1 = __float2half_rn(w), w2 = __float2half_rn(w - __half2float(w1));
        const int c = k / KCH, k8 = (k % KCH) / 8, h = f >> 7, row = f & 127;
        const size_t base = ((size_t)L * NCHUNK + c) * (W_STAGE / 2);             // in halves
        const size_t off0 = (size_t)((0 * 2 + h) * (KCH / 8) + k8) * (W_BLK / 2) + row * 8 + (k & 7);
        const size_t off1 = (size_t)((1 * 2 + h) * (KCH / 8) + k8) * (W_BLK / 2) + row * 8 + (k & 7);
        wprep[base + off0] = w1; wprep[base + off1] = w2;
    }
}

// ---------------------------------------------------------------- forward
struct FwdArgs { Dims D; const float* params; const float* img; long long N; float* out; float* zc; float* oc; const __half* wprep; };

// sin(z) for |z| < 2^22: z = q pi + r (two-FMA Cody-Waite with the FP32 split of pi), odd degree-11 polynomial on
// |r| <= pi/2, sign from the parity of q.  <= 2 ulp like sinf, 14 instructions instead of ~26 (the epilogue is issue-bound).
__device__ __forceinline__ float sin_cw(float z) {
    const float magic = 12582912.f;                                   // 1.5 * 2^23: the add rounds z/pi to the nearest integer
    const float t = fmaf(z, 0.318309886183790672f, magic);
    const float q = t - magic;
    float r = fmaf(q, -3.14159274101257324f, z);
    r = fmaf(q, 8.74227765734758577e-8f, r);
    r = __uint_as_float(__float_as_uint(r) ^ (__float_as_uint(t) << 31));   // (-1)^q
    const float s = r * r;
    float p = -2.3898405032696246e-08f;
    p = fmaf(p, s, 2.752602995315101e-06f);
    p = fmaf(p, s, -0.0001984088303288445f);
    p = fmaf(p, s, 0.008333330973982811f);
    p = fmaf(p, s, -0.1666666716337204f);
    return fmaf(r * s, p, r);
}

__device__ __forceinline__ void store_x(uint8_t* sX, int f, int p, float x) {
    const __half x1 = __float2half_rn(x), x2 = __float2half_rn(x - __half2float(x1));
    const int off = (f >> 3) * X_BLK + p * 16 + (f & 7) * 2;
    *reinterpret_cast<__half*>(sX + off) = x1;
    *reinterpret_cast<__half*>(sX + X_SPLIT + off) = x2;
}

__global__ void __launch_bounds__(NTHREADS, 1) posmlp_fwd_tc_kernel(const __grid_constant__ FwdArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Dims& D = A.D;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sX = smem + SM_X;
    float* sPT = reinterpret_cast<float*>(smem + SM_PT);          // [pixel][16] FP32 embedding
    float* sOB = reinterpret_cast<float*>(smem + SM_OB);          // [OSTRIDE][NPIX]
    const uint32_t bar0 = smem_u32(smem + SM_BAR);
    const uint32_t bar_wfull = bar0, bar_wempty = bar0 + 8 * NSTAGE, bar_xready = bar0 + 16 * NSTAGE, bar_dfull = bar_xready + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 16 * NSTAGE + 16);

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
        mbar_init(bar_xready, NEPI); mbar_init(bar_dfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NEPI / 32 + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long ntiles = (A.N + NPIX - 1) / NPIX;

    if (warp < NEPI / 32) {
        // ===================================================== epilogue: thread <-> (feature f, pixel half)
        const int f = 128 * ((warp >> 2) & 1) + 32 * (warp & 3) + lane;
        const int pbeg = (warp >> 3) * (NPIX / 2), pend = pbeg + NPIX / 2;
        const uint32_t t_lane = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(((warp >> 2) & 1) * NPIX);
        uint32_t ph_d = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long n0 = tile * NPIX;
            // ---- positional embedding of the tile (mlps.py:190-209: raw (row, col), sin/cos of p and 2p, colour):
            //      FP32 copy in sPT[p][0..15] and, minus the two coordinates, as the FP16-split input image of lin0
            if (tid < NPIX) {
                const long long n = n0 + tid;
                float e[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) e[k] = 0.f;
                if (n < A.N) {
                    const float row = (float)(n / D.W), col = (float)(n % D.W);
                    e[0] = row; e[1] = col;
                    e[2] = sinf(row); e[3] = sinf(col); e[4] = cosf(row); e[5] = cosf(col);
                    e[6] = sinf(row * 2.f); e[7] = sinf(col * 2.f); e[8] = cosf(row * 2.f); e[9] = cosf(col * 2.f);
#pragma unroll
                    for (int c = 0; c < 6; ++c) if (c < D.n_color) e[10 + c] = A.img[n * D.n_color + c];
                }
#pragma unroll
                for (int k = 0; k < 16; k += 4) *reinterpret_cast<float4*>(sPT + tid * 16 + k) = make_float4(e[k], e[k + 1], e[k + 2], e[k + 3]);
#pragma unroll
                for (int k = 0; k < 16; ++k) store_x(sX, k, tid, k >= 2 ? e[k] : 0.f);
            }
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(bar_xready);
            named_bar_sync(1, NEPI);                                  // sPT visible to every epilogue thread
            // ---- lin0 .. lin3: z = acc + b + w_row row + w_col col ; cache z ; x = sin z -> input image of the next layer
            for (int L = 0; L < 4; ++L) {
                const int n_out = (L == 0 || L == 2) ? D.h0 : HID;
                const bool live = f < n_out;
                const int ld = (L == 0) ? D.d0 : HID, emb0 = (L == 0) ? 0 : D.h0;
                float w_row = 0.f, w_col = 0.f, bias = 0.f;
                if (live) {
                    bias = __ldg(A.params + D.ob[L] + f);
                    if (L != 2) { w_row = __ldg(A.params + D.oW[L] + f * ld + emb0); w_col = __ldg(A.params + D.oW[L] + f * ld + emb0 + 1); }
                }
                const int ke = f - n_out;                             // embedding index carried by this feature slot when !live
                mbar_wait(bar_dfull, ph_d); ph_d ^= 1; tc_fence_after();
                for (int p0 = pbeg; p0 < pend; p0 += 16) {
                    float acc[16];
                    tmem_ld16(t_lane + (uint32_t)p0, acc);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int p = p0 + i;
                        const float2 rc = *reinterpret_cast<const float2*>(sPT + p * 16);
                        const float z = fmaf(w_row, rc.x, fmaf(w_col, rc.y, acc[i] + bias));
                        if (A.zc && n0 + p < A.N) A.zc[(n0 + p) * ZSTRIDE + L * HID + f] = live ? z : 0.f;
                        // features >= n_out of the next layer's input are the concatenated embedding (coordinates: FP32 side term)
                        const float x = live ? sin_cw(z) : (ke >= 2 ? sPT[p * 16 + ke] : 0.f);
                        store_x(sX, f, p, x);
                    }
                }
                tc_fence_before();
                fence_proxy_async();                                  // generic-proxy smem writes -> visible to the tensor core (async proxy)
                mbar_arrive(bar_xready);
            }
            // ---- lin4 accumulators (half 0, lanes 0..n_out-1) -> sOB, then the output activation on all epilogue threads
            mbar_wait(bar_dfull, ph_d); ph_d ^= 1; tc_fence_after();
            if (warp == 0) {
                for (int p0 = 0; p0 < NPIX; p0 += 16) {
                    float acc[16];
                    tmem_ld16(tmem_base + (uint32_t)p0, acc);
                    if (lane < OSTRIDE) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) sOB[lane * NPIX + p0 + i] = acc[i];
                    }
                }
            }
            tc_fence_before();
            named_bar_sync(1, NEPI);
            for (int i = tid; i < D.n_out * NPIX; i += NEPI) {
                const int o = i / NPIX, p = i % NPIX; const long long n = n0 + p;
                if (n < A.N) {
                    const float v = sOB[o * NPIX + p] + __ldg(A.params + D.ob[4] + o);
                    if (A.oc) A.oc[n * OSTRIDE + o] = v;
                    float y;
                    if (D.otype == 0) y = v > 20.f ? v : log1pf(expf(v));                                   // nn.Softplus (threshold 20)
                    else y = fminf(fmaxf(1.3f * tanhf(v) + A.img[n * D.n_color + o], 0.f), 1.f);           // straight-through clamp: value
                    A.out[n * D.n_out + o] = y;
                }
            }
            named_bar_sync(1, NEPI);                                  // sOB / sPT / the image are rewritten by the next tile
        }
    } else if (warp == NEPI / 32) {
        // ===================================================== producer: weight chunks global -> smem
        if (lane == 0) {
            uint32_t it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int L = 0; L < NLAYER; ++L) {
                    const int nch = (L == 0) ? 1 : NCHUNK;            // lin0: K = 16, one chunk
                    for (int c = 0; c < nch; ++c, ++it) {
                        const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
                        mbar_wait(bar_wempty + 8 * s, ph ^ 1);
                        mbar_expect_tx(bar_wfull + 8 * s, W_STAGE);
                        bulk_g2s(smem_u32(smem + SM_W + s * W_STAGE), reinterpret_cast<const uint8_t*>(A.wprep) + ((size_t)L * NCHUNK + c) * W_STAGE,
                                 W_STAGE, bar_wfull + 8 * s);
                    }
                }
            }
        }
    } else {
        // ===================================================== MMA issuer (one lane of the last warp)
        if (lane == 0) {
            uint32_t it = 0, ph_x = 0;
            const uint32_t xa = smem_u32(sX), wa = smem_u32(smem + SM_W);
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int L = 0; L < NLAYER; ++L) {
                    mbar_wait(bar_xready, ph_x); ph_x ^= 1;
                    tc_fence_after();
                    const int nhalf = (L == NLAYER - 1) ? 1 : 2;      // lin4: n_out <= 8 rows live in half 0
                    const int nch = (L == 0) ? 1 : NCHUNK, nj = (L == 0) ? 1 : KCH / 16;
                    for (int c = 0; c < nch; ++c, ++it) {
                        const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
                        mbar_wait(bar_wfull + 8 * s, ph);
                        tc_fence_after();
                        const uint32_t wst = wa + s * W_STAGE;
                        for (int h = 0; h < nhalf; ++h) {
                            for (int j = 0; j < nj; ++j) {
                                const uint32_t kb = (uint32_t)(c * (KCH / 8) + 2 * j);          // first k8 block of this K = 16 step
                                const uint64_t b1 = umma_desc(xa + kb * X_BLK, X_BLK, 128), b2 = umma_desc(xa + X_SPLIT + kb * X_BLK, X_BLK, 128);
                                const uint64_t a1 = umma_desc(wst + (uint32_t)((0 * 2 + h) * (KCH / 8) + 2 * j) * W_BLK, W_BLK, 128);
                                const uint64_t a2 = umma_desc(wst + (uint32_t)((1 * 2 + h) * (KCH / 8) + 2 * j) * W_BLK, W_BLK, 128);
                                const uint32_t d = tmem_base + (uint32_t)(h * NPIX);
                                umma_f16(d, a1, b1, kIdesc, (c | j) != 0);        // w1 x1
                                umma_f16(d, a1, b2, kIdesc, 1);                   // w1 x2
                                umma_f16(d, a2, b1, kIdesc, 1);                   // w2 x1
                            }
                        }
                        tc_commit(bar_wempty + 8 * s);                // stage free once these MMAs have read it
                    }
                    tc_commit(bar_dfull);                             // accumulators complete, activation image free
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NEPI / 32 + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace

size_t tc_workspace_bytes() { return (size_t)NLAYER * NCHUNK * W_STAGE; }

int tc_forward(const Dims& D, const float* params, const float* img, long long N, float* out, float* zc, float* oc, void* wprep, cudaStream_t st) {
    if (((uintptr_t)wprep & 15) != 0) return MB200_EINVAL;
    posmlp_prep_kernel<<<160, 256, 0, st>>>(D, params, reinterpret_cast<__half*>(wprep));
    int rc = mb200_check_launch();
    if (rc) return rc;
    rc = mb200_check(cudaFuncSetAttribute(posmlp_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    if (rc) return rc;
    FwdArgs A; A.D = D; A.params = params; A.img = img; A.N = N; A.out = out; A.zc = zc; A.oc = oc; A.wprep = reinterpret_cast<const __half*>(wprep);
    const long long ntiles = (N + NPIX - 1) / NPIX;
    const int grid = (int)(ntiles < (long long)mb200_sm_count() ? ntiles : (long long)mb200_sm_count());
    posmlp_fwd_tc_kernel<<<grid, NTHREADS, SM_TOTAL, st>>>(A);
    return mb200_check_launch();
}

}  // namespace posmlp
