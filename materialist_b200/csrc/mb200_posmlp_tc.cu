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
// Roles (608 threads): warps 0-15 = epilogue (TMEM lane quarter = warp % 4, feature half = (warp / 4) % 2, pixel half =
// warp / 8: two threads per output feature, 64 pixels each — 4 warps per scheduler hide the LDS / MUFU / STG latencies),
// warp 16 = producer (one thread streams pre-split weight chunks global -> smem with cp.async.bulk + mbarrier tx-count),
// warp 17 = TMEM owner + MMA issuer (one elected thread issues tcgen05.mma, tcgen05.commit signals the barriers),
// warp 18 = image store (bulk-copies the FP16-split activation / gradient images smem -> global for the weight-gradient GEMM).
//
// Backward = two kernels.  posmlp_bwd_data_tc_kernel walks the layers in reverse with the same skeleton (A = W^T images,
// B = the dL/dz image, all gradients carried scaled by a power of two so that they sit in FP16 range); bias gradients, the
// lin0 / lin4 weight gradients and the coordinate columns of lin1 / lin3 are exact-FP32 per-thread register accumulators.
// posmlp_wgrad_tc_kernel then forms gW_l = G_l X_l^T for the three 256x256 layers: M = output feature, N = input feature,
// K = PIXELS, both operands MN-major straight from the stored images, 256x256 FP32 accumulators = all 512 TMEM columns.
//
// Shared memory operand images are the no-swizzle K-major canonical layout of the UMMA smem descriptor
// (8 rows x 16 bytes core matrices): element (row r, k) at  (k/8)*LBO + (r/8)*128 + (r%8)*16 + (k%8)*2  bytes.
#include <cuda_fp16.h>
#include "mb200_posmlp.h"

namespace posmlp {
namespace {

constexpr int NPIX = 128;                 // pixels per tile (UMMA N)
#ifndef MB200_POSMLP_KCH
#define MB200_POSMLP_KCH 32
#endif
#ifndef MB200_POSMLP_NSTAGE
#define MB200_POSMLP_NSTAGE 2
#endif
// (measured, profiles/r5s: 5 stages of 16 KB — all the shared memory that is left — instead of 2 of 32 KB: forward 0.740 -> 0.799 ms;
// the weight stream is not what the kernel waits for, twice as many chunk barriers cost more than the deeper ring gains)
constexpr int KCH = MB200_POSMLP_KCH;     // k per weight stage
constexpr int NCHUNK = HID / KCH;         // chunks per layer
constexpr int NSTAGE = MB200_POSMLP_NSTAGE;
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
constexpr int SM_TOTAL = SM_BAR + 256;
static_assert(SM_TOTAL <= 232448, "shared memory per CTA");
constexpr int NEPI = 512;                 // epilogue threads
constexpr int NTHREADS = NEPI + 96;       // + producer warp, MMA warp, image-store warp
constexpr int WARP_PROD = NEPI / 32, WARP_MMA = NEPI / 32 + 1, WARP_STORE = NEPI / 32 + 2;
// Operand images kept for the weight-gradient GEMM (mb200_posmlp_wgrad_tc_kernel): per (layer, tile, pixel quarter)
// one contiguous 32 KB piece [split][k8][32 px][8 halves] — the MN-major canonical layout with K = pixel.
constexpr int IMG_PIECE = 2 * (HID / 8) * 32 * 16;         // 32768 B
constexpr int IMG_TILE = 4 * IMG_PIECE;                     // 131072 B per (layer, tile)
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
// same, with the descriptors as (low word, high word): the MMA issuer keeps the constant high words in registers and only ADDS
// 16-byte-unit offsets to the low words (start address field) inside its loops — rebuilding four 64-bit descriptors per K step
// with shifts and masks cost the single issuing thread ~12 k dependent instructions per tile (profiles/r5r), a serial bottleneck
// that the epilogue warps ended up waiting for
__device__ __forceinline__ void umma_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
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
constexpr int NSUB = 2, SPIX = NPIX / NSUB;        // sub-tiles of 64 pixels: the MMAs of one run under the epilogue of the other
constexpr uint32_t kIdescSub = (1u << 4) | ((uint32_t)(SPIX >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// smem -> global bulk store (async proxy); completion of the smem READS is awaited with wait_group.read
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// MN-major, SWIZZLE_NONE descriptor: sbo = stride between groups of 8 M/N elements, lbo = stride between groups of 8 K elements
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---------------------------------------------------------------- weight pre-split
// wprep[L][chunk][term][half][k8][row 0..127][8 halves]  (L = lin0..lin4), w1 = fp16(w), w2 = fp16(w - w1).
// Tensor-core k index: lin0 uses k = embedding index (one K = 16 step); lin1 / lin3 see [h (h0) | embedding (d0)].
// Rows beyond the layer's outputs and the two coordinate columns (embedding index 0, 1: FP32 in the epilogue) are zero.
// TRANSPOSED = true writes the images of the data-gradient pass instead: slot t = 0,1,2 <-> W3^T, W2^T, W1^T with
// rows = input feature k, reduction index = output feature c.
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256) posmlp_prep_kernel(const Dims D, const float* __restrict__ params, __half* __restrict__ wprep) {
    const int nl = TRANSPOSED ? 3 : NLAYER;
    const int total = nl * HID * HID;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int slot = i / (HID * HID), r = (i / HID) % HID, k = i % HID;        // image row r, reduction index k
        float w = 0.f;
        if (!TRANSPOSED) {
            const int L = slot;
            const int n_out = (L == 0 || L == 2) ? D.h0 : (L == 4 ? D.n_out : HID);
            const int n_in = (L == 0) ? D.d0 : HID;
            const int emb0 = (L == 0) ? 0 : ((L == 1 || L == 3) ? D.h0 : HID);    // first embedding column of this layer's input
            if (r < n_out && k < n_in && k != emb0 && k != emb0 + 1) w = params[D.oW[L] + r * n_in + k];
        } else {
            const int L = 3 - slot;                                               // lin3, lin2, lin1
            const int n_out = (L == 2) ? D.h0 : HID;
            if (k < n_out) w = params[D.oW[L] + k * HID + r];                     // A[r = input feature][k = output feature] = W_L[k][r]
        }
        const __half w1 = __float2half_rn(w), w2 = __float2half_rn(w - __half2float(w1));
        const int c = k / KCH, k8 = (k % KCH) / 8, h = r >> 7, row = r & 127;
        const size_t base = ((size_t)slot * NCHUNK + c) * (W_STAGE / 2);          // in halves
        const size_t off0 = (size_t)((0 * 2 + h) * (KCH / 8) + k8) * (W_BLK / 2) + row * 8 + (k & 7);
        const size_t off1 = (size_t)((1 * 2 + h) * (KCH / 8) + k8) * (W_BLK / 2) + row * 8 + (k & 7);
        wprep[base + off0] = w1; wprep[base + off1] = w2;
    }
}

// ---------------------------------------------------------------- math helpers
// sin(z) for |z| < 2^22: z = q pi + r (two-FMA Cody-Waite with the FP32 split of pi), odd degree-11 polynomial on
// |r| <= pi/2, sign from the parity of q.  <= 2 ulp like sinf, 14 instructions instead of ~26 (the epilogue is issue-bound).
__device__ __forceinline__ float reduce_pi(float z, uint32_t& sign) {
    const float magic = 12582912.f;                                   // 1.5 * 2^23: the add rounds z/pi to the nearest integer
    const float t = fmaf(z, 0.318309886183790672f, magic);
    const float q = t - magic;
    float r = fmaf(q, -3.14159274101257324f, z);
    r = fmaf(q, 8.74227765734758577e-8f, r);
    sign = __float_as_uint(t) << 31;                                  // parity of q -> sign bit
    return r;
}
__device__ __forceinline__ float sin_poly(float r) {
    const float s = r * r;
    float p = -2.3898405032696246e-08f;
    p = fmaf(p, s, 2.752602995315101e-06f);
    p = fmaf(p, s, -0.0001984088303288445f);
    p = fmaf(p, s, 0.008333330973982811f);
    p = fmaf(p, s, -0.1666666716337204f);
    return fmaf(r * s, p, r);
}
__device__ __forceinline__ float cos_poly(float r) {
    const float s = r * r;
    float p = 1.9910568749281765e-09f;
    p = fmaf(p, s, -2.752503291958419e-07f);
    p = fmaf(p, s, 2.4801054678391665e-05f);
    p = fmaf(p, s, -0.0013888884568586946f);
    p = fmaf(p, s, 0.0416666679084301f);
    p = fmaf(p, s, -0.5f);
    return fmaf(s, p, 1.f);
}
__device__ __forceinline__ float sin_cw(float z) {
    uint32_t sg; const float r = reduce_pi(z, sg);
    return sin_poly(__uint_as_float(__float_as_uint(r) ^ sg));
}
__device__ __forceinline__ void sincos_cw(float z, float& sn, float& cs) {
    uint32_t sg; const float r = reduce_pi(z, sg);
    sn = __uint_as_float(__float_as_uint(sin_poly(r)) ^ sg);
    cs = __uint_as_float(__float_as_uint(cos_poly(r)) ^ sg);
}

__device__ __forceinline__ void store_x(uint8_t* sX, int f, int p, float x) {
    const __half x1 = __float2half_rn(x), x2 = __float2half_rn(x - __half2float(x1));
    const int off = (f >> 3) * X_BLK + p * 16 + (f & 7) * 2;
    *reinterpret_cast<__half*>(sX + off) = x1;
    *reinterpret_cast<__half*>(sX + X_SPLIT + off) = x2;
}

// positional embedding of pixel n (mlps.py:190-209: raw (row, col), sin/cos of p and 2p, colour), zero padded to 16
__device__ __forceinline__ void embed_pixel(const Dims& D, const float* __restrict__ img, long long n, long long N, float (&e)[16]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) e[k] = 0.f;
    if (n < N) {
        const float row = (float)(n / D.W + D.row0), col = (float)(n % D.W);
        e[0] = row; e[1] = col;
        e[2] = sinf(row); e[3] = sinf(col); e[4] = cosf(row); e[5] = cosf(col);
        e[6] = sinf(row * 2.f); e[7] = sinf(col * 2.f); e[8] = cosf(row * 2.f); e[9] = cosf(col * 2.f);
#pragma unroll
        for (int c = 0; c < 6; ++c) if (c < D.n_color) e[10 + c] = img[n * D.n_color + c];
    }
}

// ---------------------------------------------------------------- pipeline pieces shared by the forward and the data-gradient kernels
// Two 64-pixel SUB-TILES per 128-pixel tile, each with its own barriers: while the 16 epilogue warps turn the accumulators of
// sub-tile s into the next layer's operand rows, the MMA issuer is already working on sub-tile 1 - s (its operand rows and its TMEM
// columns are disjoint).  Before, MMA and epilogue of a tile strictly alternated: the tensor pipe was busy 22 % of the time and the
// epilogue warps waited for it the rest of that (ncu: profiles/r5q).  Price: the weight chunks of a layer are streamed from L2 once
// per sub-tile (2.1 MB per tile instead of 1.05 MB).
struct Pipe {
    uint32_t wfull, wempty, xready0, dfull0, sdone0; uint32_t* tmem_slot;
    __device__ __forceinline__ uint32_t xready(int u) const { return xready0 + 8u * (uint32_t)u; }
    __device__ __forceinline__ uint32_t dfull(int u) const { return dfull0 + 8u * (uint32_t)u; }
    __device__ __forceinline__ uint32_t sdone(int u) const { return sdone0 + 8u * (uint32_t)u; }
};

__device__ __forceinline__ Pipe pipe_setup(uint8_t* smem, int tid, int warp) {
    Pipe P;
    const uint32_t bar0 = smem_u32(smem + SM_BAR);
    P.wfull = bar0; P.wempty = bar0 + 8 * NSTAGE;
    const uint32_t b1 = bar0 + 16 * NSTAGE;
    P.xready0 = b1; P.dfull0 = b1 + 8 * NSUB; P.sdone0 = b1 + 16 * NSUB;
    P.tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 16 * NSTAGE + 24 * NSUB);
    static_assert(16 * NSTAGE + 24 * NSUB + 4 <= 256, "barrier area");
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(P.wfull + 8 * s, 1); mbar_init(P.wempty + 8 * s, 1); }
        for (int u = 0; u < NSUB; ++u) { mbar_init(P.xready(u), NEPI); mbar_init(P.dfull(u), 1); mbar_init(P.sdone(u), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(P.tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return P;
}
__device__ __forceinline__ void pipe_teardown(int warp, uint32_t tmem_base) {
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}
// phases per tile and their shapes.  FWD: lin0 (K = 16), lin1, lin2, lin3, lin4 (one half).  BWD: W3^T, W2^T, W1^T.
template <bool FWD> __device__ __forceinline__ int n_phase() { return FWD ? NLAYER : 3; }
template <bool FWD> __device__ __forceinline__ int phase_chunks(int ph) { return (FWD && ph == 0) ? 1 : NCHUNK; }

// producer: one thread streams the weight chunks of every phase of every tile through the stage ring
template <bool FWD>
__device__ __forceinline__ void producer_loop(const Pipe& P, uint8_t* smem, const __half* wimg, long long ntiles) {
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int ph = 0; ph < n_phase<FWD>(); ++ph)
            for (int u = 0; u < NSUB; ++u) {
                const int nch = phase_chunks<FWD>(ph);
                for (int c = 0; c < nch; ++c, ++it) {
                    const uint32_t s = it % NSTAGE, par = (it / NSTAGE) & 1;
                    mbar_wait(P.wempty + 8 * s, par ^ 1);
                    mbar_expect_tx(P.wfull + 8 * s, W_STAGE);
                    bulk_g2s(smem_u32(smem + SM_W + s * W_STAGE), reinterpret_cast<const uint8_t*>(wimg) + ((size_t)ph * NCHUNK + c) * W_STAGE,
                             W_STAGE, P.wfull + 8 * s);
                }
            }
}
// MMA issuer: one thread; per phase waits for the operand image, then 3 split-term MMAs per K = 16 step and half
template <bool FWD>
__device__ __forceinline__ void mma_loop(const Pipe& P, uint8_t* smem, uint32_t tmem_base, long long ntiles) {
    uint32_t it = 0, ph_x = 0;
    const uint32_t xa = smem_u32(smem + SM_X), wa = smem_u32(smem + SM_W);
    // descriptor words (umma_desc): low = start >> 4 | (LBO >> 4) << 16, high = SBO >> 4 | version 1 << 14
    const uint32_t b_hi = (128u >> 4) | (1u << 14), a_hi = b_hi;
    const uint32_t b_lo0 = ((xa >> 4) & 0x3FFFu) | (((uint32_t)X_BLK >> 4) << 16);          // split term 1, k8 block 0, pixel row 0
    const uint32_t a_lo0 = ((wa >> 4) & 0x3FFFu) | (((uint32_t)W_BLK >> 4) << 16);          // stage 0, term 1, half 0, k8 block 0
    constexpr uint32_t kXBlk16 = X_BLK >> 4, kXSplit16 = X_SPLIT >> 4, kWBlk16 = W_BLK >> 4, kWStage16 = W_STAGE >> 4;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int ph = 0; ph < n_phase<FWD>(); ++ph, ph_x ^= 1)
            for (int u = 0; u < NSUB; ++u) {
                mbar_wait(P.xready(u), ph_x);
                tc_fence_after();
                const int nhalf = (FWD && ph == NLAYER - 1) ? 1 : 2;      // lin4: n_out <= 8 rows live in half 0
                const int nch = phase_chunks<FWD>(ph), nj = (FWD && ph == 0) ? 1 : KCH / 16;
                const uint32_t b_u = b_lo0 + (uint32_t)(u * SPIX);        // operand rows (pixels) of this sub-tile: 16 bytes per row
                for (int c = 0; c < nch; ++c, ++it) {
                    const uint32_t s = it % NSTAGE, par = (it / NSTAGE) & 1;
                    mbar_wait(P.wfull + 8 * s, par);
                    tc_fence_after();
                    const uint32_t a_s = a_lo0 + s * kWStage16;
                    for (int h = 0; h < nhalf; ++h)
                        for (int j = 0; j < nj; ++j) {
                            const uint32_t kb = (uint32_t)(c * (KCH / 8) + 2 * j);          // first k8 block of this K = 16 step
                            const uint32_t b1 = b_u + kb * kXBlk16, b2 = b1 + kXSplit16;
                            const uint32_t a1 = a_s + (uint32_t)((0 * 2 + h) * (KCH / 8) + 2 * j) * kWBlk16;
                            const uint32_t a2 = a_s + (uint32_t)((1 * 2 + h) * (KCH / 8) + 2 * j) * kWBlk16;
                            const uint32_t d = tmem_base + (uint32_t)(h * NPIX + u * SPIX);
                            if (FWD) {
                                umma_f16_w(d, a1, a_hi, b1, b_hi, kIdescSub, (c | j) != 0);        // w1 x1
                                umma_f16_w(d, a1, a_hi, b2, b_hi, kIdescSub, 1);                   // w1 x2
                                umma_f16_w(d, a2, a_hi, b1, b_hi, kIdescSub, 1);                   // w2 x1
                            } else {
                                // (the data-gradient kernel keeps the descriptors as 64-bit values: with the split form its register
                                // allocation tips into 1.7 KB of spills in the epilogue branch — measured 1.90 -> 2.31 ms fwd + bwd)
                                const uint64_t A1 = ((uint64_t)a_hi << 32) | a1, A2 = ((uint64_t)a_hi << 32) | a2;
                                const uint64_t B1 = ((uint64_t)b_hi << 32) | b1, B2 = ((uint64_t)b_hi << 32) | b2;
                                umma_f16(d, A1, B1, kIdescSub, (c | j) != 0);
                                umma_f16(d, A1, B2, kIdescSub, 1);
                                umma_f16(d, A2, B1, kIdescSub, 1);
                            }
                        }
                    tc_commit(P.wempty + 8 * s);                  // stage free once these MMAs have read it
                }
                tc_commit(P.dfull(u));                            // accumulators of this sub-tile complete, its operand rows free
            }
}
// image-store warp: after every operand image is complete, optionally bulk-copy it to global as 4 pixel-quarter pieces
// ([split][k8][32 px][8 halves], 32 KB each; lane <-> k8 block), then release the image for the next writer.
// slot_of(ph) = image slot (0..2) to store in phase ph, or -1.
template <bool FWD>
__device__ __forceinline__ void store_loop(const Pipe& P, uint8_t* smem, uint8_t* imgs, long long ntiles, int lane) {
    uint32_t ph_x = 0;
    const uint32_t xa = smem_u32(smem + SM_X);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int ph = 0; ph < n_phase<FWD>(); ++ph, ph_x ^= 1)
            for (int u = 0; u < NSUB; ++u) {
                mbar_wait(P.xready(u), ph_x);
                // FWD: phases 1,2,3 hold X_1, X_2, X_3 (inputs of lin1..lin3) -> slots 0,1,2.  BWD: phases 0,1,2 hold G_3, G_2, G_1 -> slots 2,1,0
                const int slot = FWD ? ((ph >= 1 && ph <= 3) ? ph - 1 : -1) : 2 - ph;
                if (imgs && slot >= 0) {
                    uint8_t* dst = imgs + ((size_t)slot * ntiles + tile) * IMG_TILE;
#pragma unroll
                    for (int q = u * (4 / NSUB); q < (u + 1) * (4 / NSUB); ++q)          // the pixel quarters of this sub-tile
#pragma unroll
                        for (int sp = 0; sp < 2; ++sp)
                            bulk_s2g(dst + (size_t)q * IMG_PIECE + sp * (IMG_PIECE / 2) + lane * 512, xa + sp * X_SPLIT + lane * X_BLK + q * 512, 512);
                    bulk_commit_wait_read();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(P.sdone(u));
            }
}

// ---------------------------------------------------------------- forward
struct FwdArgs { Dims D; const float* params; const float* img; long long N; float* out; float* zc; float* oc; const __half* wprep; uint8_t* ximg; };

__global__ void __launch_bounds__(NTHREADS, 1) posmlp_fwd_tc_kernel(const __grid_constant__ FwdArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Dims& D = A.D;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sX = smem + SM_X;
    float* sPT = reinterpret_cast<float*>(smem + SM_PT);          // [pixel][16] FP32 embedding
    float* sOB = reinterpret_cast<float*>(smem + SM_OB);          // [OSTRIDE][NPIX]
    const Pipe P = pipe_setup(smem, tid, warp);
    const uint32_t tmem_base = *P.tmem_slot;
    const long long ntiles = (A.N + NPIX - 1) / NPIX;

    if (warp < NEPI / 32) {
        // ===================================================== epilogue: thread <-> (feature f, pixel quarter of each sub-tile)
        const int f = 128 * ((warp >> 2) & 1) + 32 * (warp & 3) + lane;
        const int pq = (warp >> 3) * (SPIX / 2);                     // first pixel of this thread's 32 inside a sub-tile
        const uint32_t t_lane = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(((warp >> 2) & 1) * NPIX);
        uint32_t ph_d = 0, ph_s = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long n0 = tile * NPIX;
            // ---- embedding: FP32 copy in sPT[p][0..15] and, minus the two coordinates, the FP16-split input image of lin0
            if (tid < NPIX) {
                float e[16];
                embed_pixel(D, A.img, n0 + tid, A.N, e);
#pragma unroll
                for (int k = 0; k < 16; k += 4) *reinterpret_cast<float4*>(sPT + tid * 16 + k) = make_float4(e[k], e[k + 1], e[k + 2], e[k + 3]);
#pragma unroll
                for (int k = 0; k < 16; ++k) store_x(sX, k, tid, k >= 2 ? e[k] : 0.f);
            }
            tc_fence_before();
            fence_proxy_async();
            named_bar_sync(1, NEPI);                                  // sPT and the lin0 image complete (written by warps 0..3 for both sub-tiles)
#pragma unroll
            for (int u = 0; u < NSUB; ++u) mbar_arrive(P.xready(u));
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
                for (int u = 0; u < NSUB; ++u) {
                    mbar_wait(P.dfull(u), ph_d); tc_fence_after();
                    mbar_wait(P.sdone(u), ph_s);                      // the previous image rows of this sub-tile have been copied out
                    const int pbeg = u * SPIX + pq;
                    for (int p0 = pbeg; p0 < pbeg + SPIX / 2; p0 += 16) {
                        float acc[16];
                        tmem_ld16(t_lane + (uint32_t)p0, acc);
                        // pre-activation cache: one pointer per 16-pixel block, compile-time offsets inside it (the per-element 64-bit
                        // index arithmetic was 14 % of the kernel's instructions, profiles/r5r)
                        float* const zp = A.zc ? A.zc + (n0 + p0) * ZSTRIDE + L * HID + f : nullptr;
                        const int nval = zp ? (int)min((long long)16, A.N - (n0 + p0)) : 0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int p = p0 + i;
                            const float2 rc = *reinterpret_cast<const float2*>(sPT + p * 16);
                            const float z = fmaf(w_row, rc.x, fmaf(w_col, rc.y, acc[i] + bias));
                            if (i < nval) zp[i * ZSTRIDE] = live ? z : 0.f;
                            // features >= n_out of the next layer's input are the concatenated embedding (coordinates: FP32 side term)
                            const float x = live ? sin_cw(z) : (ke >= 2 ? sPT[p * 16 + ke] : 0.f);
                            store_x(sX, f, p, x);
                        }
                    }
                    tc_fence_before();
                    fence_proxy_async();                              // generic-proxy smem writes -> visible to the async proxy (UMMA, bulk store)
                    mbar_arrive(P.xready(u));
                }
                ph_d ^= 1; ph_s ^= 1;
            }
            // ---- lin4 accumulators (half 0, lanes 0..n_out-1) -> sOB, then the output activation on all epilogue threads
#pragma unroll
            for (int u = 0; u < NSUB; ++u) { mbar_wait(P.dfull(u), ph_d); }
            ph_d ^= 1;
            tc_fence_after();
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
#pragma unroll
            for (int u = 0; u < NSUB; ++u) { mbar_wait(P.sdone(u), ph_s); }
            ph_s ^= 1;
            named_bar_sync(1, NEPI);                                  // sOB / sPT / the image are rewritten by the next tile
        }
    } else if (warp == WARP_PROD) {
        if (lane == 0) producer_loop<true>(P, smem, A.wprep, ntiles);
    } else if (warp == WARP_MMA) {
        if (lane == 0) mma_loop<true>(P, smem, tmem_base, ntiles);
    } else {
        store_loop<true>(P, smem, A.ximg, ntiles, lane);
    }
    pipe_teardown(warp, tmem_base);
}

// ---------------------------------------------------------------- backward: data gradients (+ all the small exact-FP32 gradients)
struct BwdArgs {
    Dims D; const float* params; const float* img; long long N; const float* zc; const float* oc; const float* g_out;
    float* g_params; const __half* wprepT; uint8_t* gimg; const float* gmax;
};
// power-of-two scale that puts max|g_out| in [8, 16): gradients are carried scaled through the whole backward pass
__device__ __forceinline__ void grad_scale(float gmax, float& s, float& inv_s) {
    const uint32_t e = (__float_as_uint(gmax) >> 23) & 0xFFu;
    if (e < 3u || e > 250u) { s = 1.f; inv_s = 1.f; return; }
    s = __uint_as_float((257u - e) << 23); inv_s = __uint_as_float((e - 3u) << 23);
}
__device__ __forceinline__ float cos_cw(float z) {
    uint32_t sg; const float r = reduce_pi(z, sg);
    return __uint_as_float(__float_as_uint(cos_poly(r)) ^ sg);
}

__global__ void __launch_bounds__(NTHREADS, 1) posmlp_bwd_data_tc_kernel(const __grid_constant__ BwdArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Dims& D = A.D;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sX = smem + SM_X;
    float* sPT = reinterpret_cast<float*>(smem + SM_PT);          // [pixel][16] FP32 embedding
    float* sG4 = reinterpret_cast<float*>(smem + SM_OB);          // [pixel][OSTRIDE] scaled dL/d(lin4 output)
    const Pipe P = pipe_setup(smem, tid, warp);
    const uint32_t tmem_base = *P.tmem_slot;
    const long long ntiles = (A.N + NPIX - 1) / NPIX;
    float gs, inv_gs; grad_scale(__ldg(A.gmax), gs, inv_gs);

    if (warp < NEPI / 32) {
        const int f = 128 * ((warp >> 2) & 1) + 32 * (warp & 3) + lane;
        const int pq = (warp >> 3) * (SPIX / 2);                     // first pixel of this thread's 32 inside a sub-tile
        const uint32_t t_lane = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(((warp >> 2) & 1) * NPIX);
        uint32_t ph_d = 0, ph_s = 0;
        float w4[OSTRIDE];
#pragma unroll
        for (int o = 0; o < OSTRIDE; ++o) w4[o] = o < D.n_out ? __ldg(A.params + D.oW[4] + o * HID + f) : 0.f;
        // per-thread accumulators (scaled by gs): bias gradients of lin0..lin3 at feature f, gW4[:, f], the two coordinate
        // columns of gW1 / gW3 at row f, gW0[f, :], and (threads < n_out, pixel half 0) the lin4 bias gradient
        float gb[4] = {0.f, 0.f, 0.f, 0.f}, gW4a[OSTRIDE], gWc1[2] = {0.f, 0.f}, gWc3[2] = {0.f, 0.f}, gW0a[16], gb4 = 0.f;
#pragma unroll
        for (int o = 0; o < OSTRIDE; ++o) gW4a[o] = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) gW0a[j] = 0.f;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long n0 = tile * NPIX;
            if (tid < NPIX) {
                float e[16];
                embed_pixel(D, A.img, n0 + tid, A.N, e);
#pragma unroll
                for (int k = 0; k < 16; k += 4) *reinterpret_cast<float4*>(sPT + tid * 16 + k) = make_float4(e[k], e[k + 1], e[k + 2], e[k + 3]);
            }
            // dL/do = g_out * act'(o), scaled
            for (int i = tid; i < OSTRIDE * NPIX; i += NEPI) {
                const int p = i / OSTRIDE, o = i % OSTRIDE; const long long n = n0 + p;
                float g = 0.f;
                if (o < D.n_out && n < A.N) {
                    const float v = A.oc[n * OSTRIDE + o], gy = A.g_out[n * D.n_out + o];
                    if (D.otype == 0) g = gy * (v > 20.f ? 1.f : 1.f / (1.f + expf(-v)));
                    else { const float t = tanhf(v); g = gy * 1.3f * (1.f - t * t); }
                }
                sG4[i] = g * gs;
            }
            named_bar_sync(1, NEPI);
            if (tid < OSTRIDE) { float sum = 0.f; for (int p = 0; p < NPIX; ++p) sum += sG4[p * OSTRIDE + tid]; gb4 += sum; }
            // ---- through lin4 and the sine of lin3 (no MMA):  gz3 = (W4^T G4) * cos z3 ; gW4 += G4 sin z3
            for (int u = 0; u < NSUB; ++u) {
                const int pbeg = u * SPIX + pq;
                for (int p0 = pbeg; p0 < pbeg + SPIX / 2; p0 += 16) {
                    float zv[16];
                    const float* const zp = A.zc + (n0 + p0) * ZSTRIDE + 3 * HID + f;
                    const int nval = (int)min((long long)16, A.N - (n0 + p0));
#pragma unroll
                    for (int i = 0; i < 16; ++i) zv[i] = i < nval ? __ldg(zp + i * ZSTRIDE) : 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int p = p0 + i;
                        const float4 ga = *reinterpret_cast<const float4*>(sG4 + p * OSTRIDE), gb_ = *reinterpret_cast<const float4*>(sG4 + p * OSTRIDE + 4);
                        const float g4[OSTRIDE] = {ga.x, ga.y, ga.z, ga.w, gb_.x, gb_.y, gb_.z, gb_.w};
                        float gx = 0.f;
#pragma unroll
                        for (int o = 0; o < OSTRIDE; ++o) gx = fmaf(w4[o], g4[o], gx);
                        float sn, cs; sincos_cw(zv[i], sn, cs);
                        const float gz = gx * cs;
#pragma unroll
                        for (int o = 0; o < OSTRIDE; ++o) gW4a[o] = fmaf(g4[o], sn, gW4a[o]);
                        const float2 rc = *reinterpret_cast<const float2*>(sPT + p * 16);
                        gb[3] += gz; gWc3[0] = fmaf(gz, rc.x, gWc3[0]); gWc3[1] = fmaf(gz, rc.y, gWc3[1]);
                        store_x(sX, f, p, gz);
                    }
                }
                tc_fence_before(); fence_proxy_async(); mbar_arrive(P.xready(u));
            }
            // ---- lin3, lin2, lin1 backward: the accumulators hold W_L^T G_L (feature f = an INPUT of layer L = an output of layer L-1)
            for (int L = 3; L >= 1; --L) {
                const int h_prev = (L == 3 || L == 1) ? D.h0 : HID;  // sine units feeding layer L; the rest of its input is the embedding
                const bool live = f < h_prev;
                for (int u = 0; u < NSUB; ++u) {
                    mbar_wait(P.dfull(u), ph_d); tc_fence_after();
                    mbar_wait(P.sdone(u), ph_s);
                    const int pbeg = u * SPIX + pq;
                    for (int p0 = pbeg; p0 < pbeg + SPIX / 2; p0 += 16) {
                        float acc[16], zv[16];
                        const float* const zp = A.zc + (n0 + p0) * ZSTRIDE + (L - 1) * HID + f;
                        const int nval = live ? (int)min((long long)16, A.N - (n0 + p0)) : 0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) zv[i] = i < nval ? __ldg(zp + i * ZSTRIDE) : 0.f;
                        tmem_ld16(t_lane + (uint32_t)p0, acc);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int p = p0 + i;
                            const float gz = live ? acc[i] * cos_cw(zv[i]) : 0.f;       // dL/dz_{L-1}
                            gb[L - 1] += gz;
                            if (L == 2) {                                                // z1: lin1's output -> coordinate columns of gW1
                                const float2 rc = *reinterpret_cast<const float2*>(sPT + p * 16);
                                gWc1[0] = fmaf(gz, rc.x, gWc1[0]); gWc1[1] = fmaf(gz, rc.y, gWc1[1]);
                            }
                            if (L > 1) store_x(sX, f, p, gz);
                            else {                                                       // z0: lin0's weight gradient, K = embedding, exact FP32
                                const float4* e4 = reinterpret_cast<const float4*>(sPT + p * 16);
                                const float4 e0 = e4[0], e1 = e4[1], e2 = e4[2], e3 = e4[3];
                                gW0a[0] = fmaf(gz, e0.x, gW0a[0]); gW0a[1] = fmaf(gz, e0.y, gW0a[1]); gW0a[2] = fmaf(gz, e0.z, gW0a[2]); gW0a[3] = fmaf(gz, e0.w, gW0a[3]);
                                gW0a[4] = fmaf(gz, e1.x, gW0a[4]); gW0a[5] = fmaf(gz, e1.y, gW0a[5]); gW0a[6] = fmaf(gz, e1.z, gW0a[6]); gW0a[7] = fmaf(gz, e1.w, gW0a[7]);
                                gW0a[8] = fmaf(gz, e2.x, gW0a[8]); gW0a[9] = fmaf(gz, e2.y, gW0a[9]); gW0a[10] = fmaf(gz, e2.z, gW0a[10]); gW0a[11] = fmaf(gz, e2.w, gW0a[11]);
                                gW0a[12] = fmaf(gz, e3.x, gW0a[12]); gW0a[13] = fmaf(gz, e3.y, gW0a[13]); gW0a[14] = fmaf(gz, e3.z, gW0a[14]); gW0a[15] = fmaf(gz, e3.w, gW0a[15]);
                            }
                        }
                    }
                    if (L > 1) { tc_fence_before(); fence_proxy_async(); mbar_arrive(P.xready(u)); }
                }
                ph_d ^= 1; ph_s ^= 1;
            }
            tc_fence_before();
            named_bar_sync(1, NEPI);                                  // sPT / sG4 / the image are rewritten by the next tile
        }
        // ---- flush the register accumulators (unscaled) into the packed gradient vector
        float* gp = A.g_params;
#pragma unroll
        for (int L = 0; L < 4; ++L) {
            const int n_out = (L == 0 || L == 2) ? D.h0 : HID;
            if (f < n_out) atomicAdd(gp + D.ob[L] + f, gb[L] * inv_gs);
        }
#pragma unroll
        for (int o = 0; o < OSTRIDE; ++o) if (o < D.n_out) atomicAdd(gp + D.oW[4] + o * HID + f, gW4a[o] * inv_gs);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            atomicAdd(gp + D.oW[1] + f * HID + D.h0 + c, gWc1[c] * inv_gs);
            atomicAdd(gp + D.oW[3] + f * HID + D.h0 + c, gWc3[c] * inv_gs);
        }
        if (f < D.h0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j < D.d0) atomicAdd(gp + D.oW[0] + f * D.d0 + j, gW0a[j] * inv_gs);
        }
        if (tid < D.n_out) atomicAdd(gp + D.ob[4] + tid, gb4 * inv_gs);
    } else if (warp == WARP_PROD) {
        if (lane == 0) producer_loop<false>(P, smem, A.wprepT, ntiles);
    } else if (warp == WARP_MMA) {
        if (lane == 0) mma_loop<false>(P, smem, tmem_base, ntiles);
    } else {
        store_loop<false>(P, smem, A.gimg, ntiles, lane);
    }
    pipe_teardown(warp, tmem_base);
}

// ---------------------------------------------------------------- backward: weight gradients of lin1..lin3
// gW_l[c][k] = sum_pixels G_l[c][p] X_l[k][p]: one CTA per (layer, pixel range) with the whole 256 x 256 FP32 result in TMEM
// (2 halves x 256 columns = all 512 columns); operands are the stored FP16-split images, MN-major, 32 pixels per stage.
constexpr int WG_STAGES = 3;
constexpr int WG_STAGE_BYTES = 2 * IMG_PIECE;                 // [G piece | X piece]
constexpr int WG_SM_BAR = WG_STAGES * WG_STAGE_BYTES;
constexpr int WG_SM_TOTAL = WG_SM_BAR + 128;
constexpr int WG_THREADS = 192;                               // producer warp, MMA warp, 4 epilogue warps
constexpr uint32_t kIdescWg = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct WgArgs { Dims D; const uint8_t* ximg; const uint8_t* gimg; long long ntiles; float* g_params; const float* gmax; };

__global__ void __launch_bounds__(WG_THREADS, 1) posmlp_wgrad_tc_kernel(const __grid_constant__ WgArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Dims& D = A.D;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slot = blockIdx.y, L = 1 + slot;
    const long long t0 = A.ntiles * blockIdx.x / gridDim.x, t1 = A.ntiles * (blockIdx.x + 1) / gridDim.x;
    const long long npieces = (t1 - t0) * 4;
    if (npieces <= 0) return;
    const uint32_t bar0 = smem_u32(smem + WG_SM_BAR);
    const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * WG_STAGES, bar_done = bar0 + 16 * WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + WG_SM_BAR + 16 * WG_STAGES + 16);
    if (tid == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (long long i = 0; i < npieces; ++i) {
                const uint32_t s = (uint32_t)(i % WG_STAGES), par = (uint32_t)((i / WG_STAGES) & 1);
                const size_t piece = (((size_t)slot * A.ntiles + (t0 + i / 4)) * 4 + (i & 3)) * IMG_PIECE;
                mbar_wait(bar_empty + 8 * s, par ^ 1);
                mbar_expect_tx(bar_full + 8 * s, WG_STAGE_BYTES);
                const uint32_t dst = smem_u32(smem + s * WG_STAGE_BYTES);
                bulk_g2s(dst, A.gimg + piece, IMG_PIECE, bar_full + 8 * s);
                bulk_g2s(dst + IMG_PIECE, A.ximg + piece, IMG_PIECE, bar_full + 8 * s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (long long i = 0; i < npieces; ++i) {
                const uint32_t s = (uint32_t)(i % WG_STAGES), par = (uint32_t)((i / WG_STAGES) & 1);
                mbar_wait(bar_full + 8 * s, par);
                tc_fence_after();
                const uint32_t gst = smem_u32(smem + s * WG_STAGE_BYTES), xst = gst + IMG_PIECE;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {                              // K = 16 pixels per MMA, 32 per piece
                        const uint64_t a1 = umma_desc_mn(gst + h * 8192 + j * 256, 128, 512), a2 = umma_desc_mn(gst + IMG_PIECE / 2 + h * 8192 + j * 256, 128, 512);
                        const uint64_t b1 = umma_desc_mn(xst + j * 256, 128, 512), b2 = umma_desc_mn(xst + IMG_PIECE / 2 + j * 256, 128, 512);
                        const uint32_t d = tmem_base + (uint32_t)(h * HID);
                        umma_f16(d, a1, b1, kIdescWg, (i | j) != 0);            // g1 x1
                        umma_f16(d, a1, b2, kIdescWg, 1);                       // g1 x2
                        umma_f16(d, a2, b1, kIdescWg, 1);                       // g2 x1
                    }
                tc_commit(bar_empty + 8 * s);
            }
            tc_commit(bar_done);
        }
    } else {
        float gs, inv_gs; grad_scale(__ldg(A.gmax), gs, inv_gs);
        const int quarter = warp & 3;
        const int n_out = (L == 2) ? D.h0 : HID;
        mbar_wait(bar_done, 0); tc_fence_after();
        for (int h = 0; h < 2; ++h) {
            const int c = h * 128 + quarter * 32 + lane;
            float* row = A.g_params + D.oW[L] + (size_t)c * HID;
            for (int k0 = 0; k0 < HID; k0 += 16) {
                float acc[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(h * HID + k0), acc);
                if (c < n_out) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(row + k0 + i, acc[i] * inv_gs);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ out) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = fabsf(x[i]);
        if (v < 3.0e38f) m = fmaxf(m, v);                     // ignores inf / nan
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));   // non-negative floats order like their bit patterns
}

constexpr size_t kWsFwd = (size_t)NLAYER * NCHUNK * W_STAGE, kWsBwd = (size_t)3 * NCHUNK * W_STAGE;

}  // namespace

size_t tc_workspace_bytes() { return kWsFwd + kWsBwd + 256; }
size_t tc_image_bytes(long long N) { return (size_t)3 * (size_t)((N + NPIX - 1) / NPIX) * IMG_TILE; }

int tc_forward(const Dims& D, const float* params, const float* img, long long N, float* out, float* zc, float* oc, void* ximg,
               void* workspace, cudaStream_t st) {
    if (((uintptr_t)workspace & 15) != 0 || ((uintptr_t)ximg & 15) != 0) return MB200_EINVAL;
    posmlp_prep_kernel<false><<<160, 256, 0, st>>>(D, params, reinterpret_cast<__half*>(workspace));
    int rc = mb200_check_launch();
    if (rc) return rc;
    rc = mb200_check(cudaFuncSetAttribute(posmlp_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    if (rc) return rc;
    FwdArgs A; A.D = D; A.params = params; A.img = img; A.N = N; A.out = out; A.zc = zc; A.oc = oc;
    A.wprep = reinterpret_cast<const __half*>(workspace); A.ximg = reinterpret_cast<uint8_t*>(ximg);
    const long long ntiles = (N + NPIX - 1) / NPIX;
    const int grid = (int)(ntiles < (long long)mb200_sm_count() ? ntiles : (long long)mb200_sm_count());
    posmlp_fwd_tc_kernel<<<grid, NTHREADS, SM_TOTAL, st>>>(A);
    return mb200_check_launch();
}

int tc_backward(const Dims& D, const float* params, const float* img, long long N, const float* zc, const float* oc, const void* ximg,
                void* gimg, const float* g_out, float* g_params, void* workspace, cudaStream_t st) {
    if (((uintptr_t)workspace & 15) != 0 || ((uintptr_t)ximg & 15) != 0 || ((uintptr_t)gimg & 15) != 0) return MB200_EINVAL;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    __half* wprepT = reinterpret_cast<__half*>(ws + kWsFwd);
    unsigned int* gmax = reinterpret_cast<unsigned int*>(ws + kWsFwd + kWsBwd);
    int rc = mb200_check(cudaMemsetAsync(gmax, 0, 4, st));
    if (rc) return rc;
    absmax_kernel<<<mb200_sm_count() * 2, 256, 0, st>>>(g_out, N * D.n_out, gmax);
    posmlp_prep_kernel<true><<<96, 256, 0, st>>>(D, params, wprepT);
    rc = mb200_check_launch();
    if (rc) return rc;
    rc = mb200_check(cudaFuncSetAttribute(posmlp_bwd_data_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    if (rc) return rc;
    rc = mb200_check(cudaFuncSetAttribute(posmlp_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SM_TOTAL));
    if (rc) return rc;
    const long long ntiles = (N + NPIX - 1) / NPIX;
    BwdArgs B; B.D = D; B.params = params; B.img = img; B.N = N; B.zc = zc; B.oc = oc; B.g_out = g_out; B.g_params = g_params;
    B.wprepT = wprepT; B.gimg = reinterpret_cast<uint8_t*>(gimg); B.gmax = reinterpret_cast<const float*>(gmax);
    const int grid = (int)(ntiles < (long long)mb200_sm_count() ? ntiles : (long long)mb200_sm_count());
    posmlp_bwd_data_tc_kernel<<<grid, NTHREADS, SM_TOTAL, st>>>(B);
    rc = mb200_check_launch();
    if (rc) return rc;
    WgArgs W; W.D = D; W.ximg = reinterpret_cast<const uint8_t*>(ximg); W.gimg = reinterpret_cast<const uint8_t*>(gimg);
    W.ntiles = ntiles; W.g_params = g_params; W.gmax = reinterpret_cast<const float*>(gmax);
    const int per_layer = mb200_sm_count() / 3;
    const int nsplit = (int)(ntiles < (long long)per_layer ? ntiles : (long long)per_layer);
    posmlp_wgrad_tc_kernel<<<dim3((unsigned)nsplit, 3), WG_THREADS, WG_SM_TOTAL, st>>>(W);
    return mb200_check_launch();
}

}  // namespace posmlp
