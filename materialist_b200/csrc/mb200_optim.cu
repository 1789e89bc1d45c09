// mb200_optim.cu — the loss and optimiser step that sit between the forward and the adjoint render of one
// BRDF-phase iteration (inverse_img_w_mi.py:388-432), fused into four small kernels so that the iteration is
// ~9 launches instead of ~100 elementwise torch launches and never syncs with the host:
//
//   image_sum      Σ pred                       -> ratio = gt.mean() / pred.detach().mean()      (:388)
//   srgb_sums      Σ diff², Σ |diff|            -> loss_mse, loss_l1, scale_raito                 (:391-395, :415)
//   srgb_grad      d loss / d pred              (autograd of :389-418 written out)
//   adam_clamped   clamp backward + aux L1 gradient + Adam + clamp forward of the next iteration (:370-376, :398-417, :428)
//
// All three are HBM-bound streaming passes over (rows*W*3) floats; reductions are two-stage with a FIXED
// summation order (per-thread strided partial -> warp shuffle -> block -> the last block sums the per-block
// partials), so results are bitwise reproducible run to run.
#include <math.h>
#include "mb200_host.h"

namespace {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 1024;
constexpr float kSrgbExp = 1.0f / 2.2f;        // myutils/misc.py:167-170  image ** (1/2.2)

struct RedScratch { float partial[2][kRedMaxBlocks]; unsigned int ticket; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-level reduce of NV values, then the last block to arrive reduces the per-block partials in index order
template <int NV>
__device__ __forceinline__ bool finish_reduce(float (&v)[NV], RedScratch* sc, float* out, float* s_tot = nullptr) {
    __shared__ float s_part[NV][kRedThreads / 32];
    __shared__ bool s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const float w = warp_sum(v[k]);
        if (lane == 0) s_part[k][warp] = w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float s = 0.f;
            for (int i = 0; i < kRedThreads / 32; ++i) s += s_part[k][i];
            sc->partial[k][blockIdx.x] = s;
        }
        __threadfence();
        const unsigned int t = atomicAdd(&sc->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    if (warp < NV) {
        // one warp per value: lane-strided partial sums (fixed order), then a shuffle tree
        float s = 0.f;
        for (int i = lane; i < (int)gridDim.x; i += 32) s += __ldcg(&sc->partial[warp][i]);
        s = warp_sum(s);
        if (lane == 0) { out[warp] = s; if (s_tot) s_tot[warp] = s; }
    }
    if (threadIdx.x == 0) sc->ticket = 0u;           // ready for the next launch on the same stream
    return true;
}

// ---------------------------------------------------------------- multi-GPU: the exchange steps fused into these kernels
// One process per GPU; every rank owns a MAILBOX in its own HBM that its peers write straight into over NVLink / NVSwitch
// (peer pointers from CUDA IPC, mb200_peer_open).  The three scalar sums of an iteration (Σ pred; Σ diff², Σ |diff|) are not
// all-reduced by a collective library: the last block of the producing kernel stores the rank's partial into slot [rank] of EVERY
// peer's mailbox and then raises that peer's flag (release, system scope); the consuming kernel — the very next kernel of the
// iteration — spins on its own mailbox until all `world` flags carry this iteration's sequence number (acquire) and adds the
// partials in RANK order, so every rank computes the bitwise identical total.  No extra launch, no host involvement, ~2 us of
// NVLink latency instead of two NCCL launches.  Slots are double-buffered by sequence parity: a rank can run at most one exchange
// ahead of a peer (it needs that peer's partial of the same iteration to go on), so parity keeps the peer's unread values intact.
struct PeerBox {
    float pred[2][MB200_MAX_PEERS]; float sums[2][MB200_MAX_PEERS][2];
    unsigned int f_pred[2][MB200_MAX_PEERS], f_sums[2][MB200_MAX_PEERS];
    unsigned int f_halo[2][2], f_map[2][2];          // [parity][0: written by the upper neighbour, 1: by the lower neighbour]
};
struct PeerView { int rank, world; unsigned int seq; PeerBox* box[MB200_MAX_PEERS]; };

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void spin_until(const unsigned int* flag, unsigned int seq) { while (ld_acquire_sys(flag) != seq) __nanosleep(64); }
__device__ __forceinline__ float ld_volatile_f32(const float* p) { float v; asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }

// threads r < world of the calling block: store NV totals into slot [rank] of peer r's mailbox, then raise its flag
template <int NV>
__device__ __forceinline__ void peer_publish(const PeerView& pv, const float* s_tot) {
    __syncthreads();                                   // s_tot written by the reducing warps
    const int r = threadIdx.x;
    if (r >= pv.world) return;
    PeerBox* b = pv.box[r]; const int par = pv.seq & 1u;
    if (NV == 1) { b->pred[par][pv.rank] = s_tot[0]; __threadfence_system(); st_release_sys(&b->f_pred[par][pv.rank], pv.seq); }
    else { b->sums[par][pv.rank][0] = s_tot[0]; b->sums[par][pv.rank][1] = s_tot[1]; __threadfence_system(); st_release_sys(&b->f_sums[par][pv.rank], pv.seq); }
}
// thread 0 of the calling block: wait for every rank's flag, add the partials in rank order -> s_out[0..NV) (then __syncthreads)
template <int NV>
__device__ __forceinline__ void peer_collect(const PeerView& pv, float* s_out) {
    if (threadIdx.x == 0) {
        const PeerBox* b = pv.box[pv.rank]; const int par = pv.seq & 1u;
        float t0 = 0.f, t1 = 0.f;
        for (int r = 0; r < pv.world; ++r) {
            if (NV == 1) { spin_until(&b->f_pred[par][r], pv.seq); t0 += ld_volatile_f32(&b->pred[par][r]); }
            else { spin_until(&b->f_sums[par][r], pv.seq); t0 += ld_volatile_f32(&b->sums[par][r][0]); t1 += ld_volatile_f32(&b->sums[par][r][1]); }
        }
        s_out[0] = t0; if (NV == 2) s_out[1] = t1;
    }
    __syncthreads();
}

template <bool PEER>
__global__ void __launch_bounds__(kRedThreads) image_sum_kernel(const float* __restrict__ img, long long n, RedScratch* sc, float* out,
                                                                const __grid_constant__ PeerView pv) {
    __shared__ float s_tot[2];
    float v[1] = {0.f};
    const long long n4 = n >> 2;
    const float4* img4 = reinterpret_cast<const float4*>(img);
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kRedThreads) {
        const float4 q = __ldg(img4 + i);
        v[0] += (q.x + q.y) + (q.z + q.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) v[0] += img[(n4 << 2) + threadIdx.x];
    const bool last = finish_reduce<1>(v, sc, out, s_tot);
    if (PEER && last) peer_publish<1>(pv, s_tot);      // out[0] stays this rank's partial; the global sum is formed by the consumer
}

// ratio = num/den read from device memory: scal[0] = Σ gt (global), scal[1] = Σ pred (global)
__device__ __forceinline__ float srgb_of(float x) { return x > 0.f ? powf(x, kSrgbExp) : 0.f; }

template <bool PEER>
__global__ void __launch_bounds__(kRedThreads) srgb_sums_kernel(const float* __restrict__ img, const float* __restrict__ gt_srgb, long long n,
                                                                float* __restrict__ scal, RedScratch* sc, float* out2,
                                                                float* __restrict__ pred_srgb, const __grid_constant__ PeerView pv) {
    __shared__ float s_tot[2];
    float pred_sum;
    if (PEER) {          // Σ pred over all ranks: collected from the mailbox (every block, so that nobody waits on a global write)
        peer_collect<1>(pv, s_tot);
        pred_sum = s_tot[0];
        if (blockIdx.x == 0 && threadIdx.x == 0) scal[1] = pred_sum;        // for the host-visible record (same value on every rank)
        __syncthreads();
    } else pred_sum = scal[1];
    const float ratio = __fdiv_rn(scal[0], pred_sum);
    float v[2] = {0.f, 0.f};
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kRedThreads) {
        const float y = srgb_of(__ldg(img + i) * ratio);
        const float d = y - __ldg(gt_srgb + i);
        v[0] = fmaf(d, d, v[0]); v[1] += fabsf(d);
        if (pred_srgb) pred_srgb[i] = y;
    }
    const bool last = finish_reduce<2>(v, sc, out2, s_tot);
    if (PEER && last) peer_publish<2>(pv, s_tot);
}

// loss = 3 * (S1/S0) * Σdiff²/n_total + Σ|diff|/n_total with S1/S0 and ratio detached (:388, :415-417)
template <bool PEER>
__global__ void __launch_bounds__(256) srgb_grad_kernel(const float* __restrict__ img, const float* __restrict__ gt_srgb, long long n,
                                                        const float* __restrict__ scal, float* __restrict__ sums2, float inv_n_total,
                                                        float* __restrict__ grad, const __grid_constant__ PeerView pv) {
    __shared__ float s_tot[2];
    float S0, S1;
    if (PEER) {          // Σ diff², Σ |diff| over all ranks
        peer_collect<2>(pv, s_tot);
        S0 = s_tot[0]; S1 = s_tot[1];
        if (blockIdx.x == 0 && threadIdx.x == 0) { sums2[0] = S0; sums2[1] = S1; }
    } else { S0 = sums2[0]; S1 = sums2[1]; }
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float ratio = __fdiv_rn(scal[0], scal[1]);
    const float k_mse = 6.f * __fdiv_rn(S1, S0) * inv_n_total;                  // 3 * scale_raito * 2 / n
    const float x = __ldg(img + i) * ratio;
    float g = 0.f;
    if (x > 0.f) {
        const float y = powf(x, kSrgbExp);
        const float d = y - __ldg(gt_srgb + i);
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        const float dy = fmaf(k_mse, d, sgn * inv_n_total);
        g = dy * (kSrgbExp * __fdiv_rn(y, x)) * ratio;                          // d(x^e)/dx = e * x^(e-1) = e * y / x
    }
    grad[i] = g;
}

struct AdamSeg { float* p; float* mat; const float* g; const float* ori; float* m; float* v; long long n; float lo, hi, aux; };
struct AdamParams { AdamSeg seg[MB200_ADAM_MAX_SEGS]; int nseg; float lr_over_bc1, inv_sqrt_bc2, beta1, beta2, eps; };

// torch.optim.Adam (single-tensor formulas) on p with gradient
//   mask(lo <= p <= hi) * (g_render + aux * sign(clamp(p) - ori))       [clamp backward + NF.l1_loss(mat, ori) * scale_delta]
// then mat = clamp(p_new) — the map the next iteration renders with.
__global__ void __launch_bounds__(256) adam_clamped_kernel(const __grid_constant__ AdamParams P) {
    const AdamSeg& s = P.seg[blockIdx.y];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += (long long)gridDim.x * blockDim.x) {
        float p = s.p[i];
        float g = 0.f;
        if (p >= s.lo && p <= s.hi) {
            const float d = p - s.ori[i];                   // inside the clamp range clamp(p) == p
            g = s.g[i] + s.aux * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        }
        float m = s.m[i], v = s.v[i];
        m = fmaf(g - m, 1.f - P.beta1, m);                  // exp_avg.lerp_(grad, 1 - beta1)
        v = fmaf(v, P.beta2, (1.f - P.beta2) * g * g);      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = __fsqrt_rn(v) * P.inv_sqrt_bc2 + P.eps;
        p = p - P.lr_over_bc1 * __fdiv_rn(m, denom);
        s.m[i] = m; s.v[i] = v; s.p[i] = p;
        s.mat[i] = fminf(fmaxf(p, s.lo), s.hi);
    }
}

// Halo rows: `nseg` contiguous float4 runs copied from this rank's memory straight into a neighbour's (peer stores over NVLink); the
// last block to finish raises the neighbours' flags.  Used for the 2-row film halo of d(loss)/d(image) (between the loss and the
// adjoint render) and for the stepped material maps of the boundary rows (after Adam).
struct PushSeg { const float4* src; float4* dst; long long n4; };
struct PushParams { PushSeg seg[MB200_PEER_MAX_PUSH]; int nseg; unsigned int* flag[2]; unsigned int seq; unsigned int* ticket; };
__global__ void __launch_bounds__(256) peer_push_kernel(const __grid_constant__ PushParams P) {
    for (int k = 0; k < P.nseg; ++k) {
        const PushSeg& s = P.seg[k];
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s.n4; i += (long long)gridDim.x * blockDim.x) s.dst[i] = s.src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(P.ticket, 1u);
        if (t == gridDim.x - 1) {
            *P.ticket = 0u;
            __threadfence_system();
            if (P.flag[0]) st_release_sys(P.flag[0], P.seq);
            if (P.flag[1]) st_release_sys(P.flag[1], P.seq);
        }
    }
}
__global__ void peer_wait_kernel(const unsigned int* f0, const unsigned int* f1, unsigned int seq) {
    if (f0) spin_until(f0, seq);
    if (f1) spin_until(f1, seq);
}

int fill_peer(const mb200_peer* p, PeerView& pv) {
    if (!p || p->world < 2 || p->world > MB200_MAX_PEERS || p->rank < 0 || p->rank >= p->world || p->seq == 0) return MB200_EINVAL;
    pv.rank = p->rank; pv.world = p->world; pv.seq = p->seq;
    for (int r = 0; r < MB200_MAX_PEERS; ++r) pv.box[r] = nullptr;
    for (int r = 0; r < p->world; ++r) { if (!p->box[r]) return MB200_EINVAL; pv.box[r] = (PeerBox*)p->box[r]; }
    return MB200_OK;
}

int red_grid(long long n) {
    long long b = (n + kRedThreads * 8 - 1) / (kRedThreads * 8);
    const long long cap = (long long)mb200_sm_count() * 4;
    if (b > cap) b = cap;
    if (b > kRedMaxBlocks) b = kRedMaxBlocks;
    return b < 1 ? 1 : (int)b;
}

}  // namespace

extern "C" {

size_t mb200_reduce_scratch_bytes(void) { return sizeof(RedScratch); }

int mb200_image_sum(const float* img, int64_t n, float* out, void* scratch, void* stream) {
    if (!img || !out || !scratch || n <= 0) return MB200_EINVAL;
    if (((uintptr_t)img & 15) != 0) return MB200_EINVAL;
    image_sum_kernel<false><<<red_grid(n), kRedThreads, 0, (cudaStream_t)stream>>>(img, (long long)n, (RedScratch*)scratch, out, PeerView{});
    return mb200_check_launch();
}

int mb200_loss_srgb_sums(const float* img, const float* gt_srgb, int64_t n, const float* gt_pred_sums, float* out2,
                         float* pred_srgb_opt, void* scratch, void* stream) {
    if (!img || !gt_srgb || !gt_pred_sums || !out2 || !scratch || n <= 0) return MB200_EINVAL;
    srgb_sums_kernel<false><<<red_grid(n), kRedThreads, 0, (cudaStream_t)stream>>>(img, gt_srgb, (long long)n, const_cast<float*>(gt_pred_sums),
                                                                                   (RedScratch*)scratch, out2, pred_srgb_opt, PeerView{});
    return mb200_check_launch();
}

int mb200_loss_srgb_grad(const float* img, const float* gt_srgb, int64_t n, const float* gt_pred_sums, const float* sums2,
                         int64_t n_total, float* grad, void* stream) {
    if (!img || !gt_srgb || !gt_pred_sums || !sums2 || !grad || n <= 0 || n_total <= 0) return MB200_EINVAL;
    const int tb = 256;
    srgb_grad_kernel<false><<<(unsigned)((n + tb - 1) / tb), tb, 0, (cudaStream_t)stream>>>(img, gt_srgb, (long long)n, gt_pred_sums,
                                                                                            const_cast<float*>(sums2), 1.f / (float)n_total, grad, PeerView{});
    return mb200_check_launch();
}

// ---- the same three kernels with the cross-rank sums exchanged through the peers' mailboxes (see PeerBox above)
size_t mb200_peer_box_bytes(void) { return (sizeof(PeerBox) + 255) & ~(size_t)255; }

int mb200_image_sum_peer(const float* img, int64_t n, float* out, void* scratch, const mb200_peer* peer, void* stream) {
    if (!img || !out || !scratch || n <= 0 || ((uintptr_t)img & 15) != 0) return MB200_EINVAL;
    PeerView pv; int rc = fill_peer(peer, pv); if (rc) return rc;
    image_sum_kernel<true><<<red_grid(n), kRedThreads, 0, (cudaStream_t)stream>>>(img, (long long)n, (RedScratch*)scratch, out, pv);
    return mb200_check_launch();
}
int mb200_loss_srgb_sums_peer(const float* img, const float* gt_srgb, int64_t n, float* gt_pred_sums, float* out2, float* pred_srgb_opt,
                              void* scratch, const mb200_peer* peer, void* stream) {
    if (!img || !gt_srgb || !gt_pred_sums || !out2 || !scratch || n <= 0) return MB200_EINVAL;
    PeerView pv; int rc = fill_peer(peer, pv); if (rc) return rc;
    srgb_sums_kernel<true><<<red_grid(n), kRedThreads, 0, (cudaStream_t)stream>>>(img, gt_srgb, (long long)n, gt_pred_sums, (RedScratch*)scratch,
                                                                                  out2, pred_srgb_opt, pv);
    return mb200_check_launch();
}
int mb200_loss_srgb_grad_peer(const float* img, const float* gt_srgb, int64_t n, const float* gt_pred_sums, float* sums2, int64_t n_total,
                              float* grad, const mb200_peer* peer, void* stream) {
    if (!img || !gt_srgb || !gt_pred_sums || !sums2 || !grad || n <= 0 || n_total <= 0) return MB200_EINVAL;
    PeerView pv; int rc = fill_peer(peer, pv); if (rc) return rc;
    const int tb = 256;
    srgb_grad_kernel<true><<<(unsigned)((n + tb - 1) / tb), tb, 0, (cudaStream_t)stream>>>(img, gt_srgb, (long long)n, gt_pred_sums, sums2,
                                                                                           1.f / (float)n_total, grad, pv);
    return mb200_check_launch();
}
// which: MB200_PEER_HALO (image-gradient halo) or MB200_PEER_MAP (material-map halo).  Copies the segments (16-byte aligned, n_float4
// float4s each; dst = the neighbour's memory as mapped here) and raises the flag of `which` in the mailboxes of the neighbours
// `to_up` / `to_down` (rank or -1).  `ticket`: one zero-initialised uint32 of device memory owned by the caller.
int mb200_peer_push(const mb200_peer* peer, int which, const mb200_push_seg* segs, int nseg, int to_up, int to_down, void* ticket, void* stream) {
    PeerView pv; int rc = fill_peer(peer, pv); if (rc) return rc;
    if (!segs || nseg < 0 || nseg > MB200_PEER_MAX_PUSH || !ticket || (which != MB200_PEER_HALO && which != MB200_PEER_MAP)) return MB200_EINVAL;
    if (to_up >= pv.world || to_down >= pv.world) return MB200_EINVAL;
    PushParams P; memset(&P, 0, sizeof(P));
    long long nmax = 1;
    for (int k = 0; k < nseg; ++k) {
        if (!segs[k].src || !segs[k].dst || segs[k].n_float4 <= 0 || (((uintptr_t)segs[k].src | (uintptr_t)segs[k].dst) & 15)) return MB200_EINVAL;
        P.seg[k].src = (const float4*)segs[k].src; P.seg[k].dst = (float4*)segs[k].dst; P.seg[k].n4 = segs[k].n_float4;
        if (segs[k].n_float4 > nmax) nmax = segs[k].n_float4;
    }
    P.nseg = nseg; P.seq = pv.seq; P.ticket = (unsigned int*)ticket;
    const int par = pv.seq & 1u;
    // the upper neighbour sees this rank as ITS lower neighbour (slot 1), and vice versa
    if (to_up >= 0)   P.flag[0] = which == MB200_PEER_HALO ? &pv.box[to_up]->f_halo[par][1]   : &pv.box[to_up]->f_map[par][1];
    if (to_down >= 0) P.flag[1] = which == MB200_PEER_HALO ? &pv.box[to_down]->f_halo[par][0] : &pv.box[to_down]->f_map[par][0];
    long long b = (nmax + 255) / 256; if (b > 64) b = 64;
    peer_push_kernel<<<(unsigned)b, 256, 0, (cudaStream_t)stream>>>(P);
    return mb200_check_launch();
}
// orders the stream after the arrival of the neighbours' pushes of `which` with sequence number peer->seq
int mb200_peer_wait(const mb200_peer* peer, int which, int from_up, int from_down, void* stream) {
    PeerView pv; int rc = fill_peer(peer, pv); if (rc) return rc;
    if (which != MB200_PEER_HALO && which != MB200_PEER_MAP) return MB200_EINVAL;
    PeerBox* own = pv.box[pv.rank]; const int par = pv.seq & 1u;
    const unsigned int* f0 = from_up >= 0 ? (which == MB200_PEER_HALO ? &own->f_halo[par][0] : &own->f_map[par][0]) : nullptr;
    const unsigned int* f1 = from_down >= 0 ? (which == MB200_PEER_HALO ? &own->f_halo[par][1] : &own->f_map[par][1]) : nullptr;
    if (!f0 && !f1) return MB200_OK;
    peer_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(f0, f1, pv.seq);
    return mb200_check_launch();
}

// ---- peer-visible device memory (CUDA IPC): the one place this library allocates, because a mailbox must be cudaMalloc memory to be
// exportable.  handle64: 64 bytes (cudaIpcMemHandle_t) to hand to the other ranks by any host-side channel.
int mb200_peer_alloc(size_t bytes, void** dev_ptr, void* handle64) {
    if (!dev_ptr || !handle64 || bytes == 0) return MB200_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    int rc = mb200_check(cudaMalloc(&p, bytes)); if (rc) return rc;
    if ((rc = mb200_check(cudaMemset(p, 0, bytes))) != MB200_OK) { cudaFree(p); return rc; }
    cudaIpcMemHandle_t h;
    if ((rc = mb200_check(cudaIpcGetMemHandle(&h, p))) != MB200_OK) { cudaFree(p); return rc; }
    memcpy(handle64, &h, 64); *dev_ptr = p;
    return mb200_check(cudaDeviceSynchronize());
}
int mb200_peer_open(const void* handle64, void** dev_ptr) {
    if (!handle64 || !dev_ptr) return MB200_EINVAL;
    cudaIpcMemHandle_t h; memcpy(&h, handle64, 64);
    return mb200_check(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
}
int mb200_peer_close(void* dev_ptr) { return dev_ptr ? mb200_check(cudaIpcCloseMemHandle(dev_ptr)) : MB200_OK; }
int mb200_peer_free(void* dev_ptr) { return dev_ptr ? mb200_check(cudaFree(dev_ptr)) : MB200_OK; }

int mb200_adam_clamped(const mb200_adam_seg* segs, int nseg, float lr, float beta1, float beta2, float eps, int step, void* stream) {
    if (!segs || nseg <= 0 || nseg > MB200_ADAM_MAX_SEGS || step <= 0) return MB200_EINVAL;
    AdamParams P; memset(&P, 0, sizeof(P));
    long long nmax = 0;
    for (int k = 0; k < nseg; ++k) {
        const mb200_adam_seg& s = segs[k];
        if (!s.p || !s.mat || !s.g || !s.m || !s.v || s.n <= 0 || (s.aux_coeff != 0.f && !s.ori)) return MB200_EINVAL;
        AdamSeg& d = P.seg[k];
        d.p = s.p; d.mat = s.mat; d.g = s.g; d.ori = s.ori ? s.ori : s.p; d.m = s.m; d.v = s.v; d.n = s.n; d.lo = s.lo; d.hi = s.hi; d.aux = s.aux_coeff;
        if (s.n > nmax) nmax = s.n;
    }
    P.nseg = nseg; P.beta1 = beta1; P.beta2 = beta2; P.eps = eps;
    P.lr_over_bc1 = (float)((double)lr / (1.0 - pow((double)beta1, (double)step)));
    P.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)step)));
    long long b = (nmax + 255) / 256; const long long cap = (long long)mb200_sm_count() * 8;
    if (b > cap) b = cap;
    adam_clamped_kernel<<<dim3((unsigned)b, (unsigned)nseg), 256, 0, (cudaStream_t)stream>>>(P);
    return mb200_check_launch();
}

}  // extern "C"
