// emul.cpp — the G-buffer kernels' own per-sample device code (mb200_device.cuh, mb200_shade.cuh) compiled for the HOST.
// TEST INFRASTRUCTURE (tests/test_host_emulation.py): checks, without a GPU, that the decisions the device source takes
// (hierarchy cell, texel, lobe, envmap cells, the bits of both sampled directions) equal the oracle's on every lane, i.e.
// that the source mirrors the oracle operation for operation.  What it cannot see is nvcc's own contraction of an
// expression that was left un-annotated; the -m gpu test test_gpu_render_parity.py::test_sample_record_bit_exact does.
#include "cuda_shim.h"
#include "mb200_shade.cuh"

int mb200_sm_count() { return 1; }
int mb200_check_launch() { return 0; }
int mb200_check(cudaError_t) { return 0; }

extern "C" int emul_sample_record(const mb200_cfg* c, const float* gpos, const float* gnrm, const float* a, const float* r, const float* m,
                                  const float* n_opt, const float* env4, const float* hier, const mb200_hier_desc* d,
                                  int32_t* out, float* out_L) {
    RenderParams P; int rc = fill_params(c, gpos, gnrm, a, r, m, n_opt, env4, hier, d, P);
    if (rc) return rc;
    P.prow0 = c->row0; P.prows = c->rows;
    const long long n = (long long)P.prows * P.W * P.spp;
    const bool ad = (c->flags & MB200_FLAG_AD_WEIGHTS) != 0;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) {
        const int s = (int)(i % P.spp); const long long pix = i / P.spp;
        const int py = P.prow0 + (int)(pix / P.W), px = (int)(pix % P.W);
        const int gpix = py * P.W + px;
        const PixelCtx ctx = load_pixel(P, gpix);
        SampleDbg g; float jx, jy;
        const uint32_t lane = (uint32_t)gpix * (uint32_t)P.spp + (uint32_t)s;
        const float3 L = ad ? shade_sample<true, false, true>(P, ctx, px, py, lane, jx, jy, &g)
                            : shade_sample<false, false, true>(P, ctx, px, py, lane, jx, jy, &g);
        int32_t* o = out + 12 * i;
        o[0] = (int32_t)g.ox; o[1] = (int32_t)g.oy; o[2] = (int32_t)g.flat; o[3] = g.lobe; o[4] = g.em_i00; o[5] = g.bs_i00;
        o[6] = __float_as_int(g.d_em.x); o[7] = __float_as_int(g.d_em.y); o[8] = __float_as_int(g.d_em.z);
        o[9] = __float_as_int(g.d_bs.x); o[10] = __float_as_int(g.d_bs.y); o[11] = __float_as_int(g.d_bs.z);
        if (out_L) { out_L[3 * i] = L.x; out_L[3 * i + 1] = L.y; out_L[3 * i + 2] = L.z; }
    }
    return 0;
}
