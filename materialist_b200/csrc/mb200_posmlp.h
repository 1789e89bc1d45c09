// mb200_posmlp.h — shapes / parameter packing shared by the PosMLP translation units (FFMA and tcgen05 kernels).
#pragma once
#include "mb200_host.h"

namespace posmlp {

constexpr int HID = 256;
constexpr int ZSTRIDE = 4 * HID;  // cached pre-activations per pixel (lin0..lin3)
constexpr int OSTRIDE = 8;        // cached lin4 outputs per pixel

struct Dims {
    int n_color, n_out, d0, h0, H, W, otype, row0;
    int oW[5], ob[5];             // offsets into the packed parameter vector
};

inline Dims make_dims(const mb200_posmlp_desc* d) {
    Dims D; D.n_color = d->n_color; D.n_out = d->n_out; D.d0 = 2 + 4 * d->n_freq + d->n_color; D.h0 = HID - D.d0;
    D.H = d->H; D.W = d->W; D.otype = d->output_type; D.row0 = d->row0;
    const int in[5] = {D.d0, HID, HID, HID, HID}, out[5] = {D.h0, HID, D.h0, HID, D.n_out};
    int off = 0;
    for (int l = 0; l < 5; ++l) { D.oW[l] = off; off += in[l] * out[l]; D.ob[l] = off; off += out[l]; }
    return D;
}
inline bool valid_desc(const mb200_posmlp_desc* d) {
    return d && d->hidden == HID && d->n_freq == 2 && d->n_color >= 1 && d->n_color <= 6 && d->n_out >= 1 && d->n_out <= OSTRIDE &&
           d->H > 0 && d->W > 0 && d->row0 >= 0 && d->row0 < d->H && (d->output_type == 0 || d->output_type == 1) && (d->output_type == 0 || d->n_out == d->n_color) &&
           (d->impl == MB200_POSMLP_TCGEN05 || d->impl == MB200_POSMLP_FFMA);
}

// tcgen05 path (mb200_posmlp_tc.cu).  workspace = tc_workspace_bytes() bytes of device scratch (pre-split weight images);
// ximg / gimg = tc_image_bytes(N) bytes each: the FP16-split activation / gradient images kept for the weight-gradient GEMM.
size_t tc_workspace_bytes();
size_t tc_image_bytes(long long N);
int tc_forward(const Dims& D, const float* params, const float* img, long long N, float* out, float* zc, float* oc, void* ximg,
               void* workspace, cudaStream_t st);
int tc_backward(const Dims& D, const float* params, const float* img, long long N, const float* zc, const float* oc, const void* ximg,
                void* gimg, const float* g_out, float* g_params, void* workspace, cudaStream_t st);

}  // namespace posmlp
