// temporary: PosMLP / CDF / SH entry points (replaced by mb200_posmlp.cu / mb200_envutils.cu)
#include "mb200_host.h"
extern "C" {
int64_t mb200_posmlp_param_count(const mb200_posmlp_desc*) { return -1; }
size_t  mb200_posmlp_cache_bytes(const mb200_posmlp_desc*, int64_t) { return 0; }
int mb200_posmlp_fwd(const mb200_posmlp_desc*, const float*, const float*, int64_t, float*, void*, void*) { return MB200_EUNSUPPORTED; }
int mb200_posmlp_bwd(const mb200_posmlp_desc*, const float*, const float*, int64_t, const void*, const float*, float*, float*, void*) { return MB200_EUNSUPPORTED; }
int mb200_cdf_build(const float*, int, int, float*, float*, void*) { return MB200_EUNSUPPORTED; }
int mb200_cdf_sample(const float*, const float*, int, int, const float*, int64_t, float*, float*, int64_t*, int64_t*, void*) { return MB200_EUNSUPPORTED; }
int mb200_sh_project(const double*, int, int, const double*, int64_t, double*, void*) { return MB200_EUNSUPPORTED; }
int mb200_sh_reconstruct(const double*, int, int, int, double*, void*) { return MB200_EUNSUPPORTED; }
}
