// mb200_host.h — host-side helpers shared by the translation units of libmaterialist_b200.so
#pragma once
#include <cuda_runtime.h>
#include <string.h>
#include "../../include/materialist_b200.h"

// number of SMs of the current device (cached per device); 148 on B200
int mb200_sm_count();
// cudaGetLastError() -> MB200_OK / MB200_ELAUNCH, remembering the message for mb200_last_cuda_error()
int mb200_check_launch();
int mb200_check(cudaError_t e);
