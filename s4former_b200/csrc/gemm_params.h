// Internal view of the public GEMM parameter block + launcher prototypes.
#pragma once
#include "../../include/s4former.h"

int s4_gemm_simt_launch(const S4GemmParams& p, cudaStream_t stream);
// returns S4_ERR_UNSUPPORTED (without setting an error) when the tcgen05 path cannot run p
int s4_gemm_tc_launch(const S4GemmParams& p, cudaStream_t stream);
bool s4_gemm_tc_supported(const S4GemmParams& p);
