// s4_gemm dispatcher: tcgen05 path when the problem fits it, CUDA-core path otherwise.
#include <stdlib.h>

#include "common.cuh"
#include "gemm_params.h"

static int env_force_simt() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("S4_FORCE_SIMT");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

extern "C" int s4_gemm_uses_tc(const S4GemmParams* p) {
  if (p->backend == S4_BACKEND_SIMT || env_force_simt()) return 0;
  return s4_gemm_tc_supported(*p) ? 1 : 0;
}

extern "C" int s4_gemm(const S4GemmParams* p, cudaStream_t stream) {
  S4_REQUIRE(p->M >= 0 && p->N >= 0 && p->K >= 0, "gemm: negative dimension");
  S4_REQUIRE(p->nb1 >= 1 && p->nb2 >= 1, "gemm: batch counts must be >= 1");
  if (s4_gemm_uses_tc(p)) return s4_gemm_tc_launch(*p, stream);
  if (p->colsum) {
    s4_set_error("gemm: the fused column sum exists on the tcgen05 path only (check s4_gemm_uses_tc)");
    return S4_ERR_UNSUPPORTED;
  }
  if (p->backend == S4_BACKEND_TC) {
    s4_set_error("gemm: tcgen05 path does not support this problem (M=%d N=%d K=%d dtype=%d)",
                 p->M, p->N, p->K, p->dtype);
    return S4_ERR_UNSUPPORTED;
  }
  return s4_gemm_simt_launch(*p, stream);
}
