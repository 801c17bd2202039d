// Fused flash-style attention with the in-tile PASA bias on tcgen05 (placeholder dispatch:
// until the fused kernel lands every shape is served by the composed path in attention.cu).
#include "common.cuh"

bool s4_attention_tc_supported(int B, int H, int L, int hd, int dtype) { return false; }

int s4_attention_tc_fwd(const void* qkv, const float* u0, const float* gate, float w, void* out,
                        float* lse, int B, int H, int L, int hd, cudaStream_t st) {
  s4_set_error("attention_tc_fwd: not available");
  return S4_ERR_UNSUPPORTED;
}

int s4_attention_tc_bwd(const void* dout, const void* qkv, const void* out, const float* lse,
                        const float* u0, const float* gate, float w, void* dqkv, void* ws,
                        size_t ws_bytes, int B, int H, int L, int hd, cudaStream_t st) {
  s4_set_error("attention_tc_bwd: not available");
  return S4_ERR_UNSUPPORTED;
}
