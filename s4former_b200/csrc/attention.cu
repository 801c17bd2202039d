// Multi-head self-attention with S4Former's patch-adaptive (PASA) additive bias.
//
// Reference: mmcv MultiheadAttention -> torch nn.MultiheadAttention as called from
// mmseg/models/backbones/vit.py:119, with the float attn_mask built at vit.py:519-535.  The
// reference materialises a [B*heads, L, L] mask; here the bias is applied in its rank-1 form
//     bias[b,h,q,k] = w * gate[b,q] * u0[b,k]
// inside the softmax kernel, so no L x L mask tensor ever exists.
//
// This file is the COMPOSED path (contractions through s4_gemm, softmax kernels here): it backs
// the fp32 validation mode and any shape the fused tcgen05 kernel (attention_tc.cu) rejects.
// Only lse is kept between forward and backward; probabilities are recomputed.
#include "common.cuh"
#include "gemm_params.h"

static inline int pad8(int L) { return (L + 7) & ~7; }

// one warp per row: p[k] = exp(s[k] + wg*u0[k] - m) / sum ; lse = m + log(sum)
// if lse_in != null the saved lse is used instead (backward recompute).
template <typename T>
__global__ void __launch_bounds__(256)
softmax_bias_kernel(const float* __restrict__ S, T* __restrict__ P, const float* __restrict__ u0,
                    const float* __restrict__ gate, float w, float* __restrict__ lse_out,
                    const float* __restrict__ lse_in, int H, int L, int Lp, long long rows) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int q = (int)(row % L);
  const long long bh = row / L;
  const int b = (int)(bh / H);
  const float* s = S + row * Lp;
  T* p = P + row * Lp;
  const float* ub = u0 ? u0 + (size_t)b * L : nullptr;
  const float wg = (u0 && gate) ? w * gate[(size_t)b * L + q] : (u0 ? w : 0.f);
  float lse;
  if (lse_in) {
    lse = lse_in[row];
  } else {
    float m = -INFINITY;
    for (int k = lane; k < L; k += 32) {
      float v = s[k];
      if (ub) v = fmaf(wg, ub[k], v);
      m = fmaxf(m, v);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int k = lane; k < L; k += 32) {
      float v = s[k];
      if (ub) v = fmaf(wg, ub[k], v);
      sum += expf(v - m);
    }
    sum = warp_sum(sum);
    lse = m + logf(sum);
    if (lane == 0 && lse_out) lse_out[row] = lse;
  }
  for (int k = lane; k < Lp; k += 32) {
    float o = 0.f;
    if (k < L) {
      float v = s[k];
      if (ub) v = fmaf(wg, ub[k], v);
      o = expf(v - lse);
    }
    p[k] = from_f32<T>(o);   // padded tail is written as 0 so K-padded GEMMs stay exact
  }
}

// dS[k] = P[k] * (dP[k] - sum_j dP[j] P[j])
template <typename T>
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ dP, const T* __restrict__ P, T* __restrict__ dS,
                   int L, int Lp, long long rows) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* dp = dP + row * Lp;
  const T* p = P + row * Lp;
  T* ds = dS + row * Lp;
  float dot = 0.f;
  for (int k = lane; k < L; k += 32) dot = fmaf(dp[k], to_f32<T>(p[k]), dot);
  dot = warp_sum(dot);
  for (int k = lane; k < Lp; k += 32) {
    float o = 0.f;
    if (k < L) o = to_f32<T>(p[k]) * (dp[k] - dot);
    ds[k] = from_f32<T>(o);
  }
}

bool s4_attention_tc_fwd_supported(int B, int H, int L, int hd, int dtype);
bool s4_attention_tc_bwd_supported(int B, int H, int L, int hd, int dtype);
size_t s4_attention_tc_bwd_workspace(int B, int H, int L);

static size_t composed_workspace(int B, int H, int L, int dtype) {
  const size_t e = (size_t)B * H * L * pad8(L);
  const size_t ts = dtype == S4_BF16 ? 2 : 4;
  return e * 4 + 2 * e * ts + 256;
}

extern "C" size_t s4_attention_workspace(int B, int H, int L, int hd, int dtype, int backend) {
  if (backend != S4_BACKEND_SIMT && s4_attention_tc_fwd_supported(B, H, L, hd, dtype) &&
      s4_attention_tc_bwd_supported(B, H, L, hd, dtype))
    return s4_attention_tc_bwd_workspace(B, H, L);
  return composed_workspace(B, H, L, dtype);
}

static void base_params(S4GemmParams& g, int dtype, int backend) {
  g = S4GemmParams{};
  g.nb1 = 1; g.nb2 = 1; g.alpha = 1.f; g.dtype = dtype; g.c_dtype = dtype;
  g.backend = backend; g.split_k = 1;
}

template <typename T>
static int attention_fwd_t(const T* qkv, const float* u0, const float* gate, float w, T* out,
                           float* lse, void* ws, int B, int H, int L, int hd, int dtype,
                           int backend, cudaStream_t st) {
  const int D = H * hd, Lp = pad8(L);
  const size_t e = (size_t)B * H * L * Lp;
  float* S = (float*)ws;
  T* P = (T*)((char*)ws + e * 4);
  S4GemmParams g{};
  // S = scale * Q K^T
  base_params(g, dtype, backend);
  g.a = qkv; g.b = qkv + D; g.c = S; g.c_dtype = S4_F32;
  g.M = L; g.N = L; g.K = hd; g.nb1 = B; g.nb2 = H;
  g.a_sm = 3 * D; g.a_sk = 1; g.a_b1 = (long long)L * 3 * D; g.a_b2 = hd;
  g.b_sk = 1; g.b_sn = 3 * D; g.b_b1 = (long long)L * 3 * D; g.b_b2 = hd;
  g.c_sm = Lp; g.c_b1 = (long long)H * L * Lp; g.c_b2 = (long long)L * Lp;
  g.alpha = 1.0f / sqrtf((float)hd);
  int rc = s4_gemm(&g, st);
  if (rc) return rc;
  const long long rows = (long long)B * H * L;
  const int blocks = (int)((rows + 7) / 8);
  softmax_bias_kernel<T><<<blocks, 256, 0, st>>>(S, P, u0, gate, w, lse, nullptr, H, L, Lp, rows);
  rc = s4_check_launch("softmax_bias");
  if (rc) return rc;
  // O = P V
  base_params(g, dtype, backend);
  g.a = P; g.b = qkv + 2 * D; g.c = out;
  g.M = L; g.N = hd; g.K = L; g.nb1 = B; g.nb2 = H;
  g.a_sm = Lp; g.a_sk = 1; g.a_b1 = (long long)H * L * Lp; g.a_b2 = (long long)L * Lp;
  g.b_sk = 3 * D; g.b_sn = 1; g.b_b1 = (long long)L * 3 * D; g.b_b2 = hd;
  g.c_sm = D; g.c_b1 = (long long)L * D; g.c_b2 = hd;
  return s4_gemm(&g, st);
}

template <typename T>
static int attention_bwd_t(const T* dout, const T* qkv, const float* lse, const float* u0,
                           const float* gate, float w, T* dqkv, void* ws, int B, int H, int L,
                           int hd, int dtype, int backend, cudaStream_t st) {
  const int D = H * hd, Lp = pad8(L);
  const size_t e = (size_t)B * H * L * Lp;
  float* S = (float*)ws;
  T* P = (T*)((char*)ws + e * 4);
  T* dS = P + e;
  const float scale = 1.0f / sqrtf((float)hd);
  const long long qkv_b1 = (long long)L * 3 * D;
  const long long pb1 = (long long)H * L * Lp, pb2 = (long long)L * Lp;
  S4GemmParams g{};
  int rc;
  // recompute P from the saved lse
  base_params(g, dtype, backend);
  g.a = qkv; g.b = qkv + D; g.c = S; g.c_dtype = S4_F32;
  g.M = L; g.N = L; g.K = hd; g.nb1 = B; g.nb2 = H;
  g.a_sm = 3 * D; g.a_sk = 1; g.a_b1 = qkv_b1; g.a_b2 = hd;
  g.b_sk = 1; g.b_sn = 3 * D; g.b_b1 = qkv_b1; g.b_b2 = hd;
  g.c_sm = Lp; g.c_b1 = pb1; g.c_b2 = pb2; g.alpha = scale;
  if ((rc = s4_gemm(&g, st))) return rc;
  const long long rows = (long long)B * H * L;
  const int blocks = (int)((rows + 7) / 8);
  softmax_bias_kernel<T><<<blocks, 256, 0, st>>>(S, P, u0, gate, w, nullptr, lse, H, L, Lp, rows);
  if ((rc = s4_check_launch("softmax_bias(recompute)"))) return rc;
  // dV[k,:] = sum_q P[q,k] dO[q,:]
  base_params(g, dtype, backend);
  g.a = P; g.b = dout; g.c = dqkv + 2 * D;
  g.M = L; g.N = hd; g.K = L; g.nb1 = B; g.nb2 = H;
  g.a_sm = 1; g.a_sk = Lp; g.a_b1 = pb1; g.a_b2 = pb2;
  g.b_sk = D; g.b_sn = 1; g.b_b1 = (long long)L * D; g.b_b2 = hd;
  g.c_sm = 3 * D; g.c_b1 = qkv_b1; g.c_b2 = hd;
  if ((rc = s4_gemm(&g, st))) return rc;
  // dP = dO V^T   (fp32, reuses the S buffer)
  base_params(g, dtype, backend);
  g.a = dout; g.b = qkv + 2 * D; g.c = S; g.c_dtype = S4_F32;
  g.M = L; g.N = L; g.K = hd; g.nb1 = B; g.nb2 = H;
  g.a_sm = D; g.a_sk = 1; g.a_b1 = (long long)L * D; g.a_b2 = hd;
  g.b_sk = 1; g.b_sn = 3 * D; g.b_b1 = qkv_b1; g.b_b2 = hd;
  g.c_sm = Lp; g.c_b1 = pb1; g.c_b2 = pb2;
  if ((rc = s4_gemm(&g, st))) return rc;
  softmax_bwd_kernel<T><<<blocks, 256, 0, st>>>(S, P, dS, L, Lp, rows);
  if ((rc = s4_check_launch("softmax_bwd"))) return rc;
  // dQ = scale * dS K
  base_params(g, dtype, backend);
  g.a = dS; g.b = qkv + D; g.c = dqkv;
  g.M = L; g.N = hd; g.K = L; g.nb1 = B; g.nb2 = H;
  g.a_sm = Lp; g.a_sk = 1; g.a_b1 = pb1; g.a_b2 = pb2;
  g.b_sk = 3 * D; g.b_sn = 1; g.b_b1 = qkv_b1; g.b_b2 = hd;
  g.c_sm = 3 * D; g.c_b1 = qkv_b1; g.c_b2 = hd; g.alpha = scale;
  if ((rc = s4_gemm(&g, st))) return rc;
  // dK = scale * dS^T Q
  base_params(g, dtype, backend);
  g.a = dS; g.b = qkv; g.c = dqkv + D;
  g.M = L; g.N = hd; g.K = L; g.nb1 = B; g.nb2 = H;
  g.a_sm = 1; g.a_sk = Lp; g.a_b1 = pb1; g.a_b2 = pb2;
  g.b_sk = 3 * D; g.b_sn = 1; g.b_b1 = qkv_b1; g.b_b2 = hd;
  g.c_sm = 3 * D; g.c_b1 = qkv_b1; g.c_b2 = hd; g.alpha = scale;
  return s4_gemm(&g, st);
}

int s4_attention_tc_fwd(const void* qkv, const float* u0, const float* gate, float w, void* out,
                        float* lse, int B, int H, int L, int hd, cudaStream_t st);
int s4_attention_tc_bwd(const void* dout, const void* qkv, const void* out, const float* lse,
                        const float* u0, const float* gate, float w, void* dqkv, void* ws,
                        size_t ws_bytes, int B, int H, int L, int hd, cudaStream_t st);
extern "C" int s4_attention_fwd(const void* qkv, const float* u0, const float* gate,
                                float bias_weight, void* out, float* lse, void* workspace,
                                size_t ws_bytes, int B, int H, int L, int hd, int dtype,
                                int backend, cudaStream_t stream) {
  S4ProfScope prof_("attention_fwd", 0.0, 1, stream);
  if (B * H * L == 0) return S4_OK;
  if (backend != S4_BACKEND_SIMT && s4_attention_tc_fwd_supported(B, H, L, hd, dtype))
    return s4_attention_tc_fwd(qkv, u0, gate, bias_weight, out, lse, B, H, L, hd, stream);
  S4_REQUIRE(backend != S4_BACKEND_TC, "attention: fused tcgen05 path does not support this shape");
  S4_REQUIRE(ws_bytes >= composed_workspace(B, H, L, dtype), "attention_fwd: workspace too small");
  if (dtype == S4_BF16)
    return attention_fwd_t<__nv_bfloat16>((const __nv_bfloat16*)qkv, u0, gate, bias_weight,
                                          (__nv_bfloat16*)out, lse, workspace, B, H, L, hd, dtype,
                                          backend, stream);
  return attention_fwd_t<float>((const float*)qkv, u0, gate, bias_weight, (float*)out, lse,
                                workspace, B, H, L, hd, dtype, backend, stream);
}

extern "C" int s4_attention_bwd(const void* dout, const void* qkv, const void* out,
                                const float* lse, const float* u0, const float* gate,
                                float bias_weight, void* dqkv, void* workspace, size_t ws_bytes,
                                int B, int H, int L, int hd, int dtype, int backend,
                                cudaStream_t stream) {
  S4ProfScope prof_("attention_bwd", 0.0, 1, stream);
  if (B * H * L == 0) return S4_OK;
  if (backend != S4_BACKEND_SIMT && s4_attention_tc_bwd_supported(B, H, L, hd, dtype))
    return s4_attention_tc_bwd(dout, qkv, out, lse, u0, gate, bias_weight, dqkv, workspace,
                               ws_bytes, B, H, L, hd, stream);
  S4_REQUIRE(backend != S4_BACKEND_TC, "attention: fused tcgen05 path does not support this shape");
  S4_REQUIRE(ws_bytes >= composed_workspace(B, H, L, dtype), "attention_bwd: workspace too small");
  if (dtype == S4_BF16)
    return attention_bwd_t<__nv_bfloat16>((const __nv_bfloat16*)dout, (const __nv_bfloat16*)qkv,
                                          lse, u0, gate, bias_weight, (__nv_bfloat16*)dqkv,
                                          workspace, B, H, L, hd, dtype, backend, stream);
  return attention_bwd_t<float>((const float*)dout, (const float*)qkv, lse, u0, gate, bias_weight,
                                (float*)dqkv, workspace, B, H, L, hd, dtype, backend, stream);
}
