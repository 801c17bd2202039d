// Generic strided/batched GEMM on CUDA cores with fp32 accumulation.
//
// This is the VALIDATION path (fp32 activations, parity gate 1e-3) and the safety net under the
// tcgen05 kernels in gemm_tc.cu: same epilogue contract, arbitrary strides, so every contraction
// of the train step (linear fwd/dgrad/wgrad, QK^T, PV and their backward forms; reference
// vit.py:113-127 through mmcv MultiheadAttention/FFN) can be expressed with it.
//
//   C[z][m,n] = epi( alpha * sum_k A[z][m,k] * B[z][k,n] ),  z = (z1, z2)
//   epi(v): v += bias[n]; v *= gelu'(aux[m,n]); pre[m,n] = v; v = gelu(v); v += res[m,n];
//           v += C_old[m,n]  (each step optional)
#include "common.cuh"
#include "gemm_params.h"

#define TM 64
#define TN 64
#define TK 16

template <typename T, typename TC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(S4GemmParams p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int z = blockIdx.z;
  const int z1 = z / p.nb2, z2 = z % p.nb2;
  const T* A = (const T*)p.a + (size_t)z1 * p.a_b1 + (size_t)z2 * p.a_b2;
  const T* B = (const T*)p.b + (size_t)z1 * p.b_b1 + (size_t)z2 * p.b_b2;
  const size_t coff = (size_t)z1 * p.c_b1 + (size_t)z2 * p.c_b2;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += TK) {
    // A tile: TM x TK.  Choose the thread->element map so the unit-stride dim is fastest.
    for (int e = tid; e < TM * TK; e += 256) {
      int mm, kk;
      if (p.a_sk == 1) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < p.M && k < p.K) v = to_f32<T>(A[(size_t)m * p.a_sm + (size_t)k * p.a_sk]);
      As[kk][mm] = v;
    }
    for (int e = tid; e < TN * TK; e += 256) {
      int nn, kk;
      if (p.b_sk == 1) { kk = e % TK; nn = e / TK; } else { nn = e % TN; kk = e / TN; }
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < p.N && k < p.K) v = to_f32<T>(B[(size_t)k * p.b_sk + (size_t)n * p.b_sn]);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  TC* C = (TC*)p.c + coff;
  const T* aux = p.aux ? (const T*)p.aux + coff : nullptr;
  const T* res = p.res ? (const T*)p.res + coff : nullptr;
  T* pre = p.pre ? (T*)p.pre + coff : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      const size_t o = (size_t)m * p.c_sm + n;
      float v = p.alpha * acc[i][j];
      if (p.bias) v += p.bias[n];
      if (aux) v *= gelu_erf_grad(to_f32<T>(aux[o]));
      if (pre) pre[o] = from_f32<T>(v);
      if (p.act == S4_ACT_GELU) v = gelu_erf(v);
      if (res) v += to_f32<T>(res[o]);
      if (p.accumulate) v += to_f32<TC>(C[o]);
      C[o] = from_f32<TC>(v);
    }
  }
}

int s4_gemm_simt_launch(const S4GemmParams& p, cudaStream_t stream) {
  if (p.M == 0 || p.N == 0 || p.nb1 * p.nb2 == 0) return S4_OK;
  dim3 grid((p.N + TN - 1) / TN, (p.M + TM - 1) / TM, p.nb1 * p.nb2);
  S4_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm_simt: grid too large");
  S4ProfScope prof("gemm_simt", 2.0 * p.M * p.N * (double)p.K * p.nb1 * p.nb2, 0, stream);
  if (p.dtype == S4_BF16) {
    if (p.c_dtype == S4_F32)
      gemm_simt_kernel<__nv_bfloat16, float><<<grid, 256, 0, stream>>>(p);
    else
      gemm_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>(p);
  } else {
    S4_REQUIRE(p.c_dtype == S4_F32, "gemm_simt: f32 inputs need f32 output");
    gemm_simt_kernel<float, float><<<grid, 256, 0, stream>>>(p);
  }
  return s4_check_launch("gemm_simt");
}
