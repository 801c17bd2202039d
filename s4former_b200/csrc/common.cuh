// Shared helpers for the s4former_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define S4_OK 0
#define S4_ERR_ARG (-1)
#define S4_ERR_CUDA (-2)
#define S4_ERR_UNSUPPORTED (-3)

// dtype codes used across the C ABI
#define S4_F32 0
#define S4_BF16 1

void s4_set_error(const char* fmt, ...);
int s4_check_launch(const char* what);   // also counts one kernel launch
void s4_count_launches(int n);           // extra launches behind a single check

// ---- optional per-kernel-family device timing (s4_prof_* in include/s4former.h) ---------------
// When enabled, a scope records a CUDA event pair on the launching stream around the launches
// it brackets and books `work` (FLOPs or bytes: the ALGORITHMIC figure, not the executed one)
// under `name`.  Disabled (the default) it costs one relaxed atomic load.
bool s4_prof_on();
void s4_prof_begin(const char* name, double work, int unit /*0 = flop, 1 = byte*/, cudaStream_t st,
                   void** cookie);
void s4_prof_end(void* cookie, cudaStream_t st);
struct S4ProfScope {
  void* cookie = nullptr;
  cudaStream_t st;
  S4ProfScope(const char* name, double work, int unit, cudaStream_t s) : st(s) {
    if (s4_prof_on()) s4_prof_begin(name, work, unit, s, &cookie);
  }
  ~S4ProfScope() {
    if (cookie) s4_prof_end(cookie, st);
  }
};

#define S4_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      s4_set_error(__VA_ARGS__);         \
      return S4_ERR_ARG;                 \
    }                                    \
  } while (0)

static inline int s4_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` must hold >= 32 floats. Result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// exact (erf) GELU, matching torch.nn.GELU() default
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// GELU with Phi(-a) = 2^(-g(a)), g a degree-8 polynomial fit of -log2(Phi(-a)) on [0, 6]
// (max |error| of GELU 6e-7, relative error of the negative tail 5e-6: one MUFU instead of erff).
// Used by the bf16 tensor-core epilogues; the fp32 validation path keeps erff.
__device__ __forceinline__ float norm_cdf_fast(float x) {
  const float a = fminf(fabsf(x), 6.0f);
  float g = 7.981907579335257e-09f;
  g = fmaf(g, a, 1.698060486887698e-06f);
  g = fmaf(g, a, -6.081363608245738e-05f);
  g = fmaf(g, a, 0.0009293854236602783f);
  g = fmaf(g, a, -0.00851279217749834f);
  g = fmaf(g, a, 0.05397995561361313f);
  g = fmaf(g, a, 0.4584403336048126f);
  g = fmaf(g, a, 1.1512603759765625f);
  g = fmaf(g, a, 0.9999947547912598f);
  float pm;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pm) : "f"(-g));
  return x < 0.f ? pm : 1.f - pm;
}
__device__ __forceinline__ float gelu_fast(float x) { return x * norm_cdf_fast(x); }
// GELU for bf16 OUTPUTS: 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) with the hardware
// tanh.approx (one MUFU): 6 instructions instead of 15.  |error| vs the erf form <= 5e-4 absolute,
// an eighth of a bf16 ulp at |gelu| ~ 1; the fc1 epilogue (K = 768: a tile's MMAs take only ~6 100
// clk) was issue-bound on the degree-8 form.  The fp32 validation path keeps erff.
__device__ __forceinline__ float gelu_tanh_bf16(float x) {
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  float pdf;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pdf) : "f"(fmaf(x * x, -0.7213475204444817f, -1.3257480647361595f)));
  return fmaf(x, pdf, norm_cdf_fast(x));
}

// 16-byte vector of 8 bf16 / 4 f32 helpers
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ __forceinline__ float get(int i) const { return (&v.x)[i]; }
  __device__ __forceinline__ void set(int i, float f) { (&v.x)[i] = f; }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ float get(int i) const {
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&v);
    return __bfloat162float(h[i]);
  }
  __device__ __forceinline__ void set(int i, float f) {
    __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(&v);
    h[i] = __float2bfloat16_rn(f);
  }
};
