// LayerNorm forward/backward over the last dim (D <= 1024, D % 8 == 0), fp32 statistics.
// One warp per row, 16-byte vector loads.  An optional row map gathers source rows, which folds
// the backbone feature tap (drop the cls token, reference vit.py:556-562) and the PatchMix
// feature un-shuffle (decode_head.py:186-212) into the SETR head's LayerNorm
// (setr_up_head.py:96-103).  Reference for the op itself: vit.py:119-120 (eps 1e-6).
#include "common.cuh"

#define LN_MAX_VPL 8  // vectors per lane: D <= 32 lanes * 8 vec * 4 (f32) = 1024

template <typename T, int VPL>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const T* __restrict__ x, const int* __restrict__ row_map,
              const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ y,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int D,
              float eps) {
  constexpr int VN = Vec16<T>::N;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int src = row_map ? row_map[warp] : warp;
  const T* xr = x + (size_t)src * D;
  const int nvec = D / VN;
  Vec16<T> v[VPL];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
      v[k].load(xr + vi * VN);
#pragma unroll
      for (int e = 0; e < VN; ++e) sum += v[k].get(e);
    }
  }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const float d = v[k].get(e) - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
  if (lane == 0) {
    if (mean_out) mean_out[warp] = mean;
    if (rstd_out) rstd_out[warp] = rstd;
  }
  T* yr = y + (size_t)warp * D;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
      Vec16<T> o;
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const int c = vi * VN + e;
        o.set(e, (v[k].get(e) - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c));
      }
      o.store(yr + vi * VN);
    }
  }
}

// dx (written at the mapped source row of a pre-zeroed buffer when row_map != null),
// dgamma/dbeta accumulated with atomics (buffers must be zero-initialised or hold a running sum).
template <typename T, int VPL>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const int* __restrict__ row_map,
              const float* __restrict__ gamma, const float* __restrict__ mean,
              const float* __restrict__ rstd, const T* __restrict__ dres, T* __restrict__ dx,
              float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int D) {
  constexpr int VN = Vec16<T>::N;
  extern __shared__ float sh[];  // [2][D]
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nvec = D / VN;
  float ag[VPL][VN], ab[VPL][VN];
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int e = 0; e < VN; ++e) { ag[k][e] = 0.f; ab[k][e] = 0.f; }
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const int src = row_map ? row_map[row] : row;
    const T* xr = x + (size_t)src * D;
    const T* gr = dy + (size_t)row * D;
    const float mu = mean[row], rs = rstd[row];
    Vec16<T> xv[VPL], gv[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (vi < nvec) {
        xv[k].load(xr + vi * VN);
        gv[k].load(gr + vi * VN);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const float xh = (xv[k].get(e) - mu) * rs;
          const float go = gv[k].get(e);
          const float g = go * __ldg(gamma + vi * VN + e);
          s1 += g;
          s2 += g * xh;
          ag[k][e] += go * xh;
          ab[k][e] += go;
        }
      }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
    T* dr = dx + (size_t)src * D;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (vi < nvec) {
        Vec16<T> o, rv;
        if (dres) rv.load(dres + (size_t)src * D + vi * VN);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const float xh = (xv[k].get(e) - mu) * rs;
          const float g = gv[k].get(e) * __ldg(gamma + vi * VN + e);
          float val = rs * (g - s1 - xh * s2);
          if (dres) val += rv.get(e);
          o.set(e, val);
        }
        o.store(dr + vi * VN);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        atomicAdd(&sh[vi * VN + e], ag[k][e]);
        atomicAdd(&sh[D + vi * VN + e], ab[k][e]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, sh[i]);
    atomicAdd(dbeta + i, sh[D + i]);
  }
}

#define LN_DISPATCH_VPL(vpl, ...)                     \
  do {                                                \
    if (vpl <= 1) { constexpr int VPL = 1; __VA_ARGS__; }      \
    else if (vpl <= 2) { constexpr int VPL = 2; __VA_ARGS__; } \
    else if (vpl <= 3) { constexpr int VPL = 3; __VA_ARGS__; } \
    else if (vpl <= 4) { constexpr int VPL = 4; __VA_ARGS__; } \
    else if (vpl <= 6) { constexpr int VPL = 6; __VA_ARGS__; } \
    else { constexpr int VPL = 8; __VA_ARGS__; }               \
  } while (0)

extern "C" int s4_layernorm_fwd(const void* x, const int* row_map, const float* gamma,
                                const float* beta, void* y, float* mean, float* rstd, int rows,
                                int D, float eps, int dtype, cudaStream_t stream) {
  S4ProfScope prof_("layernorm_fwd", 0.0, 1, stream);
  const int vn = dtype == S4_BF16 ? 8 : 4;
  S4_REQUIRE(D % vn == 0 && D / vn <= 32 * LN_MAX_VPL, "layernorm: unsupported D=%d", D);
  if (rows == 0) return S4_OK;
  const int blocks = (rows + 7) / 8;
  const int vpl = (D / vn + 31) / 32;
  if (dtype == S4_BF16) {
    LN_DISPATCH_VPL(vpl, (ln_fwd_kernel<__nv_bfloat16, VPL><<<blocks, 256, 0, stream>>>(
        (const __nv_bfloat16*)x, row_map, gamma, beta, (__nv_bfloat16*)y, mean, rstd, rows, D, eps)));
  } else {
    LN_DISPATCH_VPL(vpl, (ln_fwd_kernel<float, VPL><<<blocks, 256, 0, stream>>>(
        (const float*)x, row_map, gamma, beta, (float*)y, mean, rstd, rows, D, eps)));
  }
  return s4_check_launch("layernorm_fwd");
}

extern "C" int s4_layernorm_bwd(const void* dy, const void* x, const int* row_map,
                                const float* gamma, const float* mean, const float* rstd,
                                const void* dres, void* dx, float* dgamma, float* dbeta, int rows,
                                int D, int dtype, cudaStream_t stream) {
  S4ProfScope prof_("layernorm_bwd", 0.0, 1, stream);
  const int vn = dtype == S4_BF16 ? 8 : 4;
  S4_REQUIRE(D % vn == 0 && D / vn <= 32 * LN_MAX_VPL, "layernorm: unsupported D=%d", D);
  if (rows == 0) return S4_OK;
  int blocks = (rows + 7) / 8;
  const int cap = s4_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  const size_t smem = 2 * (size_t)D * sizeof(float);
  const int vpl = (D / vn + 31) / 32;
  if (dtype == S4_BF16) {
    LN_DISPATCH_VPL(vpl, (ln_bwd_kernel<__nv_bfloat16, VPL><<<blocks, 256, smem, stream>>>(
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, row_map, gamma, mean, rstd,
        (const __nv_bfloat16*)dres, (__nv_bfloat16*)dx, dgamma, dbeta, rows, D)));
  } else {
    LN_DISPATCH_VPL(vpl, (ln_bwd_kernel<float, VPL><<<blocks, 256, smem, stream>>>(
        (const float*)dy, (const float*)x, row_map, gamma, mean, rstd, (const float*)dres,
        (float*)dx, dgamma, dbeta, rows, D)));
  }
  return s4_check_launch("layernorm_bwd");
}
