// SETR-PUP decode head kernels (NHWC activations).
// Reference: mmseg/models/decode_heads/setr_up_head.py:49-77,92-111 (ConvModule = conv3x3
// without bias -> SyncBN -> ReLU, then Upsample(bilinear, align_corners=False), n times, then the
// 1x1 conv_seg of decode_head.py:107-111,311-316); mmseg/ops/wrappers.py:30-51.
//
// Layout decisions (B200-first):
//  * BN + ReLU are never materialised on their own: they are folded into the load of the
//    consumer (bilinear upsample, or the 1x1 classifier of the last stage).
//  * The last stage applies conv_seg BEFORE the final upsample (both are linear and bilinear
//    weights sum to 1, so bias commutes): the [B,256,512,512] tensor of the reference
//    (268 MB/img fp32) never exists; a 21-channel map is upsampled instead.
//  * CUDA-core 3x3 convolutions here are the fp32 validation path; gemm_tc.cu holds the
//    tcgen05 implicit-GEMM version.
#include <algorithm>

#include "common.cuh"

// head_cls.cu: bf16 tensor-core kernels of the last stage
bool s4_cls_tc_supported(int C, int NC, int dtype);
int s4_cls_fwd_tc(const void* y, const float* scale, const float* shift, const float* w, const float* bias,
                  float* z, long long rows, int C, int NC, cudaStream_t st);
int s4_cls_bwd_reduce_tc(const void* dz16, const void* y, const float* scale, const float* shift,
                         const float* mean, const float* invstd, const float* w, float* dw, float* dbias,
                         float* dsum, float* ddot, long long rows, int C, int NC, cudaStream_t st);
int s4_cls_bwd_apply_tc(const void* dz16, const void* y, const float* scale, const float* shift,
                        const float* mean, const float* invstd, const float* gamma, const float* w,
                        const float* dsum, const float* ddot, double count, void* dy, long long rows,
                        int C, int NC, cudaStream_t st);
int s4_cls_upsample_fwd(const float* z, float* logits, int B, int H, int W, int NC, int s, cudaStream_t st);
int s4_cls_upsample_bwd(const float* dlogits, void* dz16, int B, int H, int W, int NC, int s, cudaStream_t st);
#include "gemm_params.h"

// head_stream.cu: second-generation streaming kernels (bf16); false = shape outside the fast path
bool s4_stream_upsample_fwd(const void* x, const float* scale, const float* shift, void* out, int B, int H,
                            int W, int C, int s, cudaStream_t st);
bool s4_stream_upsample_bwd(const void* dout, const void* x, const float* scale, const float* shift,
                            const float* mean, const float* invstd, void* dact, float* dsum, float* ddot,
                            int B, int H, int W, int C, int s, cudaStream_t st);
bool s4_stream_bn_bwd_apply(const void* dact, const void* x, const float* gamma, const float* mean,
                            const float* invstd, const float* dsum, const float* ddot, double count,
                            void* dy, long long rows, int C, cudaStream_t st);
bool s4_stream_pack_conv_weight(const float* w, void* wf, void* wd, int Cin, int Cout, cudaStream_t st);

// ------------------------------------------------------------------------------------------
// 3x3 convolution as implicit GEMM on CUDA cores (fp32 accumulate)
//   fwd  : M = B*H*W pixels, N = Cout, K = 9*Cin ; A gathered from x, B = w_packed [N][K]
//   wgrad: M = Cout, N = 9*Cin, K = pixels ; A = dy^T, B gathered from x ; fp32 accumulate
// ------------------------------------------------------------------------------------------
#define CT 64
#define CK 16

template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_simt_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y, int B,
                    int H, int W, int Cin, int Cout) {
  __shared__ float As[CK][CT + 4];
  __shared__ float Bs[CK][CT + 4];
  const long long M = (long long)B * H * W;
  const int K = 9 * Cin;
  const long long m0 = (long long)blockIdx.y * CT;
  const int n0 = blockIdx.x * CT;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += CK) {
    for (int e = tid; e < CT * CK; e += 256) {
      const int kk = e % CK, mm = e / CK;
      const long long m = m0 + mm;
      const int k = k0 + kk;
      float v = 0.f;
      if (m < M && k < K) {
        const int tap = k / Cin, ci = k % Cin;
        const int px = (int)(m % W), py = (int)((m / W) % H);
        const long long b = m / ((long long)W * H);
        const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
          v = to_f32<T>(x[((b * H + yy) * W + xx) * Cin + ci]);
      }
      As[kk][mm] = v;
    }
    for (int e = tid; e < CT * CK; e += 256) {
      const int kk = e % CK, nn = e / CK;
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < Cout && k < K) v = to_f32<T>(w[(size_t)n * K + k]);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CK; ++kk) {
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < Cout) y[m * Cout + n] = from_f32<T>(acc[i][j]);
    }
  }
}

// dw[co][ci][ky][kx] += sum_pix dy[pix][co] * x[pix + tap][ci];  blockIdx.z splits the pixels
template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_wgrad_simt_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                          int B, int H, int W, int Cin, int Cout, long long pix_per_split) {
  __shared__ float As[CK][CT + 4];
  __shared__ float Bs[CK][CT + 4];
  const long long P = (long long)B * H * W;
  const int N = 9 * Cin;
  const int m0 = blockIdx.y * CT, n0 = blockIdx.x * CT;
  const long long p_begin = (long long)blockIdx.z * pix_per_split;
  const long long p_end = min(P, p_begin + pix_per_split);
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[4][4] = {};
  for (long long k0 = p_begin; k0 < p_end; k0 += CK) {
    for (int e = tid; e < CT * CK; e += 256) {
      const int mm = e % CT, kk = e / CT;
      const int m = m0 + mm;
      const long long pix = k0 + kk;
      float v = 0.f;
      if (m < Cout && pix < p_end) v = to_f32<T>(dy[pix * Cout + m]);
      As[kk][mm] = v;
    }
    for (int e = tid; e < CT * CK; e += 256) {
      const int nn = e % CT, kk = e / CT;
      const int n = n0 + nn;
      const long long pix = k0 + kk;
      float v = 0.f;
      if (n < N && pix < p_end) {
        const int tap = n / Cin, ci = n % Cin;
        const int px = (int)(pix % W), py = (int)((pix / W) % H);
        const long long b = pix / ((long long)W * H);
        const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
          v = to_f32<T>(x[((b * H + yy) * W + xx) * Cin + ci]);
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CK; ++kk) {
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = m0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const int tap = n / Cin, ci = n % Cin;
      atomicAdd(dw + ((size_t)co * Cin + ci) * 9 + tap, acc[i][j]);
    }
  }
}

int s4_conv3x3_tc(const void* x, const void* w_packed, void* y, int B, int H, int W, int Cin,
                  int Cout, cudaStream_t stream);
bool s4_conv3x3_tc_supported(int B, int H, int W, int Cin, int Cout, int dtype);
int s4_conv3x3_tc_stats(const void* x, const void* w_packed, void* y, float* sum, float* sumsq, int B,
                        int H, int W, int Cin, int Cout, cudaStream_t stream);
bool s4_conv3x3_tc_stats_supported(int B, int H, int W, int Cin, int Cout, int dtype);
int s4_conv3x3_wgrad_tc(const void* x, const void* dy, float* dw, int B, int H, int W, int Cin,
                        int Cout, cudaStream_t stream);
bool s4_conv3x3_wgrad_tc_supported(int B, int H, int W, int Cin, int Cout, int dtype);

static int conv3x3_any(const void* x, const void* w, void* y, int B, int H, int W, int Cin,
                       int Cout, int dtype, int backend, cudaStream_t stream) {
  if ((long long)B * H * W == 0) return S4_OK;
  if (backend != S4_BACKEND_SIMT && s4_conv3x3_tc_supported(B, H, W, Cin, Cout, dtype))
    return s4_conv3x3_tc(x, w, y, B, H, W, Cin, Cout, stream);
  S4_REQUIRE(backend != S4_BACKEND_TC, "conv3x3: tcgen05 path does not support this shape");
  const long long M = (long long)B * H * W;
  dim3 grid((Cout + CT - 1) / CT, (unsigned)((M + CT - 1) / CT));
  if (dtype == S4_BF16)
    conv3x3_simt_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (__nv_bfloat16*)y, B, H, W, Cin, Cout);
  else
    conv3x3_simt_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, (const float*)w,
                                                         (float*)y, B, H, W, Cin, Cout);
  return s4_check_launch("conv3x3_simt");
}

extern "C" int s4_conv3x3_fwd(const void* x, const void* w_packed, void* y, int B, int H, int W,
                              int Cin, int Cout, int dtype, int backend, cudaStream_t stream) {
  S4ProfScope prof_("conv3x3_fwd", 0.0, 1, stream);
  return conv3x3_any(x, w_packed, y, B, H, W, Cin, Cout, dtype, backend, stream);
}

extern "C" int s4_colsum(const void* x, float* sum, float* sumsq, long long rows, int cols, int dtype,
                         cudaStream_t stream);

extern "C" int s4_conv3x3_fwd_stats(const void* x, const void* w_packed, void* y, float* sum,
                                    float* sumsq, int B, int H, int W, int Cin, int Cout, int dtype,
                                    int backend, cudaStream_t stream) {
  if ((long long)B * H * W == 0) return S4_OK;
  if (backend != S4_BACKEND_SIMT && s4_conv3x3_tc_stats_supported(B, H, W, Cin, Cout, dtype)) {
    S4ProfScope prof_("conv3x3_fwd", 0.0, 1, stream);
    return s4_conv3x3_tc_stats(x, w_packed, y, sum, sumsq, B, H, W, Cin, Cout, stream);
  }
  int rc = s4_conv3x3_fwd(x, w_packed, y, B, H, W, Cin, Cout, dtype, backend, stream);
  if (rc) return rc;
  return s4_colsum(y, sum, sumsq, (long long)B * H * W, Cout, dtype, stream);
}

extern "C" int s4_conv3x3_dgrad(const void* dy, const void* w_dgrad, void* dx, int B, int H, int W,
                                int Cin, int Cout, int dtype, int backend, cudaStream_t stream) {
  S4ProfScope prof_("conv3x3_dgrad", 0.0, 1, stream);
  // dgrad of a stride-1 pad-1 3x3 conv is a 3x3 conv of dy with flipped, transposed weights
  return conv3x3_any(dy, w_dgrad, dx, B, H, W, Cout, Cin, dtype, backend, stream);
}

extern "C" int s4_conv3x3_wgrad(const void* x, const void* dy, float* dw, int B, int H, int W,
                                int Cin, int Cout, int dtype, int backend, cudaStream_t stream) {
  S4ProfScope prof_("conv3x3_wgrad", 0.0, 1, stream);
  const long long P = (long long)B * H * W;
  if (P == 0) return S4_OK;
  if (backend != S4_BACKEND_SIMT && s4_conv3x3_wgrad_tc_supported(B, H, W, Cin, Cout, dtype))
    return s4_conv3x3_wgrad_tc(x, dy, dw, B, H, W, Cin, Cout, stream);
  S4_REQUIRE(backend != S4_BACKEND_TC, "conv3x3_wgrad: tcgen05 path does not support this shape");
  const int gx = (9 * Cin + CT - 1) / CT, gy = (Cout + CT - 1) / CT;
  int splits = (s4_num_sms() * 4 + gx * gy - 1) / (gx * gy);
  if (splits < 1) splits = 1;
  long long pps = (P + splits - 1) / splits;
  pps = ((pps + CK - 1) / CK) * CK;
  splits = (int)((P + pps - 1) / pps);
  dim3 grid(gx, gy, splits);
  if (dtype == S4_BF16)
    conv3x3_wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw, B, H, W, Cin, Cout, pps);
  else
    conv3x3_wgrad_simt_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, (const float*)dy,
                                                               dw, B, H, W, Cin, Cout, pps);
  return s4_check_launch("conv3x3_wgrad_simt");
}

// w [Cout][Cin][3][3] f32 -> fwd [Cout][tap*Cin+ci], dgrad [Cin][(8-tap)*Cout+co]
template <typename T>
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, T* __restrict__ wf,
                                        T* __restrict__ wd, int Cin, int Cout) {
  const int total = Cout * Cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, ci = (i / 9) % Cin, co = i / (9 * Cin);
    const T v = from_f32<T>(w[i]);
    if (wf) wf[(size_t)co * 9 * Cin + tap * Cin + ci] = v;
    if (wd) wd[(size_t)ci * 9 * Cout + (8 - tap) * Cout + co] = v;
  }
}

extern "C" int s4_pack_conv3x3_weight(const float* w, void* w_fwd, void* w_dgrad, int Cin,
                                      int Cout, int dtype, cudaStream_t stream) {
  S4ProfScope prof_("pack_conv3x3_weight", 0.0, 1, stream);
  const int total = Cout * Cin * 9;
  if (total == 0) return S4_OK;
  if (dtype == S4_BF16 && s4_stream_pack_conv_weight(w, w_fwd, w_dgrad, Cin, Cout, stream))
    return s4_check_launch("pack_conv3x3_weight");
  const int grid = min((total + 255) / 256, s4_num_sms() * 8);
  if (dtype == S4_BF16)
    pack_conv_weight_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(w, (__nv_bfloat16*)w_fwd, (__nv_bfloat16*)w_dgrad, Cin, Cout);
  else
    pack_conv_weight_kernel<float><<<grid, 256, 0, stream>>>(w, (float*)w_fwd, (float*)w_dgrad, Cin, Cout);
  return s4_check_launch("pack_conv3x3_weight");
}

// ------------------------------------------------------------------------------------------
// BatchNorm statistics
// ------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq,
                                   double count, float eps, float momentum,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ rmean, float* __restrict__ rvar,
                                   long long* __restrict__ nbt, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (c == 0 && nbt) *nbt += 1;        // BatchNorm's num_batches_tracked, no extra launch
  const double mu = (double)sum[c] / count;
  double var = (double)sumsq[c] / count - mu * mu;
  if (var < 0) var = 0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)mu;
  invstd[c] = is;
  const float sc = gamma[c] * is;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mu * sc;
  if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mu;
  if (rvar) {
    const double unb = count > 1 ? var * count / (count - 1.0) : var;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
  }
}

extern "C" int s4_bn_finalize(const float* sum, const float* sumsq, double count, float eps,
                              float momentum, const float* gamma, const float* beta, float* mean,
                              float* invstd, float* scale, float* shift, float* running_mean,
                              float* running_var, long long* num_batches_tracked, int C,
                              cudaStream_t stream) {
  S4ProfScope prof_("bn_finalize", 0.0, 1, stream);
  if (C == 0) return S4_OK;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sum, sumsq, count, eps, momentum, gamma,
                                                         beta, mean, invstd, scale, shift,
                                                         running_mean, running_var,
                                                         num_batches_tracked, C);
  return s4_check_launch("bn_finalize");
}

__global__ void bn_eval_affine_kernel(const float* __restrict__ rm, const float* __restrict__ rv,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float eps, float* __restrict__ scale, float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}

extern "C" int s4_bn_eval_affine(const float* running_mean, const float* running_var,
                                 const float* gamma, const float* beta, float eps, float* scale,
                                 float* shift, int C, cudaStream_t stream) {
  S4ProfScope prof_("bn_eval_affine", 0.0, 1, stream);
  if (C == 0) return S4_OK;
  bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, stream>>>(running_mean, running_var, gamma, beta, eps, scale, shift, C);
  return s4_check_launch("bn_eval_affine");
}

// ------------------------------------------------------------------------------------------
// bilinear helpers (align_corners=False, integer scale), matching ATen's upsample_bilinear2d
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int o, int s, int n_in, int& i0, int& i1, float& l) {
  float src = ((float)o + 0.5f) / (float)s - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  i1 = min(i0 + 1, n_in - 1);
  l = src - (float)i0;
}
// weight with which input index i contributes to output index o
__device__ __forceinline__ float bilinear_w(int o, int i, int s, int n_in) {
  int i0, i1;
  float l;
  bilinear_src(o, s, n_in, i0, i1, l);
  return (i == i0 ? 1.f - l : 0.f) + (i == i1 ? l : 0.f);
}

// out[b,oy,ox,c] = bilinear(relu(x*scale+shift));  one thread per (output pixel, 16-byte vector)
template <typename T>
__global__ void __launch_bounds__(256)
bn_relu_upsample_fwd_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                            const float* __restrict__ shift, T* __restrict__ out, int B, int H,
                            int W, int C, int s) {
  constexpr int VN = Vec16<T>::N;
  const int cv = C / VN;
  const int OH = H * s, OW = W * s;
  const size_t total = (size_t)B * OH * OW * cv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    const size_t pix = i / cv;
    const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH);
    const size_t b = pix / ((size_t)OW * OH);
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(oy, s, H, y0, y1, ly);
    bilinear_src(ox, s, W, x0, x1, lx);
    const T* base = x + b * (size_t)H * W * C + (size_t)v * VN;
    Vec16<T> a00, a01, a10, a11, o;
    a00.load(base + ((size_t)y0 * W + x0) * C);
    a01.load(base + ((size_t)y0 * W + x1) * C);
    a10.load(base + ((size_t)y1 * W + x0) * C);
    a11.load(base + ((size_t)y1 * W + x1) * C);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      const float sc = __ldg(scale + v * VN + e), sh = __ldg(shift + v * VN + e);
      const float r00 = fmaxf(fmaf(a00.get(e), sc, sh), 0.f);
      const float r01 = fmaxf(fmaf(a01.get(e), sc, sh), 0.f);
      const float r10 = fmaxf(fmaf(a10.get(e), sc, sh), 0.f);
      const float r11 = fmaxf(fmaf(a11.get(e), sc, sh), 0.f);
      o.set(e, (1.f - ly) * ((1.f - lx) * r00 + lx * r01) + ly * ((1.f - lx) * r10 + lx * r11));
    }
    o.store(out + pix * C + (size_t)v * VN);
  }
}

// Block-structured version (the hot one): a thread owns (input pixel, 16-byte channel vector).
// It loads the 3x3 clamped neighbourhood once, applies BN + ReLU once per loaded element and emits
// the S x S output pixels that interpolate inside it.  With clamped neighbour coordinates the
// border cases need no special weights: a clamped row/column is a copy of the centre one, so the
// interior weights reproduce ATen's clamped source index.  32-bit index math only; a warp covers
// one pixel's 512 contiguous bytes (C = 256, bf16) in every load and store.
template <typename T, int S>
__global__ void __launch_bounds__(256)
bn_relu_upsample_fwd_blk_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                const float* __restrict__ shift, T* __restrict__ out, int B, int H,
                                int W, int C) {
  constexpr int VN = Vec16<T>::N;
  const int cv = C / VN, ppb = 256 / cv;
  const int v = threadIdx.x % cv, pl = threadIdx.x / cv;
  float sc[VN], sh[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) { sc[e] = __ldg(scale + v * VN + e); sh[e] = __ldg(shift + v * VN + e); }
  const int xg = (W + ppb - 1) / ppb;
  const int n_items = B * H * xg;
  const int OW = W * S;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int xgi = item % xg;
    const int r = item / xg;
    const int iy = r % H, b = r / H;
    const int ix = xgi * ppb + pl;
    if (ix >= W) continue;
    const int ys[3] = {max(iy - 1, 0), iy, min(iy + 1, H - 1)};
    const int xs[3] = {max(ix - 1, 0), ix, min(ix + 1, W - 1)};
    const T* base = x + (size_t)b * H * W * C + (size_t)v * VN;
    float a[3][3][VN];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        Vec16<T> t;
        t.load(base + ((size_t)ys[j] * W + xs[i]) * C);
#pragma unroll
        for (int e = 0; e < VN; ++e) a[j][i][e] = fmaxf(fmaf(t.get(e), sc[e], sh[e]), 0.f);
      }
    T* obase = out + (((size_t)b * H * S + (size_t)iy * S) * OW + (size_t)ix * S) * C + (size_t)v * VN;
#pragma unroll
    for (int ry = 0; ry < S; ++ry) {
      // source row = iy + (ry + 0.5)/S - 0.5: rows (iy-1, iy) for the upper half, (iy, iy+1) below
      const int j0 = ry < S / 2 ? 0 : 1;
      const float ly = (ry + 0.5f) / S - 0.5f + (ry < S / 2 ? 1.f : 0.f);
#pragma unroll
      for (int rx = 0; rx < S; ++rx) {
        const int i0 = rx < S / 2 ? 0 : 1;
        const float lx = (rx + 0.5f) / S - 0.5f + (rx < S / 2 ? 1.f : 0.f);
        Vec16<T> o;
#pragma unroll
        for (int e = 0; e < VN; ++e)
          o.set(e, (1.f - ly) * ((1.f - lx) * a[j0][i0][e] + lx * a[j0][i0 + 1][e]) +
                       ly * ((1.f - lx) * a[j0 + 1][i0][e] + lx * a[j0 + 1][i0 + 1][e]));
        o.store(obase + ((size_t)ry * OW + rx) * C);
      }
    }
  }
}

template <typename T>
static bool bn_relu_upsample_fwd_blk(const void* x, const float* scale, const float* shift, void* out,
                                     int B, int H, int W, int C, int s, cudaStream_t stream) {
  constexpr int VN = Vec16<T>::N;
  const int cv = C / VN;
  if ((s != 2 && s != 4) || cv > 256 || 256 % cv) return false;
  const int ppb = 256 / cv;
  const long long items = (long long)B * H * ((W + ppb - 1) / ppb);
  if (items >= (1ll << 31) || (long long)B * H * W * s * s >= (1ll << 31)) return false;
  const int grid = (int)std::min<long long>(items, (long long)s4_num_sms() * 16);
  if (s == 2)
    bn_relu_upsample_fwd_blk_kernel<T, 2><<<grid, 256, 0, stream>>>((const T*)x, scale, shift, (T*)out, B, H, W, C);
  else
    bn_relu_upsample_fwd_blk_kernel<T, 4><<<grid, 256, 0, stream>>>((const T*)x, scale, shift, (T*)out, B, H, W, C);
  return true;
}

extern "C" int s4_bn_relu_upsample_fwd(const void* x, const float* scale, const float* shift,
                                       void* out, int B, int H, int W, int C, int s, int dtype,
                                       cudaStream_t stream) {
  S4ProfScope prof_("bn_relu_upsample_fwd", 0.0, 1, stream);
  const int vn = dtype == S4_BF16 ? 8 : 4;
  S4_REQUIRE(C % vn == 0 && s >= 1, "bn_relu_upsample: C=%d must be a multiple of %d", C, vn);
  const size_t total = (size_t)B * H * s * W * s * (C / vn);
  if (total == 0) return S4_OK;
  if (dtype == S4_BF16 && s4_stream_upsample_fwd(x, scale, shift, out, B, H, W, C, s, stream))
    return s4_check_launch("bn_relu_upsample_fwd");
  const bool done = dtype == S4_BF16
                        ? bn_relu_upsample_fwd_blk<__nv_bfloat16>(x, scale, shift, out, B, H, W, C, s, stream)
                        : bn_relu_upsample_fwd_blk<float>(x, scale, shift, out, B, H, W, C, s, stream);
  if (done) return s4_check_launch("bn_relu_upsample_fwd");
  const int grid = (int)min((total + 255) / 256, (size_t)s4_num_sms() * 32);
  if (dtype == S4_BF16)
    bn_relu_upsample_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)x, scale, shift, (__nv_bfloat16*)out, B, H, W, C, s);
  else
    bn_relu_upsample_fwd_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, scale, shift,
                                                                 (float*)out, B, H, W, C, s);
  return s4_check_launch("bn_relu_upsample_fwd");
}

// dact[b,iy,ix,c] = [bn(x)>0] * sum_{oy,ox} wy*wx*dout[b,oy,ox,c];  per-channel sums of dact and
// dact*xhat are accumulated for the BatchNorm backward.  One thread per (input pixel, vector),
// threads of a block share the same channel vectors every `cv` threads.
template <typename T>
__global__ void __launch_bounds__(256)
bn_relu_upsample_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ x,
                            const float* __restrict__ scale, const float* __restrict__ shift,
                            const float* __restrict__ mean, const float* __restrict__ invstd,
                            T* __restrict__ dact, float* __restrict__ dsum,
                            float* __restrict__ ddot, int B, int H, int W, int C, int s) {
  constexpr int VN = Vec16<T>::N;
  extern __shared__ float sh[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int cv = C / VN;
  const int OH = H * s, OW = W * s;
  const size_t total = (size_t)B * H * W * cv;
  // blockDim.x (256) is a multiple of cv whenever cv | 256, so a thread keeps its vector index
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float as[VN], ad[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) { as[e] = 0.f; ad[e] = 0.f; }
  const int v_fixed = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) % cv);
  const bool fixed = (stride % cv) == 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = (int)(i % cv);
    const size_t pix = i / cv;
    const int ix = (int)(pix % W), iy = (int)((pix / W) % H);
    const size_t b = pix / ((size_t)W * H);
    float g[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) g[e] = 0.f;
    const int oy_lo = max(0, s * iy - s), oy_hi = min(OH, s * iy + 2 * s);
    const int ox_lo = max(0, s * ix - s), ox_hi = min(OW, s * ix + 2 * s);
    for (int oy = oy_lo; oy < oy_hi; ++oy) {
      const float wy = bilinear_w(oy, iy, s, H);
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox < ox_hi; ++ox) {
        const float wx = bilinear_w(ox, ix, s, W);
        if (wx == 0.f) continue;
        Vec16<T> d;
        d.load(dout + ((b * OH + oy) * OW + ox) * C + (size_t)v * VN);
        const float wgt = wy * wx;
#pragma unroll
        for (int e = 0; e < VN; ++e) g[e] = fmaf(wgt, d.get(e), g[e]);
      }
    }
    Vec16<T> xv, o;
    xv.load(x + pix * C + (size_t)v * VN);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      const int c = v * VN + e;
      const float xe = xv.get(e);
      const float bn = fmaf(xe, __ldg(scale + c), __ldg(shift + c));
      const float da = bn > 0.f ? g[e] : 0.f;
      o.set(e, da);
      const float xh = (xe - __ldg(mean + c)) * __ldg(invstd + c);
      if (fixed) { as[e] += da; ad[e] += da * xh; }
      else { atomicAdd(&sh[c], da); atomicAdd(&sh[C + c], da * xh); }
    }
    o.store(dact + pix * C + (size_t)v * VN);
  }
  if (fixed) {
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      atomicAdd(&sh[v_fixed * VN + e], as[e]);
      atomicAdd(&sh[C + v_fixed * VN + e], ad[e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dsum + i, sh[i]);
    atomicAdd(ddot + i, sh[C + i]);
  }
}

// Block-structured backward: a thread owns (input pixel, channel vector).  Exactly 2S x 2S output
// pixels carry weight for an input pixel (rows S*iy - S/2 .. S*iy + 3S/2 - 1); their weights are
// evaluated once per thread (2 * 2S calls of bilinear_w, which knows ATen's border clamping) and
// all 16-byte loads of a row are independent.  The per-channel BatchNorm-backward sums stay in
// registers across a thread's pixels and are folded once per block.
template <typename T, int S>
__global__ void __launch_bounds__(256)
bn_relu_upsample_bwd_blk_kernel(const T* __restrict__ dout, const T* __restrict__ x,
                                const float* __restrict__ scale, const float* __restrict__ shift,
                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                T* __restrict__ dact, float* __restrict__ dsum,
                                float* __restrict__ ddot, int B, int H, int W, int C) {
  constexpr int VN = Vec16<T>::N;
  constexpr int NT = 2 * S;
  extern __shared__ float sh[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int cv = C / VN, ppb = 256 / cv;
  const int v = threadIdx.x % cv, pl = threadIdx.x / cv;
  float sc[VN], sf[VN], mu[VN], is[VN], as[VN], ad[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    const int c = v * VN + e;
    sc[e] = __ldg(scale + c); sf[e] = __ldg(shift + c); mu[e] = __ldg(mean + c); is[e] = __ldg(invstd + c);
    as[e] = 0.f; ad[e] = 0.f;
  }
  const int xg = (W + ppb - 1) / ppb;
  const int n_items = B * H * xg;
  const int OH = H * S, OW = W * S;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int xgi = item % xg;
    const int r = item / xg;
    const int iy = r % H, b = r / H;
    const int ix = xgi * ppb + pl;
    if (ix >= W) continue;
    const int oy0 = S * iy - S / 2, ox0 = S * ix - S / 2;
    float wy[NT], wx[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int oy = oy0 + t, ox = ox0 + t;
      wy[t] = (oy >= 0 && oy < OH) ? bilinear_w(oy, iy, S, H) : 0.f;
      wx[t] = (ox >= 0 && ox < OW) ? bilinear_w(ox, ix, S, W) : 0.f;
    }
    float g[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) g[e] = 0.f;
    const T* dbase = dout + (size_t)b * OH * OW * C + (size_t)v * VN;
#pragma unroll
    for (int ty = 0; ty < NT; ++ty) {
      const int oy = min(max(oy0 + ty, 0), OH - 1);      // zero weight where clamped
      Vec16<T> d[NT];
#pragma unroll
      for (int tx = 0; tx < NT; ++tx) {
        const int ox = min(max(ox0 + tx, 0), OW - 1);
        d[tx].load(dbase + ((size_t)oy * OW + ox) * C);
      }
#pragma unroll
      for (int tx = 0; tx < NT; ++tx) {
        const float wgt = wy[ty] * wx[tx];
#pragma unroll
        for (int e = 0; e < VN; ++e) g[e] = fmaf(wgt, d[tx].get(e), g[e]);
      }
    }
    const size_t pix = ((size_t)b * H + iy) * W + ix;
    Vec16<T> xv, o;
    xv.load(x + pix * C + (size_t)v * VN);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      const float xe = xv.get(e);
      const float da = fmaf(xe, sc[e], sf[e]) > 0.f ? g[e] : 0.f;
      o.set(e, da);
      as[e] += da;
      ad[e] = fmaf(da, (xe - mu[e]) * is[e], ad[e]);
    }
    o.store(dact + pix * C + (size_t)v * VN);
  }
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    atomicAdd(&sh[v * VN + e], as[e]);
    atomicAdd(&sh[C + v * VN + e], ad[e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dsum + i, sh[i]);
    atomicAdd(ddot + i, sh[C + i]);
  }
}

template <typename T>
static bool bn_relu_upsample_bwd_blk(const void* dout, const void* x, const float* scale,
                                     const float* shift, const float* mean, const float* invstd,
                                     void* dact, float* dsum, float* ddot, int B, int H, int W, int C,
                                     int s, cudaStream_t stream) {
  constexpr int VN = Vec16<T>::N;
  const int cv = C / VN;
  if ((s != 2 && s != 4) || cv > 256 || 256 % cv) return false;
  const int ppb = 256 / cv;
  const long long items = (long long)B * H * ((W + ppb - 1) / ppb);
  if (items >= (1ll << 31) || (long long)B * H * W * s * s >= (1ll << 31)) return false;
  const int grid = (int)std::min<long long>(items, (long long)s4_num_sms() * 8);
  const size_t smem = 2 * (size_t)C * sizeof(float);
  if (s == 2)
    bn_relu_upsample_bwd_blk_kernel<T, 2><<<grid, 256, smem, stream>>>(
        (const T*)dout, (const T*)x, scale, shift, mean, invstd, (T*)dact, dsum, ddot, B, H, W, C);
  else
    bn_relu_upsample_bwd_blk_kernel<T, 4><<<grid, 256, smem, stream>>>(
        (const T*)dout, (const T*)x, scale, shift, mean, invstd, (T*)dact, dsum, ddot, B, H, W, C);
  return true;
}

extern "C" int s4_bn_relu_upsample_bwd(const void* dout, const void* x, const float* scale,
                                       const float* shift, const float* mean, const float* invstd,
                                       void* dact, float* dsum, float* ddot, int B, int H, int W,
                                       int C, int s, int dtype, cudaStream_t stream) {
  S4ProfScope prof_("bn_relu_upsample_bwd", 0.0, 1, stream);
  const int vn = dtype == S4_BF16 ? 8 : 4;
  S4_REQUIRE(C % vn == 0 && s >= 1, "bn_relu_upsample_bwd: C=%d must be a multiple of %d", C, vn);
  const size_t total = (size_t)B * H * W * (C / vn);
  if (total == 0) return S4_OK;
  if (dtype == S4_BF16 &&
      s4_stream_upsample_bwd(dout, x, scale, shift, mean, invstd, dact, dsum, ddot, B, H, W, C, s, stream))
    return s4_check_launch("bn_relu_upsample_bwd");
  const bool done =
      dtype == S4_BF16
          ? bn_relu_upsample_bwd_blk<__nv_bfloat16>(dout, x, scale, shift, mean, invstd, dact, dsum, ddot, B, H, W, C, s, stream)
          : bn_relu_upsample_bwd_blk<float>(dout, x, scale, shift, mean, invstd, dact, dsum, ddot, B, H, W, C, s, stream);
  if (done) return s4_check_launch("bn_relu_upsample_bwd");
  const int grid = (int)min((total + 255) / 256, (size_t)s4_num_sms() * 16);
  const size_t smem = 2 * (size_t)C * sizeof(float);
  if (dtype == S4_BF16)
    bn_relu_upsample_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(
        (const __nv_bfloat16*)dout, (const __nv_bfloat16*)x, scale, shift, mean, invstd,
        (__nv_bfloat16*)dact, dsum, ddot, B, H, W, C, s);
  else
    bn_relu_upsample_bwd_kernel<float><<<grid, 256, smem, stream>>>(
        (const float*)dout, (const float*)x, scale, shift, mean, invstd, (float*)dact, dsum, ddot,
        B, H, W, C, s);
  return s4_check_launch("bn_relu_upsample_bwd");
}

// dy = gamma*invstd*(dact - dsum/n - xhat*ddot/n) = A*dact + Bc*x + Cc with per-channel constants
// kept in registers (the grid stride is a multiple of the vectors per row, so a thread keeps its
// channel vector)
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ dact, const T* __restrict__ x,
                    const float* __restrict__ gamma, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ dsum,
                    const float* __restrict__ ddot, float inv_n, T* __restrict__ dy, size_t rows,
                    int C) {
  constexpr int VN = Vec16<T>::N;
  const int cv = C / VN;
  const size_t total = rows * cv;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool fixed = stride % cv == 0;
  float A[VN], Bc[VN], Cc[VN];
  auto coeffs = [&](int v) {
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      const int c = v * VN + e;
      const float is = __ldg(invstd + c), a = __ldg(gamma + c) * is;
      const float k = is * __ldg(ddot + c) * inv_n;          // xhat*ddot/n = (x - mean) * k
      A[e] = a;
      Bc[e] = -a * k;
      Cc[e] = a * (__ldg(mean + c) * k - __ldg(dsum + c) * inv_n);
    }
  };
  if (fixed) coeffs((int)(i0 % cv));
  for (size_t i = i0; i < total; i += stride) {
    if (!fixed) coeffs((int)(i % cv));
    Vec16<T> d, xv, o;
    d.load(dact + i * VN);
    xv.load(x + i * VN);
#pragma unroll
    for (int e = 0; e < VN; ++e) o.set(e, fmaf(A[e], d.get(e), fmaf(Bc[e], xv.get(e), Cc[e])));
    o.store(dy + i * VN);
  }
}

extern "C" int s4_bn_bwd_apply(const void* dact, const void* x, const float* gamma,
                               const float* mean, const float* invstd, const float* dsum,
                               const float* ddot, double count, void* dy, long long rows, int C,
                               int dtype, cudaStream_t stream) {
  S4ProfScope prof_("bn_bwd_apply", 0.0, 1, stream);
  const int vn = dtype == S4_BF16 ? 8 : 4;
  S4_REQUIRE(C % vn == 0, "bn_bwd_apply: C=%d must be a multiple of %d", C, vn);
  const size_t total = (size_t)rows * (C / vn);
  if (total == 0) return S4_OK;
  if (dtype == S4_BF16 &&
      s4_stream_bn_bwd_apply(dact, x, gamma, mean, invstd, dsum, ddot, count, dy, rows, C, stream))
    return s4_check_launch("bn_bwd_apply");
  const int grid = (int)min((total + 255) / 256, (size_t)s4_num_sms() * 32);
  const float inv_n = (float)(1.0 / count);
  if (dtype == S4_BF16)
    bn_bwd_apply_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)dact, (const __nv_bfloat16*)x, gamma, mean, invstd, dsum, ddot, inv_n,
        (__nv_bfloat16*)dy, (size_t)rows, C);
  else
    bn_bwd_apply_kernel<float><<<grid, 256, 0, stream>>>((const float*)dact, (const float*)x, gamma,
                                                         mean, invstd, dsum, ddot, inv_n, (float*)dy,
                                                         (size_t)rows, C);
  return s4_check_launch("bn_bwd_apply");
}

// ------------------------------------------------------------------------------------------
// last stage: z[row, j] = bias[j] + sum_c w[j,c] * relu(x[row,c]*scale[c] + shift[c])
// one warp per pixel row; C <= 1024, NC <= 32
// ------------------------------------------------------------------------------------------
#define MAXNC 32
template <typename T>
__global__ void __launch_bounds__(256)
bn_relu_conv1x1_fwd_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                           const float* __restrict__ shift, const float* __restrict__ w,
                           const float* __restrict__ bias, float* __restrict__ z, size_t rows,
                           int C, int NC) {
  extern __shared__ float ws[];  // [NC][C] + scale[C] + shift[C]
  float* ssc = ws + (size_t)NC * C;
  float* ssh = ssc + C;
  for (int i = threadIdx.x; i < NC * C; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { ssc[i] = scale[i]; ssh[i] = shift[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t row = warp0; row < rows; row += nwarps) {
    float acc[MAXNC];
#pragma unroll
    for (int j = 0; j < MAXNC; ++j) acc[j] = 0.f;
    for (int c = lane; c < C; c += 32) {      // lane-strided channels: conflict-free smem reads
      const float a = fmaxf(fmaf(to_f32<T>(x[row * C + c]), ssc[c], ssh[c]), 0.f);
#pragma unroll
      for (int j = 0; j < MAXNC; ++j)
        if (j < NC) acc[j] = fmaf(a, ws[j * C + c], acc[j]);
    }
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < MAXNC; ++j) {
      if (j < NC) {
        const float t = warp_sum(acc[j]);
        if (lane == j) mine = t;
      }
    }
    if (lane < NC) z[row * NC + lane] = mine + bias[lane];
  }
}

extern "C" int s4_bn_relu_conv1x1_fwd(const void* x, const float* scale, const float* shift,
                                      const float* w, const float* bias, float* z, long long rows,
                                      int C, int NC, int dtype, cudaStream_t stream) {
  S4ProfScope prof_("bn_relu_conv1x1_fwd", 0.0, 1, stream);
  S4_REQUIRE(NC >= 1 && NC <= MAXNC, "conv1x1: NC=%d not in [1,%d]", NC, MAXNC);
  if (rows == 0) return S4_OK;
  if (s4_cls_tc_supported(C, NC, dtype)) return s4_cls_fwd_tc(x, scale, shift, w, bias, z, rows, C, NC, stream);
  const size_t smem = ((size_t)NC * C + 2 * C) * sizeof(float);
  S4_REQUIRE(smem <= 200 * 1024, "conv1x1: C*NC too large for shared memory");
  const int grid = (int)min(((size_t)rows + 7) / 8, (size_t)s4_num_sms() * 8);
  if (dtype == S4_BF16) {
    auto k = bn_relu_conv1x1_fwd_kernel<__nv_bfloat16>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, 256, smem, stream>>>((const __nv_bfloat16*)x, scale, shift, w, bias, z, (size_t)rows, C, NC);
  } else {
    auto k = bn_relu_conv1x1_fwd_kernel<float>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, 256, smem, stream>>>((const float*)x, scale, shift, w, bias, z, (size_t)rows, C, NC);
  }
  return s4_check_launch("bn_relu_conv1x1_fwd");
}

// backward: thread c owns channel c for a strip of rows (C <= 1024 threads per block)
template <typename T>
__global__ void bn_relu_conv1x1_bwd_kernel(const float* __restrict__ dz, const T* __restrict__ x,
                                           const float* __restrict__ scale,
                                           const float* __restrict__ shift,
                                           const float* __restrict__ mean,
                                           const float* __restrict__ invstd,
                                           const float* __restrict__ w, T* __restrict__ dact,
                                           float* __restrict__ dw, float* __restrict__ dbias,
                                           float* __restrict__ dsum, float* __restrict__ ddot,
                                           size_t rows, int C, int NC, size_t rows_per_blk) {
  __shared__ float sdz[32][MAXNC];
  const int c = threadIdx.x;   // blockDim.x == C
  const size_t r0 = (size_t)blockIdx.x * rows_per_blk;
  const size_t r1 = min(rows, r0 + rows_per_blk);
  float wc[MAXNC], gw[MAXNC];
#pragma unroll
  for (int j = 0; j < MAXNC; ++j) { wc[j] = (j < NC) ? w[j * C + c] : 0.f; gw[j] = 0.f; }
  const float sc = scale[c], sf = shift[c], mu = mean[c], is = invstd[c];
  float s_sum = 0.f, s_dot = 0.f, s_bias = 0.f;
  for (size_t rb = r0; rb < r1; rb += 32) {
    const int nr = (int)min((size_t)32, r1 - rb);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * NC; i += blockDim.x) sdz[i / NC][i % NC] = dz[rb * NC + i];
    __syncthreads();
    if (c < NC)
      for (int r = 0; r < nr; ++r) s_bias += sdz[r][c];
    for (int r = 0; r < nr; ++r) {
      const float xe = to_f32<T>(x[(rb + r) * C + c]);
      const float bn = fmaf(xe, sc, sf);
      const float a = fmaxf(bn, 0.f);
      float g = 0.f;
#pragma unroll
      for (int j = 0; j < MAXNC; ++j)
        if (j < NC) {
          const float d = sdz[r][j];
          g = fmaf(d, wc[j], g);
          gw[j] = fmaf(d, a, gw[j]);
        }
      const float da = bn > 0.f ? g : 0.f;
      dact[(rb + r) * C + c] = from_f32<T>(da);
      s_sum += da;
      s_dot += da * (xe - mu) * is;
    }
  }
#pragma unroll
  for (int j = 0; j < MAXNC; ++j)
    if (j < NC) atomicAdd(dw + j * C + c, gw[j]);
  atomicAdd(dsum + c, s_sum);
  atomicAdd(ddot + c, s_dot);
  if (c < NC) atomicAdd(dbias + c, s_bias);
}

extern "C" int s4_bn_relu_conv1x1_bwd(const float* dz, const void* x, const float* scale,
                                      const float* shift, const float* mean, const float* invstd,
                                      const float* w, void* dact, float* dw, float* dbias,
                                      float* dsum, float* ddot, long long rows, int C, int NC,
                                      int dtype, cudaStream_t stream) {
  S4ProfScope prof_("bn_relu_conv1x1_bwd", 0.0, 1, stream);
  S4_REQUIRE(NC >= 1 && NC <= MAXNC, "conv1x1_bwd: NC=%d not in [1,%d]", NC, MAXNC);
  S4_REQUIRE(C >= NC && C <= 1024 && C % 32 == 0, "conv1x1_bwd: C=%d must be a multiple of 32 in [NC,1024]", C);
  if (rows == 0) return S4_OK;
  int blocks = s4_num_sms() * (C <= 256 ? 8 : 2);
  size_t rpb = ((size_t)rows + blocks - 1) / blocks;
  rpb = ((rpb + 31) / 32) * 32;
  blocks = (int)(((size_t)rows + rpb - 1) / rpb);
  if (dtype == S4_BF16)
    bn_relu_conv1x1_bwd_kernel<__nv_bfloat16><<<blocks, C, 0, stream>>>(
        dz, (const __nv_bfloat16*)x, scale, shift, mean, invstd, w, (__nv_bfloat16*)dact, dw, dbias,
        dsum, ddot, (size_t)rows, C, NC, rpb);
  else
    bn_relu_conv1x1_bwd_kernel<float><<<blocks, C, 0, stream>>>(
        dz, (const float*)x, scale, shift, mean, invstd, w, (float*)dact, dw, dbias, dsum, ddot,
        (size_t)rows, C, NC, rpb);
  return s4_check_launch("bn_relu_conv1x1_bwd");
}

// ------------------------------------------------------------------------------------------
// logits[B,NC,OH,OW] (NCHW f32) = bilinear_s(z[B,H,W,NC]) and the transpose
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
upsample_logits_fwd_kernel(const float* __restrict__ z, float* __restrict__ out, int B, int H,
                           int W, int NC, int s) {
  const int OH = H * s, OW = W * s;
  const size_t total = (size_t)B * NC * OH * OW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
    const int j = (int)((i / ((size_t)OW * OH)) % NC);
    const size_t b = i / ((size_t)OW * OH * NC);
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(oy, s, H, y0, y1, ly);
    bilinear_src(ox, s, W, x0, x1, lx);
    const float* zb = z + b * (size_t)H * W * NC + j;
    const float v00 = __ldg(zb + ((size_t)y0 * W + x0) * NC), v01 = __ldg(zb + ((size_t)y0 * W + x1) * NC);
    const float v10 = __ldg(zb + ((size_t)y1 * W + x0) * NC), v11 = __ldg(zb + ((size_t)y1 * W + x1) * NC);
    out[i] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

__global__ void __launch_bounds__(256)
upsample_logits_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dz, int B, int H,
                           int W, int NC, int s) {
  const int OH = H * s, OW = W * s;
  const size_t total = (size_t)B * NC * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(i % W), iy = (int)((i / W) % H);
    const int j = (int)((i / ((size_t)W * H)) % NC);
    const size_t b = i / ((size_t)W * H * NC);
    const float* d = dout + (b * NC + j) * (size_t)OH * OW;
    float g = 0.f;
    const int oy_lo = max(0, s * iy - s), oy_hi = min(OH, s * iy + 2 * s);
    const int ox_lo = max(0, s * ix - s), ox_hi = min(OW, s * ix + 2 * s);
    for (int oy = oy_lo; oy < oy_hi; ++oy) {
      const float wy = bilinear_w(oy, iy, s, H);
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox < ox_hi; ++ox) {
        const float wx = bilinear_w(ox, ix, s, W);
        if (wx != 0.f) g = fmaf(wy * wx, __ldg(d + (size_t)oy * OW + ox), g);
      }
    }
    dz[((b * H + iy) * W + ix) * NC + j] = g;
  }
}

extern "C" int s4_upsample_logits_fwd(const float* z, float* logits, int B, int H, int W, int NC,
                                      int s, cudaStream_t stream) {
  S4ProfScope prof_("upsample_logits_fwd", 0.0, 1, stream);
  const size_t total = (size_t)B * NC * H * s * W * s;
  if (total == 0) return S4_OK;
  if ((size_t)2 * W * NC * 4 <= 160 * 1024) return s4_cls_upsample_fwd(z, logits, B, H, W, NC, s, stream);
  const int grid = (int)min((total + 255) / 256, (size_t)s4_num_sms() * 32);
  upsample_logits_fwd_kernel<<<grid, 256, 0, stream>>>(z, logits, B, H, W, NC, s);
  return s4_check_launch("upsample_logits_fwd");
}

extern "C" int s4_upsample_logits_bwd(const float* dlogits, float* dz, int B, int H, int W, int NC,
                                      int s, cudaStream_t stream) {
  S4ProfScope prof_("upsample_logits_bwd", 0.0, 1, stream);
  const size_t total = (size_t)B * NC * H * W;
  if (total == 0) return S4_OK;
  const int grid = (int)min((total + 255) / 256, (size_t)s4_num_sms() * 32);
  upsample_logits_bwd_kernel<<<grid, 256, 0, stream>>>(dlogits, dz, B, H, W, NC, s);
  return s4_check_launch("upsample_logits_bwd");
}

// ---- bf16 tensor-core backward of the last stage (head_cls.cu) ---------------------------------
extern "C" int s4_cls_supported(int C, int NC, int dtype) { return s4_cls_tc_supported(C, NC, dtype) ? 1 : 0; }

extern "C" int s4_cls_upsample_bwd_padded(const float* dlogits, void* dz16, int B, int H, int W, int NC,
                                          int s, cudaStream_t stream) {
  S4ProfScope prof_("cls_upsample_bwd", 0.0, 1, stream);
  S4_REQUIRE(NC >= 1 && NC <= 32, "cls_upsample_bwd: NC=%d not in [1,32]", NC);
  S4_REQUIRE((size_t)NC * W * s * 4 <= 200 * 1024, "cls_upsample_bwd: row too wide for shared memory");
  if ((size_t)B * H * W == 0) return S4_OK;
  return s4_cls_upsample_bwd(dlogits, dz16, B, H, W, NC, s, stream);
}

extern "C" int s4_cls_bwd_reduce(const void* dz16, const void* y, const float* scale, const float* shift,
                                 const float* mean, const float* invstd, const float* w, float* dw,
                                 float* dbias, float* dsum, float* ddot, long long rows, int C, int NC,
                                 cudaStream_t stream) {
  S4ProfScope prof_("cls_bwd_reduce", 0.0, 1, stream);
  S4_REQUIRE(s4_cls_tc_supported(C, NC, S4_BF16), "cls_bwd_reduce: unsupported C=%d NC=%d", C, NC);
  if (rows == 0) return S4_OK;
  return s4_cls_bwd_reduce_tc(dz16, y, scale, shift, mean, invstd, w, dw, dbias, dsum, ddot, rows, C, NC, stream);
}

extern "C" int s4_cls_bwd_apply(const void* dz16, const void* y, const float* scale, const float* shift,
                                const float* mean, const float* invstd, const float* gamma, const float* w,
                                const float* dsum, const float* ddot, double count, void* dy,
                                long long rows, int C, int NC, cudaStream_t stream) {
  S4ProfScope prof_("cls_bwd_apply", 0.0, 1, stream);
  S4_REQUIRE(s4_cls_tc_supported(C, NC, S4_BF16), "cls_bwd_apply: unsupported C=%d NC=%d", C, NC);
  if (rows == 0) return S4_OK;
  return s4_cls_bwd_apply_tc(dz16, y, scale, shift, mean, invstd, gamma, w, dsum, ddot, count, dy, rows, C, NC, stream);
}
