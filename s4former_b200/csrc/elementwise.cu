// Small HBM-bound kernels around the backbone: patchify (im2col-free patch embedding input),
// cls/pos-embed assembly and its backward, column sums (bias / BN statistics), dtype casts,
// multi-tensor EMA and SGD, CutMix / PatchShuffle gathers.
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// patchify: img [B,3,H,W] f32 (NCHW)  ->  A [B*gh*gw, Cin*P*P] with k = (c*P + ky)*P + kx
// (the weight layout of Conv2d(3,768,16,16), reference embed.py:145-153 / 199-201).
// Non-overlapping patches => pure gather, no im2col blow-up.  Corner padding (embed.py:58-80)
// is realised by zero-filling reads beyond H/W.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void patchify_kernel(const float* __restrict__ img, T* __restrict__ out, int B, int Cin,
                                int H, int W, int P, int gh, int gw) {
  const int K = Cin * P * P;
  const size_t total = (size_t)B * gh * gw * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const size_t row = i / K;
    const int px = (int)(row % gw), py = (int)((row / gw) % gh), b = (int)(row / ((size_t)gw * gh));
    const int kx = k % P, ky = (k / P) % P, c = k / (P * P);
    const int y = py * P + ky, x = px * P + kx;
    float v = 0.f;
    if (y < H && x < W) v = __ldg(img + (((size_t)b * Cin + c) * H + y) * W + x);
    out[i] = from_f32<T>(v);
  }
}

// Vector path (P % 8 == 0, W % 4 == 0, bf16 out): one thread moves 8 consecutive pixels of a patch row
// (two 16-byte loads -> one 16-byte store) and does its index arithmetic once, in 32 bits.
__global__ void __launch_bounds__(256)
patchify_vec8_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int Cin,
                     int H, int W, int P, int gh, int gw) {
  const int P8 = P >> 3;                       // 8-pixel groups per patch row
  const int K8 = Cin * P * P8;                 // groups per output row
  const unsigned total = (unsigned)B * gh * gw * K8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned k8 = i % K8, row = i / K8;
    const int kx = (int)(k8 % P8) * 8, ky = (int)(k8 / P8) % P, c = (int)(k8 / (P8 * P));
    const int px = (int)(row % gw), py = (int)(row / gw) % gh, b = (int)(row / (gw * gh));
    const int y = py * P + ky, x = px * P + kx;
    float v[8];
    if (y < H && x + 8 <= W) {
      const float4* src = reinterpret_cast<const float4*>(img + (((size_t)b * Cin + c) * H + y) * W + x);
      const float4 a = __ldg(src), d = __ldg(src + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = d.x; v[5] = d.y; v[6] = d.z; v[7] = d.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        v[e] = (y < H && x + e < W) ? __ldg(img + (((size_t)b * Cin + c) * H + y) * W + x + e) : 0.f;
    }
    uint4 o;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    o.x = *reinterpret_cast<uint32_t*>(&h0);
    o.y = *reinterpret_cast<uint32_t*>(&h1);
    o.z = *reinterpret_cast<uint32_t*>(&h2);
    o.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(out + (size_t)i * 8) = o;
  }
}

extern "C" int s4_patchify(const float* img, void* out, int B, int Cin, int H, int W, int P,
                           int dtype, cudaStream_t stream) {
  S4ProfScope prof_("patchify", 0.0, 1, stream);
  const int gh = (H + P - 1) / P, gw = (W + P - 1) / P;
  const size_t total = (size_t)B * gh * gw * Cin * P * P;
  if (total == 0) return S4_OK;
  if (dtype == S4_BF16 && P % 8 == 0 && W % 4 == 0 && total / 8 < (1ull << 31) &&
      ((uintptr_t)img & 15) == 0 && ((uintptr_t)out & 15) == 0) {
    const size_t groups = total / 8;
    const int grid = (int)min((groups + 255) / 256, (size_t)s4_num_sms() * 16);
    patchify_vec8_kernel<<<grid, 256, 0, stream>>>(img, (__nv_bfloat16*)out, B, Cin, H, W, P, gh, gw);
    return s4_check_launch("patchify");
  }
  const int grid = (int)min((total + 255) / 256, (size_t)s4_num_sms() * 16);
  if (dtype == S4_BF16)
    patchify_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(img, (__nv_bfloat16*)out, B, Cin, H, W, P, gh, gw);
  else
    patchify_kernel<float><<<grid, 256, 0, stream>>>(img, (float*)out, B, Cin, H, W, P, gh, gw);
  return s4_check_launch("patchify");
}

// ------------------------------------------------------------------------------------------
// tokens assembly (reference vit.py:486-487, :513):
//   x[b,0,:]   = cls + pos[0]
//   x[b,1+p,:] = patch_tok[b,p,:] + pos[1+p]
// and its backward: dpos[l] = sum_b dx[b,l]; dcls = sum_b dx[b,0]; dtok = dx[:,1:]
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void assemble_tokens_kernel(const T* __restrict__ tok, const float* __restrict__ cls,
                                       const float* __restrict__ pos, T* __restrict__ x, int B,
                                       int L, int D) {
  const size_t total = (size_t)B * L * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const int l = (int)((i / D) % L);
    const size_t b = i / ((size_t)D * L);
    float v = __ldg(pos + (size_t)l * D + d);
    if (l == 0) v += __ldg(cls + d);
    else v += to_f32<T>(tok[(b * (L - 1) + (l - 1)) * D + d]);
    x[i] = from_f32<T>(v);
  }
}

template <typename T>
__global__ void assemble_tokens_bwd_kernel(const T* __restrict__ dx, T* __restrict__ dtok,
                                           float* __restrict__ dcls, float* __restrict__ dpos,
                                           int B, int L, int D) {
  const size_t total = (size_t)L * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const int l = (int)(i / D);
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      const T g = dx[((size_t)b * L + l) * D + d];
      acc += to_f32<T>(g);
      if (l > 0) dtok[((size_t)b * (L - 1) + (l - 1)) * D + d] = g;
    }
    dpos[i] += acc;
    if (l == 0) dcls[d] += acc;
  }
}

extern "C" int s4_assemble_tokens(const void* tok, const float* cls, const float* pos, void* x,
                                  int B, int L, int D, int dtype, cudaStream_t stream) {
  S4ProfScope prof_("assemble_tokens", 0.0, 1, stream);
  const size_t total = (size_t)B * L * D;
  if (total == 0) return S4_OK;
  const int grid = (int)min((total + 255) / 256, (size_t)s4_num_sms() * 16);
  if (dtype == S4_BF16)
    assemble_tokens_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)tok, cls, pos, (__nv_bfloat16*)x, B, L, D);
  else
    assemble_tokens_kernel<float><<<grid, 256, 0, stream>>>((const float*)tok, cls, pos, (float*)x, B, L, D);
  return s4_check_launch("assemble_tokens");
}

// dcls / dpos are ACCUMULATED (+=)
extern "C" int s4_assemble_tokens_bwd(const void* dx, void* dtok, float* dcls, float* dpos, int B,
                                      int L, int D, int dtype, cudaStream_t stream) {
  S4ProfScope prof_("assemble_tokens_bwd", 0.0, 1, stream);
  const size_t total = (size_t)L * D;
  if (total == 0 || B == 0) return S4_OK;
  const int grid = (int)min((total + 255) / 256, (size_t)s4_num_sms() * 16);
  if (dtype == S4_BF16)
    assemble_tokens_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)dx, (__nv_bfloat16*)dtok, dcls, dpos, B, L, D);
  else
    assemble_tokens_bwd_kernel<float><<<grid, 256, 0, stream>>>((const float*)dx, (float*)dtok, dcls, dpos, B, L, D);
  return s4_check_launch("assemble_tokens_bwd");
}

// ------------------------------------------------------------------------------------------
// column sums of a row-major [rows, cols] matrix: out[c] (+)= sum_r f(x[r,c]); optionally also
// the sum of squares (BatchNorm statistics over N*H*W of an NHWC tensor).  Threads walk columns
// (coalesced), blocks split rows, fp32 atomics combine the per-block partials.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, float* __restrict__ sum,
                              float* __restrict__ sumsq, size_t rows, int cols, int rows_per_blk) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const size_t r0 = (size_t)blockIdx.y * rows_per_blk;
  const size_t r1 = min(rows, r0 + rows_per_blk);
  float s = 0.f, q = 0.f;
  for (size_t r = r0; r < r1; ++r) {
    const float v = to_f32<T>(x[r * cols + c]);
    s += v;
    q += v * v;
  }
  atomicAdd(sum + c, s);
  if (sumsq) atomicAdd(sumsq + c, q);
}

// Vectorised version (cols % (16 B of T) == 0): a block owns a 32-vector column slab and a strip
// of rows; each of its 8 warps walks every 8th row of the strip with 16-byte loads, 4 rows in
// flight per thread; the warps fold through shared memory and the block issues one fp32 atomic
// per column.  SQ selects the sum-of-squares output at compile time.
template <typename T, bool SQ>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ x, float* __restrict__ sum, float* __restrict__ sumsq,
                  size_t rows, int cols, int rows_per_blk) {
  constexpr int VN = Vec16<T>::N;
  __shared__ float sh[8][2][32 * VN + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + lane) * VN;
  const bool col_ok = c0 < cols;
  const size_t r0 = (size_t)blockIdx.y * rows_per_blk;
  const size_t r1 = min(rows, r0 + rows_per_blk);
  float s[VN], q[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) { s[e] = 0.f; q[e] = 0.f; }
  if (col_ok) {
    const T* xp = x + c0;
    size_t r = r0 + warp;
    for (; r + 24 < r1; r += 32) {
      Vec16<T> v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u].load(xp + (r + 8 * u) * cols);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const float f = v[u].get(e);
          s[e] += f;
          if (SQ) q[e] = fmaf(f, f, q[e]);
        }
    }
    for (; r < r1; r += 8) {
      Vec16<T> v;
      v.load(xp + r * cols);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const float f = v.get(e);
        s[e] += f;
        if (SQ) q[e] = fmaf(f, f, q[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    sh[warp][0][lane * VN + e] = s[e];
    if (SQ) sh[warp][1][lane * VN + e] = q[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * VN; i += 256) {
    const int c = blockIdx.x * 32 * VN + i;
    if (c >= cols) continue;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      a += sh[w][0][i];
      if (SQ) b += sh[w][1][i];
    }
    atomicAdd(sum + c, a);
    if (SQ) atomicAdd(sumsq + c, b);
  }
}

template <typename T>
static void colsum_vec_launch(const void* x, float* sum, float* sumsq, long long rows, int cols,
                              cudaStream_t stream) {
  constexpr int VN = Vec16<T>::N;
  const int gx = (cols + 32 * VN - 1) / (32 * VN);
  int gy = (s4_num_sms() * 6 + gx - 1) / gx;
  const long long max_gy = (rows + 31) / 32;          // at least 32 rows (4 per warp) per block
  if ((long long)gy > max_gy) gy = (int)max_gy;
  if (gy < 1) gy = 1;
  const int rpb = (int)((rows + gy - 1) / gy);
  gy = (int)((rows + rpb - 1) / rpb);
  dim3 grid(gx, gy);
  if (sumsq)
    colsum_vec_kernel<T, true><<<grid, 256, 0, stream>>>((const T*)x, sum, sumsq, (size_t)rows, cols, rpb);
  else
    colsum_vec_kernel<T, false><<<grid, 256, 0, stream>>>((const T*)x, sum, nullptr, (size_t)rows, cols, rpb);
}

// sum / sumsq are ACCUMULATED (+=): zero them first when a fresh sum is wanted.
extern "C" int s4_colsum(const void* x, float* sum, float* sumsq, long long rows, int cols,
                         int dtype, cudaStream_t stream) {
  S4ProfScope prof_("colsum", 0.0, 1, stream);
  if (rows == 0 || cols == 0) return S4_OK;
  const int vn = dtype == S4_BF16 ? 8 : 4;
  if (cols % vn == 0 && (((uintptr_t)x) & 15) == 0) {
    if (dtype == S4_BF16) colsum_vec_launch<__nv_bfloat16>(x, sum, sumsq, rows, cols, stream);
    else colsum_vec_launch<float>(x, sum, sumsq, rows, cols, stream);
    return s4_check_launch("colsum");
  }
  const int bx = 128;
  const int gx = (cols + bx - 1) / bx;
  int gy = (s4_num_sms() * 8 + gx - 1) / gx;
  if ((long long)gy > rows) gy = (int)rows;
  const int rpb = (int)((rows + gy - 1) / gy);
  gy = (int)((rows + rpb - 1) / rpb);
  dim3 grid(gx, gy);
  if (dtype == S4_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, bx, 0, stream>>>((const __nv_bfloat16*)x, sum, sumsq, (size_t)rows, cols, rpb);
  else
    colsum_kernel<float><<<grid, bx, 0, stream>>>((const float*)x, sum, sumsq, (size_t)rows, cols, rpb);
  return s4_check_launch("colsum");
}

// ------------------------------------------------------------------------------------------
// casts
// ------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(y)[i] = o;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __float2bfloat16_rn(x[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __bfloat162float(x[i]);
}

extern "C" int s4_cast(const void* x, void* y, long long n, int src_dtype, int dst_dtype,
                       cudaStream_t stream) {
  S4ProfScope prof_("cast", 0.0, 1, stream);
  if (n == 0) return S4_OK;
  const int grid = (int)min(((size_t)n / 4 + 255) / 256 + 1, (size_t)s4_num_sms() * 16);
  if (src_dtype == S4_F32 && dst_dtype == S4_BF16)
    cast_f32_bf16_kernel<<<grid, 256, 0, stream>>>((const float*)x, (__nv_bfloat16*)y, (size_t)n);
  else if (src_dtype == S4_BF16 && dst_dtype == S4_F32)
    cast_bf16_f32_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (float*)y, (size_t)n);
  else {
    s4_set_error("cast: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
    return S4_ERR_UNSUPPORTED;
  }
  return s4_check_launch("cast");
}

// 2-D transpose  y[c, r] = x[r, c]   (weights for dgrad, activations for wgrad fallbacks)
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ x, T* __restrict__ y, int rows, int cols,
                                 size_t xbs, size_t ybs) {
  __shared__ T tile[32][33];
  x += blockIdx.z * xbs;
  y += blockIdx.z * ybs;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = x[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) y[(size_t)c * rows + r] = tile[threadIdx.x][j];
  }
}

extern "C" int s4_transpose(const void* x, void* y, int batch, int rows, int cols, int dtype,
                            cudaStream_t stream) {
  S4ProfScope prof_("transpose", 0.0, 1, stream);
  if (batch == 0 || rows == 0 || cols == 0) return S4_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch), block(32, 8);
  const size_t bs = (size_t)rows * cols;
  if (dtype == S4_BF16)
    transpose_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, rows, cols, bs, bs);
  else
    transpose_kernel<float><<<grid, block, 0, stream>>>((const float*)x, (float*)y, rows, cols, bs, bs);
  return s4_check_launch("transpose");
}

// ------------------------------------------------------------------------------------------
// multi-tensor EMA / SGD  (reference encoder_decoder.py:1044-1066; torch.optim.SGD as driven by
// mmcv OptimizerHook with paramwise lr_mult).  One launch for the whole model: a chunk table
// maps each block to (tensor, offset).
// ------------------------------------------------------------------------------------------
struct S4TensorTable {
  void* const* a;         // dst / param
  void* const* b;         // src / grad
  void* const* c;         // momentum buffer (SGD) or null
  const long long* size;  // elements per tensor
  const float* scalar;    // per-tensor lr (SGD) or null
  const int* chunk_tensor;
  const long long* chunk_off;
};

#define S4_CHUNK 16384

__global__ void __launch_bounds__(256)
ema_multi_kernel(S4TensorTable t, void* const* shadow, float momentum, float one_minus) {
  const int ti = t.chunk_tensor[blockIdx.x];
  const long long off = t.chunk_off[blockIdx.x];
  const long long n = min((long long)S4_CHUNK, t.size[ti] - off);
  float* dst = (float*)t.a[ti] + off;
  const float* src = (const float*)t.b[ti] + off;
  __nv_bfloat16* sh = (shadow && shadow[ti]) ? (__nv_bfloat16*)shadow[ti] + off : nullptr;
  // chunk offsets are multiples of S4_CHUNK, tensor bases are 16B aligned (torch allocator)
  const bool vec = ((((uintptr_t)dst) | ((uintptr_t)src)) & 15) == 0 && (((uintptr_t)sh) & 7) == 0;
  if (vec) {
    const long long n4 = n / 4;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 d = reinterpret_cast<float4*>(dst)[i];
      const float4 s = reinterpret_cast<const float4*>(src)[i];
      d.x = fmaf(s.x, one_minus, d.x * momentum);
      d.y = fmaf(s.y, one_minus, d.y * momentum);
      d.z = fmaf(s.z, one_minus, d.z * momentum);
      d.w = fmaf(s.w, one_minus, d.w * momentum);
      reinterpret_cast<float4*>(dst)[i] = d;
      if (sh) {
        __nv_bfloat162 a = __floats2bfloat162_rn(d.x, d.y), b = __floats2bfloat162_rn(d.z, d.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&a);
        o.y = *reinterpret_cast<uint32_t*>(&b);
        reinterpret_cast<uint2*>(sh)[i] = o;
      }
    }
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      dst[i] = fmaf(src[i], one_minus, dst[i] * momentum);
      if (sh) sh[i] = __float2bfloat16_rn(dst[i]);
    }
  } else {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      dst[i] = fmaf(src[i], one_minus, dst[i] * momentum);
      if (sh) sh[i] = __float2bfloat16_rn(dst[i]);
    }
  }
}

extern "C" int s4_ema_multi_tensor(void* const* dst_ptrs, void* const* src_ptrs,
                                   void* const* bf16_shadow, const long long* sizes,
                                   const int* chunk_tensor, const long long* chunk_off,
                                   int n_chunks, float momentum, float one_minus_momentum,
                                   cudaStream_t stream) {
  S4ProfScope prof_("ema_multi_tensor", 0.0, 1, stream);
  if (n_chunks == 0) return S4_OK;
  S4TensorTable t{dst_ptrs, src_ptrs, nullptr, sizes, nullptr, chunk_tensor, chunk_off};
  ema_multi_kernel<<<n_chunks, 256, 0, stream>>>(t, bf16_shadow, momentum, one_minus_momentum);
  return s4_check_launch("ema_multi_tensor");
}

// SGD with momentum (dampening 0, no nesterov), weight decay, per-tensor lr:
//   g' = g + wd*p ; buf = mu*buf + g' (buf = g' on the first step) ; p -= lr*buf
// optionally refreshes a bf16 shadow copy of the parameter in the same pass.
__global__ void __launch_bounds__(256)
sgd_multi_kernel(S4TensorTable t, void* const* shadow, float mu, float wd, int first_step) {
  const int ti = t.chunk_tensor[blockIdx.x];
  const long long off = t.chunk_off[blockIdx.x];
  const long long n = min((long long)S4_CHUNK, t.size[ti] - off);
  float* p = (float*)t.a[ti] + off;
  const float* g = (const float*)t.b[ti] + off;
  float* buf = (float*)t.c[ti] + off;
  __nv_bfloat16* sh = (shadow && shadow[ti]) ? (__nv_bfloat16*)shadow[ti] + off : nullptr;
  const float lr = t.scalar[ti];
  auto upd = [&](float gi, float& pi, float& bi) {
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    bi = first_step ? gi : fmaf(mu, bi, gi);
    pi = fmaf(-lr, bi, pi);
  };
  const bool vec = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)buf)) & 15) == 0 && (((uintptr_t)sh) & 7) == 0;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 bv = first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(buf)[i];
    upd(gv.x, pv.x, bv.x);
    upd(gv.y, pv.y, bv.y);
    upd(gv.z, pv.z, bv.z);
    upd(gv.w, pv.w, bv.w);
    reinterpret_cast<float4*>(buf)[i] = bv;
    reinterpret_cast<float4*>(p)[i] = pv;
    if (sh) {
      __nv_bfloat162 a = __floats2bfloat162_rn(pv.x, pv.y), b = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(sh)[i] = o;
    }
  }
  for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
    float pi = p[i];
    float bi = first_step ? 0.f : buf[i];
    upd(g[i], pi, bi);
    buf[i] = bi;
    p[i] = pi;
    if (sh) sh[i] = __float2bfloat16_rn(pi);
  }
}

extern "C" int s4_sgd_multi_tensor(void* const* params, void* const* grads, void* const* bufs,
                                   void* const* bf16_shadow, const long long* sizes,
                                   const float* lrs, const int* chunk_tensor,
                                   const long long* chunk_off, int n_chunks, float momentum,
                                   float weight_decay, int first_step, cudaStream_t stream) {
  S4ProfScope prof_("sgd_multi_tensor", 0.0, 1, stream);
  if (n_chunks == 0) return S4_OK;
  S4TensorTable t{params, grads, bufs, sizes, lrs, chunk_tensor, chunk_off};
  sgd_multi_kernel<<<n_chunks, 256, 0, stream>>>(t, bf16_shadow, momentum, weight_decay, first_step);
  return s4_check_launch("sgd_multi_tensor");
}

// SGD-momentum step AND the EMA-teacher update in ONE sweep.
// The reference runs update_ema_variables at the START of step t+1 on the weights SGD wrote at
// step t (encoder_decoder.py:416-423, 1044-1066): applying t <- m t + (1-m) s right after the SGD
// write of s is the same arithmetic on the same values, one step earlier in program order and
// with nothing reading the teacher in between.  Reads grad / momentum / weight / teacher once:
// 4+4+4+4 B read, 4+4+4 B written (+2+2 B bf16 shadows) per parameter instead of two sweeps.
// ema[ti] == null: a parameter without a teacher copy (auxiliary heads).
__global__ void __launch_bounds__(256)
sgd_ema_multi_kernel(S4TensorTable t, void* const* shadow, void* const* ema, void* const* ema_shadow,
                     const float* __restrict__ ema_m, float mu, float wd, int first_step) {
  const int ti = t.chunk_tensor[blockIdx.x];
  const long long off = t.chunk_off[blockIdx.x];
  const long long n = min((long long)S4_CHUNK, t.size[ti] - off);
  float* p = (float*)t.a[ti] + off;
  const float* g = (const float*)t.b[ti] + off;
  float* buf = (float*)t.c[ti] + off;
  __nv_bfloat16* sh = (shadow && shadow[ti]) ? (__nv_bfloat16*)shadow[ti] + off : nullptr;
  float* e = ema[ti] ? (float*)ema[ti] + off : nullptr;
  __nv_bfloat16* esh = (e && ema_shadow && ema_shadow[ti]) ? (__nv_bfloat16*)ema_shadow[ti] + off : nullptr;
  const float lr = t.scalar[ti];
  const float m = e ? ema_m[ti] : 0.f, om = 1.f - m;
  auto upd = [&](float gi, float& pi, float& bi) {
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    bi = first_step ? gi : fmaf(mu, bi, gi);
    pi = fmaf(-lr, bi, pi);
  };
  auto pack4 = [](const float4& v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    return o;
  };
  const bool vec = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)buf) | ((uintptr_t)e)) & 15) == 0 &&
                   ((((uintptr_t)sh) | ((uintptr_t)esh)) & 7) == 0;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 bv = first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(buf)[i];
    float4 ev = e ? reinterpret_cast<float4*>(e)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    upd(gv.x, pv.x, bv.x);
    upd(gv.y, pv.y, bv.y);
    upd(gv.z, pv.z, bv.z);
    upd(gv.w, pv.w, bv.w);
    reinterpret_cast<float4*>(buf)[i] = bv;
    reinterpret_cast<float4*>(p)[i] = pv;
    if (sh) reinterpret_cast<uint2*>(sh)[i] = pack4(pv);
    if (e) {
      ev.x = fmaf(pv.x, om, ev.x * m);
      ev.y = fmaf(pv.y, om, ev.y * m);
      ev.z = fmaf(pv.z, om, ev.z * m);
      ev.w = fmaf(pv.w, om, ev.w * m);
      reinterpret_cast<float4*>(e)[i] = ev;
      if (esh) reinterpret_cast<uint2*>(esh)[i] = pack4(ev);
    }
  }
  for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
    float pi = p[i];
    float bi = first_step ? 0.f : buf[i];
    upd(g[i], pi, bi);
    buf[i] = bi;
    p[i] = pi;
    if (sh) sh[i] = __float2bfloat16_rn(pi);
    if (e) {
      const float ei = fmaf(pi, om, e[i] * m);
      e[i] = ei;
      if (esh) esh[i] = __float2bfloat16_rn(ei);
    }
  }
}

extern "C" int s4_sgd_ema_multi_tensor(void* const* params, void* const* grads, void* const* bufs,
                                       void* const* bf16_shadow, void* const* ema_params,
                                       void* const* ema_bf16_shadow, const float* ema_momentum,
                                       const long long* sizes, const float* lrs, const int* chunk_tensor,
                                       const long long* chunk_off, int n_chunks, float momentum,
                                       float weight_decay, int first_step, cudaStream_t stream) {
  S4ProfScope prof_("sgd_ema_multi_tensor", 0.0, 1, stream);
  if (n_chunks == 0) return S4_OK;
  S4_REQUIRE(ema_params != nullptr && ema_momentum != nullptr, "sgd_ema: null EMA table");
  S4TensorTable t{params, grads, bufs, sizes, lrs, chunk_tensor, chunk_off};
  sgd_ema_multi_kernel<<<n_chunks, 256, 0, stream>>>(t, bf16_shadow, ema_params, ema_bf16_shadow, ema_momentum,
                                                     momentum, weight_decay, first_step);
  return s4_check_launch("sgd_ema_multi_tensor");
}

extern "C" int s4_chunk_elems() { return S4_CHUNK; }

// ------------------------------------------------------------------------------------------
// CutMix and PatchShuffle (reference generate_unsup_data.py:400-453, :737-819).  Boxes and
// permutations come from the HOST RNG (parity: Appendix B-7) as small device arrays.
// ------------------------------------------------------------------------------------------
// boxes[b] = (y0, y1, x0, x1): inside the box the neighbour (b+1)%B is copied.
__global__ void cutmix_kernel(const float* __restrict__ img, const long long* __restrict__ lab,
                              const int* __restrict__ boxes, float* __restrict__ out_img,
                              long long* __restrict__ out_lab, int B, int C, int H, int W) {
  const size_t plane = (size_t)H * W;
  const size_t total = (size_t)B * plane;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / plane);
    const int* bx = boxes + b * 4;
    const bool in = y >= bx[0] && y < bx[1] && x >= bx[2] && x < bx[3];
    const int sb = in ? (b + 1) % B : b;
    const size_t r = (size_t)y * W + x;
    for (int c = 0; c < C; ++c)
      out_img[((size_t)b * C + c) * plane + r] = __ldg(img + ((size_t)sb * C + c) * plane + r);
    if (lab) out_lab[(size_t)b * plane + r] = lab[(size_t)sb * plane + r];
  }
}

extern "C" int s4_cutmix(const float* img, const long long* label, const int* boxes_dev,
                         float* out_img, long long* out_label, int B, int C, int H, int W,
                         cudaStream_t stream) {
  S4ProfScope prof_("cutmix", 0.0, 1, stream);
  const size_t total = (size_t)B * H * W;
  if (total == 0) return S4_OK;
  const int grid = (int)min((total + 255) / 256, (size_t)s4_num_sms() * 16);
  cutmix_kernel<<<grid, 256, 0, stream>>>(img, label, boxes_dev, out_img, out_label, B, C, H, W);
  return s4_check_launch("cutmix");
}

// out_block[p] = in_block[perm[b, p]] on block x block pixel tiles, blocks row-major.
__global__ void patchshuffle_kernel(const float* __restrict__ img, const long long* __restrict__ perm,
                                    float* __restrict__ out, int B, int C, int H, int W, int block) {
  const size_t plane = (size_t)H * W;
  const size_t total = (size_t)B * C * plane;
  const int gw = W / block;
  const int nblk = gw * (H / block);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const size_t bc = i / plane;
    const int b = (int)(bc / C);
    const int p = (y / block) * gw + (x / block);
    const int s = (int)perm[(size_t)b * nblk + p];
    const int sy = (s / gw) * block + (y % block), sx = (s % gw) * block + (x % block);
    out[i] = __ldg(img + bc * plane + (size_t)sy * W + sx);
  }
}

extern "C" int s4_patchshuffle(const float* img, const long long* perm_dev, float* out, int B,
                               int C, int H, int W, int block, cudaStream_t stream) {
  S4ProfScope prof_("patchshuffle", 0.0, 1, stream);
  S4_REQUIRE(block > 0 && H % block == 0 && W % block == 0, "patchshuffle: H,W must be multiples of block");
  const size_t total = (size_t)B * C * H * W;
  if (total == 0) return S4_OK;
  const int grid = (int)min((total + 255) / 256, (size_t)s4_num_sms() * 16);
  patchshuffle_kernel<<<grid, 256, 0, stream>>>(img, perm_dev, out, B, C, H, W, block);
  return s4_check_launch("patchshuffle");
}
