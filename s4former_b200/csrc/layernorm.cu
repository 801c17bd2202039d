// LayerNorm forward/backward over the last dim (D <= 1024, D % 8 == 0), fp32 statistics.
// One warp per row, 16-byte vector loads.  An optional row map gathers source rows, which folds
// the backbone feature tap (drop the cls token, reference vit.py:556-562) and the PatchMix
// feature un-shuffle (decode_head.py:186-212) into the SETR head's LayerNorm
// (setr_up_head.py:96-103).  Reference for the op itself: vit.py:119-120 (eps 1e-6).
#include "common.cuh"

#define LN_MAX_VPL 8  // vectors per lane: D <= 32 lanes * 8 vec * 4 (f32) = 1024

// parameters of the columns a lane owns (vector k of the lane = columns (lane + 32k) * VN ...)
template <int VN>
__device__ __forceinline__ void load_cols(const float* __restrict__ p, int vi, float (&out)[VN]) {
#pragma unroll
  for (int q = 0; q < VN / 4; ++q) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(p + vi * VN) + q);
    out[4 * q] = f.x; out[4 * q + 1] = f.y; out[4 * q + 2] = f.z; out[4 * q + 3] = f.w;
  }
}

// pull a 16-byte chunk's line into L1 ahead of the iteration that loads it: a warp keeps only one
// row in registers, so without this every row pays the full memory latency in sequence
__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

template <typename T, int VPL>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const T* __restrict__ x, const int* __restrict__ row_map,
              const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ y,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int D,
              float eps) {
  constexpr int VN = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const int nvec = D / VN;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float gm[VPL][VN], bt[VPL][VN];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
      load_cols<VN>(gamma, vi, gm[k]);
      load_cols<VN>(beta, vi, bt[k]);
    }
  }
  const float inv_d = 1.f / (float)D;
  for (int row = warp0; row < rows; row += nwarps) {
    const int src = row_map ? row_map[row] : row;
    const T* xr = x + (size_t)src * D;
    Vec16<T> v[VPL];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (vi < nvec) v[k].load(xr + vi * VN);
    }
    if (row + nwarps < rows) {
      const int nsrc = row_map ? row_map[row + nwarps] : row + nwarps;
      const T* nx = x + (size_t)nsrc * D;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vi = lane + k * 32;
        if (vi < nvec) prefetch_l1(nx + vi * VN);
      }
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (vi < nvec) {
#pragma unroll
        for (int e = 0; e < VN; ++e) sum += v[k].get(e);
      }
    }
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (vi < nvec) {
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const float d = v[k].get(e) - mean;
          sq = fmaf(d, d, sq);
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    T* yr = y + (size_t)row * D;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (vi < nvec) {
        Vec16<T> o;
#pragma unroll
        for (int e = 0; e < VN; ++e) o.set(e, fmaf((v[k].get(e) - mean) * rstd, gm[k][e], bt[k][e]));
        o.store(yr + vi * VN);
      }
    }
  }
}

// Variant for many rows (the encoder's [B*L, D] launches): gamma / beta are NOT held in registers
// (2 x VPL x VN = 48 of the ~100 registers of the kernel above) but re-read through L1 for every
// row, so 3x more warps are resident and hide the ~1 us row round trip through L2.
// EXACT: D == VPL * 32 * VN (e.g. 768 = 3 x 32 x 8 bf16): every `vi < nvec` test is dropped at
// compile time (they cost ~20 % of the instructions of this issue-bound kernel: predicates and
// BSSY/BSYNC reconvergence pairs around each guarded group).
template <typename T, int VPL, bool EXACT>
__global__ void __launch_bounds__(128, 10)
ln_fwd_lean_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean_out,
                   float* __restrict__ rstd_out, int rows, int D, float eps) {
  constexpr int VN = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const int nvec = D / VN;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float inv_d = 1.f / (float)D;
  for (int row = warp0; row < rows; row += nwarps) {
    const T* xr = x + (size_t)row * D;
    Vec16<T> v[VPL];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (EXACT || vi < nvec) v[k].load(xr + vi * VN);
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (EXACT || vi < nvec) {
#pragma unroll
        for (int e = 0; e < VN; ++e) sum += v[k].get(e);
      }
    }
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (EXACT || vi < nvec) {
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const float d = v[k].get(e) - mean;
          sq = fmaf(d, d, sq);
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    T* yr = y + (size_t)row * D;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (EXACT || vi < nvec) {
        float gm[VN], bt[VN];
        load_cols<VN>(gamma, vi, gm);
        load_cols<VN>(beta, vi, bt);
        Vec16<T> o;
#pragma unroll
        for (int e = 0; e < VN; ++e) o.set(e, fmaf((v[k].get(e) - mean) * rstd, gm[e], bt[e]));
        o.store(yr + vi * VN);
      }
    }
  }
}

// dx (written at the mapped source row of a pre-zeroed buffer when row_map != null).
// dgamma/dbeta: every warp keeps register partials over its rows, the block folds them through
// shared memory (plain stores, one slab per warp) and issues ONE global atomic per column.
// RSUM: the column sums of `dres` (the residual-stream gradient the kernel reads anyway) are
// accumulated too - it is the bias gradient of the linear layer that produced the residual branch
// (fc2.bias / out_proj.bias in an encoder layer), which otherwise costs a pass of its own over dres.
// Those partials live in the warp's shared-memory slab (the register file is full: 128 per thread).
template <typename T, int VPL, bool EXACT, bool RSUM>
__global__ void __launch_bounds__(128, 4)
ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const int* __restrict__ row_map,
              const float* __restrict__ gamma, const float* __restrict__ mean,
              const float* __restrict__ rstd, const T* __restrict__ dres, T* __restrict__ dx,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dres_sum,
              int rows, int D) {
  constexpr int VN = Vec16<T>::N;
  constexpr int NS = RSUM ? 3 : 2;
  extern __shared__ float sh[];  // [warps][NS][D]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int nvec = D / VN;
  float* mine = sh + (size_t)wib * NS * D;
  if (RSUM) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (EXACT || vi < nvec) {
#pragma unroll
        for (int e = 0; e < VN; ++e) mine[2 * D + vi * VN + e] = 0.f;
      }
    }
  }
  float gm[VPL][VN];
  float ag[VPL][VN], ab[VPL][VN];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = lane + k * 32;
    if (EXACT || vi < nvec) load_cols<VN>(gamma, vi, gm[k]);
#pragma unroll
    for (int e = 0; e < VN; ++e) { ag[k][e] = 0.f; ab[k][e] = 0.f; }
  }
  const float inv_d = 1.f / (float)D;
  for (int row = blockIdx.x * wpb + wib; row < rows; row += gridDim.x * wpb) {
    const int src = row_map ? row_map[row] : row;
    const T* xr = x + (size_t)src * D;
    const T* gr = dy + (size_t)row * D;
    const float mu = mean[row], rs = rstd[row];
    Vec16<T> xv[VPL], gv[VPL], rv[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (EXACT || vi < nvec) {
        xv[k].load(xr + vi * VN);
        gv[k].load(gr + vi * VN);
        if (dres) rv[k].load(dres + (size_t)src * D + vi * VN);
      }
    }
    {
      const int nrow = row + gridDim.x * wpb;
      if (nrow < rows) {
        const int nsrc = row_map ? row_map[nrow] : nrow;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const int vi = lane + k * 32;
          if (EXACT || vi < nvec) {
            prefetch_l1(x + (size_t)nsrc * D + vi * VN);
            prefetch_l1(dy + (size_t)nrow * D + vi * VN);
            if (dres) prefetch_l1(dres + (size_t)nsrc * D + vi * VN);
          }
        }
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (EXACT || vi < nvec) {
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const float xh = (xv[k].get(e) - mu) * rs;
          const float go = gv[k].get(e);
          const float g = go * gm[k][e];
          s1 += g;
          s2 = fmaf(g, xh, s2);
          ag[k][e] = fmaf(go, xh, ag[k][e]);
          ab[k][e] += go;
        }
      }
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
    T* dr = dx + (size_t)src * D;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      if (EXACT || vi < nvec) {
        Vec16<T> o;
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const float xh = (xv[k].get(e) - mu) * rs;
          const float g = gv[k].get(e) * gm[k][e];
          float val = rs * (g - s1 - xh * s2);
          if (dres) val += rv[k].get(e);
          o.set(e, val);
        }
        o.store(dr + vi * VN);
        if (RSUM) {
          // lane-private columns of the warp-private slab: no synchronisation needed
          float4* acc4 = reinterpret_cast<float4*>(mine + 2 * D + vi * VN);
#pragma unroll
          for (int q = 0; q < VN / 4; ++q) {
            float4 a = acc4[q];
            a.x += rv[k].get(4 * q);
            a.y += rv[k].get(4 * q + 1);
            a.z += rv[k].get(4 * q + 2);
            a.w += rv[k].get(4 * q + 3);
            acc4[q] = a;
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = lane + k * 32;
    if (EXACT || vi < nvec) {
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        mine[vi * VN + e] = ag[k][e];
        mine[D + vi * VN + e] = ab[k][e];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NS * D; i += blockDim.x) {
    float acc = 0.f;
    for (int w2 = 0; w2 < wpb; ++w2) acc += sh[(size_t)w2 * NS * D + i];
    float* dst = i < D ? dgamma + i : (i < 2 * D ? dbeta + (i - D) : dres_sum + (i - 2 * D));
    atomicAdd(dst, acc);
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
  // 4-warp blocks: the kernel holds gamma/beta in registers (~100 regs/thread), small blocks pack
  // more warps per SM; ~5 rows per warp so the next-row prefetch has something to hide behind
  constexpr int FT = 128;
  int blocks = (rows + 3) / 4;
  const int fcap = s4_num_sms() * 5;
  if (blocks > fcap) blocks = fcap;
  const int vpl = (D / vn + 31) / 32;
  if (dtype == S4_BF16 && row_map == nullptr && rows >= 4096) {
    int lb = (rows + 3) / 4;
    const int lcap = s4_num_sms() * 10;
    if (lb > lcap) lb = lcap;
    if (D == vpl * 32 * vn) {
      LN_DISPATCH_VPL(vpl, (ln_fwd_lean_kernel<__nv_bfloat16, VPL, true><<<lb, FT, 0, stream>>>(
          (const __nv_bfloat16*)x, gamma, beta, (__nv_bfloat16*)y, mean, rstd, rows, D, eps)));
    } else {
      LN_DISPATCH_VPL(vpl, (ln_fwd_lean_kernel<__nv_bfloat16, VPL, false><<<lb, FT, 0, stream>>>(
          (const __nv_bfloat16*)x, gamma, beta, (__nv_bfloat16*)y, mean, rstd, rows, D, eps)));
    }
    return s4_check_launch("layernorm_fwd");
  }
  if (dtype == S4_BF16) {
    LN_DISPATCH_VPL(vpl, (ln_fwd_kernel<__nv_bfloat16, VPL><<<blocks, FT, 0, stream>>>(
        (const __nv_bfloat16*)x, row_map, gamma, beta, (__nv_bfloat16*)y, mean, rstd, rows, D, eps)));
  } else {
    LN_DISPATCH_VPL(vpl, (ln_fwd_kernel<float, VPL><<<blocks, FT, 0, stream>>>(
        (const float*)x, row_map, gamma, beta, (float*)y, mean, rstd, rows, D, eps)));
  }
  return s4_check_launch("layernorm_fwd");
}

extern "C" int s4_layernorm_bwd(const void* dy, const void* x, const int* row_map,
                                const float* gamma, const float* mean, const float* rstd,
                                const void* dres, void* dx, float* dgamma, float* dbeta,
                                float* dres_sum, int rows, int D, int dtype, cudaStream_t stream) {
  S4ProfScope prof_("layernorm_bwd", 0.0, 1, stream);
  S4_REQUIRE(dres_sum == nullptr || (dres != nullptr && row_map == nullptr),
             "layernorm_bwd: dres_sum needs dres and no row_map");
  const int vn = dtype == S4_BF16 ? 8 : 4;
  S4_REQUIRE(D % vn == 0 && D / vn <= 32 * LN_MAX_VPL, "layernorm: unsupported D=%d", D);
  if (rows == 0) return S4_OK;
  // 4-warp blocks (the kernel needs ~150 registers per thread: three fit per SM, two 8-warp
  // blocks would not)
  constexpr int BT = 128;
  int blocks = (rows + 3) / 4;
  const int cap = s4_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  const size_t smem = (BT / 32) * (dres_sum ? 3 : 2) * (size_t)D * sizeof(float);
  if (smem > 48 * 1024) {
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(ln_bwd_kernel<__nv_bfloat16, 3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<__nv_bfloat16, 3, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<__nv_bfloat16, 3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<__nv_bfloat16, 3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<__nv_bfloat16, 4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<__nv_bfloat16, 4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<__nv_bfloat16, 4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<__nv_bfloat16, 4, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<float, 6, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<float, 6, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<float, 6, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<float, 6, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<float, 8, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<float, 8, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<float, 8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(ln_bwd_kernel<float, 8, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      attr_done = true;
    }
  }
  const int vpl = (D / vn + 31) / 32;
  const bool exact = D == vpl * 32 * vn;
#define S4_LN_BWD(TT, EX, RS)                                                                  \
  LN_DISPATCH_VPL(vpl, (ln_bwd_kernel<TT, VPL, EX, RS><<<blocks, BT, smem, stream>>>(          \
      (const TT*)dy, (const TT*)x, row_map, gamma, mean, rstd, (const TT*)dres, (TT*)dx, dgamma, \
      dbeta, dres_sum, rows, D)))
#define S4_LN_BWD2(TT, EX) do { if (dres_sum) S4_LN_BWD(TT, EX, true); else S4_LN_BWD(TT, EX, false); } while (0)
  if (dtype == S4_BF16) {
    if (exact) S4_LN_BWD2(__nv_bfloat16, true); else S4_LN_BWD2(__nv_bfloat16, false);
  } else {
    if (exact) S4_LN_BWD2(float, true); else S4_LN_BWD2(float, false);
  }
#undef S4_LN_BWD2
#undef S4_LN_BWD
  return s4_check_launch("layernorm_bwd");
}
