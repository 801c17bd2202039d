// HBM-bound per-pixel kernels of the S4Former train step:
//   * teacher pseudo-label: softmax-max-threshold -> hard label / confidence mask / per-patch
//     unconfidence  (reference encoder_decoder.py:888-901, :541-542, :547-555)
//   * masked cross-entropy + negative-class-ranking, forward and backward in one pass
//     (reference cross_entropy_loss.py:45-61, losses/utils.py:65-69,
//      encoder_decoder.py:906-954)
// Logits are NCHW fp32 (the reference-facing layout); threads walk x so every class plane is
// read with 128-byte coalesced warps.  Arithmetic order follows ATen's softmax
// (max, then sequential fp32 sum of expf(x-max), IEEE division) so the >thr comparison is
// reproducible.
#include "common.cuh"

#define S4_MAXC 32

// ------------------------------------------------------------------------------------------
// pseudo label
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
pseudo_label_kernel(const float* __restrict__ z, long long* __restrict__ hard,
                    long long* __restrict__ conf, float* __restrict__ u, int C, int H, int W,
                    int patch, float thr) {
  __shared__ int cnt[2 * 4];  // up to (16/8) x (32/8) patch cells
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * 32 + tx;
  if (tid < 8) cnt[tid] = 0;
  __syncthreads();
  const int x = blockIdx.x * 32 + tx, y = blockIdx.y * 16 + ty, b = blockIdx.z;
  const size_t plane = (size_t)H * W;
  const float* zp = z + (size_t)b * C * plane + (size_t)y * W + x;
  float v[S4_MAXC];
  float m = -INFINITY;
  int arg = 0;
#pragma unroll
  for (int c = 0; c < S4_MAXC; ++c) {
    if (c < C) {
      v[c] = __ldg(zp + c * plane);
      if (v[c] > m) { m = v[c]; arg = c; }   // first index wins on ties
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < S4_MAXC; ++c)
    if (c < C) s += expf(v[c] - m);
  const float pmax = 1.0f / s;               // expf(0)/sum
  const int confident = pmax > thr;
  const size_t o = (size_t)b * plane + (size_t)y * W + x;
  hard[o] = confident ? (long long)arg : 255ll;
  conf[o] = confident;
  const int cells_x = 32 / patch;
  if (!confident) atomicAdd(&cnt[(ty / patch) * cells_x + (tx / patch)], 1);
  __syncthreads();
  const int cells_y = 16 / patch;
  if (tid < cells_x * cells_y) {
    const int cy = tid / cells_x, cx = tid % cells_x;
    const int gh = H / patch, gw = W / patch;
    const int py = blockIdx.y * cells_y + cy, px = blockIdx.x * cells_x + cx;
    u[((size_t)b * gh + py) * gw + px] = (float)cnt[tid] / (float)(patch * patch);
  }
}

// Two pixels per thread (W % 64 == 0): every class plane is read as 8-byte vectors, a block row
// covers 256 contiguous bytes of each plane (a 32-thread row of the scalar kernel touches 128 B
// at a 2 KB stride: poor DRAM page locality, ~1 TB/s).  Arithmetic per pixel is unchanged.
template <int MAXC>
__global__ void __launch_bounds__(512)
pseudo_label_vec2_kernel(const float* __restrict__ z, long long* __restrict__ hard,
                         long long* __restrict__ conf, float* __restrict__ u, int C, int H, int W,
                         int patch, float thr) {
  __shared__ int cnt[2 * 8];  // up to (16/8) x (64/8) patch cells
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * 32 + tx;
  if (tid < 16) cnt[tid] = 0;
  __syncthreads();
  const int x = blockIdx.x * 64 + tx * 2, y = blockIdx.y * 16 + ty, b = blockIdx.z;
  const size_t plane = (size_t)H * W;
  const float* zp = z + (size_t)b * C * plane + (size_t)y * W + x;
  float2 v[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
    if (c < C) v[c] = __ldg(reinterpret_cast<const float2*>(zp + c * plane));
  long long hv[2], cv[2];
  int unconf = 0;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    float m = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) {
        const float val = e == 0 ? v[c].x : v[c].y;
        if (val > m) { m = val; arg = c; }   // first index wins on ties
      }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) s += expf((e == 0 ? v[c].x : v[c].y) - m);
    const float pmax = 1.0f / s;               // expf(0)/sum
    const int confident = pmax > thr;
    hv[e] = confident ? (long long)arg : 255ll;
    cv[e] = confident;
    unconf += confident ? 0 : 1;
  }
  const size_t o = (size_t)b * plane + (size_t)y * W + x;
  *reinterpret_cast<longlong2*>(hard + o) = make_longlong2(hv[0], hv[1]);
  *reinterpret_cast<longlong2*>(conf + o) = make_longlong2(cv[0], cv[1]);
  const int cells_x = 64 / patch;              // 2 consecutive pixels never straddle a cell
  if (unconf) atomicAdd(&cnt[(ty / patch) * cells_x + (tx * 2) / patch], unconf);
  __syncthreads();
  const int cells_y = 16 / patch;
  if (tid < cells_x * cells_y) {
    const int cy = tid / cells_x, cx = tid % cells_x;
    const int gh = H / patch, gw = W / patch;
    const int py = blockIdx.y * cells_y + cy, px = blockIdx.x * cells_x + cx;
    u[((size_t)b * gh + py) * gw + px] = (float)cnt[tid] / (float)(patch * patch);
  }
}

extern "C" int s4_pseudo_label(const float* logits, long long* hard, long long* conf, float* u,
                               int B, int C, int H, int W, int patch, float threshold,
                               cudaStream_t stream) {
  S4ProfScope prof_("pseudo_label", 0.0, 1, stream);
  S4_REQUIRE(C >= 1 && C <= S4_MAXC, "pseudo_label: C=%d not in [1,%d]", C, S4_MAXC);
  S4_REQUIRE(patch == 8 || patch == 16, "pseudo_label: patch must be 8 or 16 (got %d)", patch);
  S4_REQUIRE(H % 16 == 0 && W % 32 == 0, "pseudo_label: H%%16, W%%32 required (H=%d W=%d)", H, W);
  if (B == 0) return S4_OK;
  dim3 grid(W / 32, H / 16, B), block(32, 16);
  if (W % 64 == 0 && (((uintptr_t)logits | (uintptr_t)hard | (uintptr_t)conf) & 15) == 0) {
    grid.x = W / 64;
    if (C <= 24)
      pseudo_label_vec2_kernel<24><<<grid, block, 0, stream>>>(logits, hard, conf, u, C, H, W, patch, threshold);
    else
      pseudo_label_vec2_kernel<S4_MAXC><<<grid, block, 0, stream>>>(logits, hard, conf, u, C, H, W, patch, threshold);
    return s4_check_launch("pseudo_label");
  }
  pseudo_label_kernel<<<grid, block, 0, stream>>>(logits, hard, conf, u, C, H, W, patch, threshold);
  return s4_check_launch("pseudo_label");
}

// ------------------------------------------------------------------------------------------
// masked CE + NCR, forward + gradient in one pass
// ------------------------------------------------------------------------------------------
// partial[blk*3 + {0,1,2}] = sum nll, sum ncr distance, #valid for that block.
// One pixel per thread, the C class planes are read with coalesced 128-byte warps and all C loads
// of a pixel are in flight together.  exp / log / divide use the fast intrinsics (ex2.approx,
// lg2.approx, rcp.approx: ~1e-6 relative, far inside the 1e-3 loss gate); the bit-exactness
// contract (softmax-max > threshold) lives in pseudo_label_kernel, not here.
// WRITE_DZ: also emit d(loss)/dz_s; HAS_T: NCR term against the teacher logits.
// SKIP_IF_EQUAL: gradient fix-up launch -- returns at once unless gscale[0] != gscale[1]
// (see s4_ce_ncr_grad_fixup).
// CC: compile-time class count (0 = run-time C <= S4_MAXC, every loop predicated): the kernel is
// ISSUE-bound (ncu: 70 % issue utilisation, 1100 thread-instructions per pixel with the 32-wide
// predicated loops), so the shipped class counts (21 VOC, 19 Cityscapes) get exact-trip loops.
template <bool WRITE_DZ, bool HAS_T, bool SKIP_IF_EQUAL, int CC>
__global__ void __launch_bounds__(256, HAS_T ? 3 : 4)   // NCR: <= 85 registers (3 blocks/SM); CE only: <= 64 (4)
ce_ncr_kernel(const float* __restrict__ zs, const float* __restrict__ zt,
              const long long* __restrict__ label, float* __restrict__ dz,
              float* __restrict__ partial, int C, size_t plane, size_t npix, float ce_scale,
              float ncr_scale, int ignore_index, const float* __restrict__ gscale) {
  constexpr int NCLS = CC ? CC : S4_MAXC;
  __shared__ float red[32];
  if (gscale) {   // upstream gradients of (loss_ce, loss_ncr), device resident: no host sync
    const float g0 = gscale[0], g1 = gscale[1];
    if (SKIP_IF_EQUAL && g0 == g1) return;
    ce_scale *= g0;
    ncr_scale *= g1;
  }
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float nll = 0.f, dist = 0.f, nvalid = 0.f;
  if (pix < npix) {
    const size_t b = pix / plane, r = pix % plane;
    const float* sp = zs + b * C * plane + r;
    const long long y = label[pix];
    const bool valid = (y != ignore_index) && y >= 0 && y < C;
    const int yi = valid ? (int)y : -1;
    float v[NCLS], t[NCLS];
    float m = -INFINITY;      // max over all classes
    float ms = -INFINITY;     // max over the negative classes (c != label)
#pragma unroll
    for (int c = 0; c < NCLS; ++c)
      if (CC || c < C) {
        v[c] = __ldg(sp + c * plane);
        m = fmaxf(m, v[c]);
        if (c != yi) ms = fmaxf(ms, v[c]);
      }
    float mt = -INFINITY;
    if (HAS_T) {
      const float* tp = zt + b * C * plane + r;
#pragma unroll
      for (int c = 0; c < NCLS; ++c)
        if (CC || c < C) {
          t[c] = __ldg(tp + c * plane);
          if (c != yi) mt = fmaxf(mt, t[c]);
        }
    }
    float* gp = WRITE_DZ ? dz + b * C * plane + r : nullptr;
    if (valid) {
      nvalid = 1.f;
      if (!HAS_T) {
        float vy = 0.f, s = 0.f;
#pragma unroll
        for (int c = 0; c < NCLS; ++c)
          if (CC || c < C) {
            if (c == yi) vy = v[c];
            v[c] = __expf(v[c] - m);
            s += v[c];
          }
        nll = m + __logf(s) - vy;
        if (WRITE_DZ) {
          const float k = ce_scale * __fdividef(1.f, s);
#pragma unroll
          for (int c = 0; c < NCLS; ++c)
            if (CC || c < C) gp[c * plane] = fmaf(v[c], k, c == yi ? -ce_scale : 0.f);
        }
      } else {
        // negatives are exponentiated against THEIR max (f_c = exp(v_c - ms), the NCR softmax of
        // the C-1 negative classes is f_c / F exactly as the reference forms it); the CE softmax
        // over all classes follows from s = F exp(ms - m) + exp(v_y - m) with both exponents <= 0.
        float vy = 0.f, F = 0.f, st = 0.f;
#pragma unroll
        for (int c = 0; c < NCLS; ++c)
          if (CC || c < C) {
            if (c == yi) {
              vy = v[c];
            } else {
              v[c] = __expf(v[c] - ms);
              F += v[c];
              t[c] = __expf(t[c] - mt);
              st += t[c];
            }
          }
        const float a = __expf(ms - m), ey = __expf(vy - m);
        const float s = fmaf(F, a, ey);
        nll = m + __logf(s) - vy;
        const float inv = __fdividef(1.f, s);
        const float iF = __fdividef(1.f, F), it = __fdividef(1.f, st);
        float sq = 0.f;
        // t[c] <- d_c = p_c - q_c + eps   (torch PairwiseDistance eps, inside the norm)
#pragma unroll
        for (int c = 0; c < NCLS; ++c)
          if ((CC || c < C) && c != yi) {
            const float d = fmaf(v[c], iF, fmaf(-t[c], it, 1e-6f));
            sq = fmaf(d, d, sq);
            t[c] = d;
          }
        dist = sqrtf(sq);
        if (WRITE_DZ) {
          const float ir = __fdividef(1.f, dist);
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < NCLS; ++c)
            if ((CC || c < C) && c != yi) dot = fmaf(t[c] * ir, v[c] * iF, dot);
          const float k = ce_scale * inv * a;
#pragma unroll
          for (int c = 0; c < NCLS; ++c)
            if (CC || c < C) {
              float g;
              if (c == yi) g = ce_scale * (ey * inv - 1.f);
              else g = fmaf(ncr_scale * (v[c] * iF), fmaf(t[c], ir, -dot), v[c] * k);
              gp[c * plane] = g;
            }
        }
      }
    } else if (WRITE_DZ) {
#pragma unroll
      for (int c = 0; c < NCLS; ++c)
        if (CC || c < C) gp[c * plane] = 0.f;
    }
  }
  if (SKIP_IF_EQUAL) return;      // fix-up launches do not touch the loss partials
  const float a = block_sum(nll, red);
  const float bsum = block_sum(dist, red);
  const float cnt = block_sum(nvalid, red);
  if (threadIdx.x == 0) {
    partial[blockIdx.x * 3 + 0] = a;
    partial[blockIdx.x * 3 + 1] = bsum;
    partial[blockIdx.x * 3 + 2] = cnt;
  }
}

// deterministic final reduce: out[0]=ce_scale*sum nll, out[1]=ncr_scale*sum dist, out[2]=#valid
__global__ void ce_ncr_finalize_kernel(const float* __restrict__ partial, int nblk,
                                       float* __restrict__ out, float ce_scale, float ncr_scale) {
  __shared__ double sh[3][256];
  double a = 0, b = 0, c = 0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
    a += partial[i * 3 + 0];
    b += partial[i * 3 + 1];
    c += partial[i * 3 + 2];
  }
  sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b; sh[2][threadIdx.x] = c;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = (float)(sh[0][0] * (double)ce_scale);
    out[1] = (float)(sh[1][0] * (double)ncr_scale);
    out[2] = (float)sh[2][0];
  }
}

extern "C" size_t s4_ce_ncr_workspace(int B, int H, int W) {
  const size_t npix = (size_t)B * H * W;
  return ((npix + 255) / 256) * 3 * sizeof(float);
}

// loss_out[0] = ce_weight/P * sum_valid nll ; loss_out[1] = ncr_weight/P * sum_valid dist ;
// loss_out[2] = number of valid pixels (loss_out may be null: gradient-only call).
// dlogits (may be null) receives g0*d(loss0)/dz_s + g1*d(loss1)/dz_s with (g0,g1) = grad_scale
// (device pointer to two floats, or null for (1,1)).  Losses and gradient may be requested in the
// same call: the forward of a training step does, so the logits are streamed once.
extern "C" int s4_ce_ncr(const float* logits_s, const float* logits_t, const long long* label,
                         float* dlogits, float* loss_out, const float* grad_scale, int B, int C,
                         int H, int W, float ce_weight, float ncr_weight, int ignore_index,
                         void* workspace, size_t ws_bytes, cudaStream_t stream) {
  S4ProfScope prof_("ce_ncr", 0.0, 1, stream);
  S4_REQUIRE(C >= 1 && C <= S4_MAXC, "ce_ncr: C=%d not in [1,%d]", C, S4_MAXC);
  const size_t npix = (size_t)B * H * W;
  S4_REQUIRE(npix > 0, "ce_ncr: empty input");
  S4_REQUIRE(ws_bytes >= s4_ce_ncr_workspace(B, H, W), "ce_ncr: workspace too small");
  const int nblk = (int)((npix + 255) / 256);
  const float P = (float)npix;
  const size_t plane = (size_t)H * W;
  const float cs = ce_weight / P, ns = ncr_weight / P;
  float* part = (float*)workspace;
  const bool has_t = logits_t != nullptr && ncr_weight != 0.f;
#define S4_CE_C(WD, HT, CCV)                                                                          \
  ce_ncr_kernel<WD, HT, false, CCV><<<nblk, 256, 0, stream>>>(logits_s, logits_t, label, dlogits, part, C, \
                                                              plane, npix, cs, ns, ignore_index, grad_scale)
#define S4_CE(WD, HT)                       \
  do {                                      \
    if (C == 21) S4_CE_C(WD, HT, 21);       \
    else if (C == 19) S4_CE_C(WD, HT, 19);  \
    else S4_CE_C(WD, HT, 0);                \
  } while (0)
  if (dlogits) {
    if (has_t) S4_CE(true, true); else S4_CE(true, false);
  } else {
    if (has_t) S4_CE(false, true); else S4_CE(false, false);
  }
#undef S4_CE
#undef S4_CE_C
  if (loss_out) s4_count_launches(1);
  if (loss_out)
    ce_ncr_finalize_kernel<<<1, 256, 0, stream>>>((const float*)workspace, nblk, loss_out, cs, ns);
  return s4_check_launch("ce_ncr");
}

// dlogits *= g when the two upstream gradients are equal (the common case: the step's total loss
// is a plain sum, g = (1, 1), and nothing is touched); recomputed from the logits when they differ.
// Everything is decided on the device from grad_scale: no host synchronisation.
__global__ void ce_ncr_rescale_kernel(float* __restrict__ y, const float* __restrict__ gscale,
                                      int use_second, size_t n4, size_t n) {
  const float s = gscale[0];
  if (use_second && gscale[1] != s) return;     // the recompute launch handles it
  if (s == 1.f) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float4* y4 = reinterpret_cast<float4*>(y);
  for (size_t k = i; k < n4; k += stride) {
    float4 v = y4[k];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    y4[k] = v;
  }
  for (size_t k = n4 * 4 + i; k < n; k += stride) y[k] *= s;
}

extern "C" int s4_ce_ncr_grad_fixup(const float* logits_s, const float* logits_t,
                                    const long long* label, float* dlogits, const float* grad_scale,
                                    int B, int C, int H, int W, float ce_weight, float ncr_weight,
                                    int ignore_index, cudaStream_t stream) {
  S4ProfScope prof_("ce_ncr_fixup", 0.0, 1, stream);
  S4_REQUIRE(C >= 1 && C <= S4_MAXC, "ce_ncr: C=%d not in [1,%d]", C, S4_MAXC);
  S4_REQUIRE(grad_scale != nullptr && dlogits != nullptr, "ce_ncr_grad_fixup: null argument");
  const size_t npix = (size_t)B * H * W;
  if (npix == 0) return S4_OK;
  const bool has_t = logits_t != nullptr && ncr_weight != 0.f;
  const size_t n = npix * C;
  const int grid = s4_num_sms() * 8;
  ce_ncr_rescale_kernel<<<grid, 256, 0, stream>>>(dlogits, grad_scale, has_t ? 1 : 0, n / 4, n);
  if (has_t) {
    s4_count_launches(1);
    const int nblk = (int)((npix + 255) / 256);
    const float P = (float)npix;
    ce_ncr_kernel<true, true, true, 0><<<nblk, 256, 0, stream>>>(
        logits_s, logits_t, label, dlogits, nullptr, C, (size_t)H * W, npix, ce_weight / P,
        ncr_weight / P, ignore_index, grad_scale);
  }
  return s4_check_launch("ce_ncr_grad_fixup");
}

// y[i] *= *scale   (upstream gradient of a scalar loss applied to a saved gradient)
__global__ void scale_by_scalar_kernel(float* __restrict__ y, const float* __restrict__ scale,
                                       size_t n4, size_t n) {
  const float s = *scale;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float4* y4 = reinterpret_cast<float4*>(y);
  for (size_t k = i; k < n4; k += stride) {
    float4 v = y4[k];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    y4[k] = v;
  }
  for (size_t k = n4 * 4 + i; k < n; k += stride) y[k] *= s;
}

extern "C" int s4_scale_by_scalar(float* y, const float* scale_dev, size_t n, cudaStream_t stream) {
  S4ProfScope prof_("scale_by_scalar", 0.0, 1, stream);
  if (n == 0) return S4_OK;
  const int grid = s4_num_sms() * 8;
  scale_by_scalar_kernel<<<grid, 256, 0, stream>>>(y, scale_dev, n / 4, n);
  return s4_check_launch("scale_by_scalar");
}
