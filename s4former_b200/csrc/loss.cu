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

extern "C" int s4_pseudo_label(const float* logits, long long* hard, long long* conf, float* u,
                               int B, int C, int H, int W, int patch, float threshold,
                               cudaStream_t stream) {
  S4ProfScope prof_("pseudo_label", 0.0, 1, stream);
  S4_REQUIRE(C >= 1 && C <= S4_MAXC, "pseudo_label: C=%d not in [1,%d]", C, S4_MAXC);
  S4_REQUIRE(patch == 8 || patch == 16, "pseudo_label: patch must be 8 or 16 (got %d)", patch);
  S4_REQUIRE(H % 16 == 0 && W % 32 == 0, "pseudo_label: H%%16, W%%32 required (H=%d W=%d)", H, W);
  if (B == 0) return S4_OK;
  dim3 grid(W / 32, H / 16, B), block(32, 16);
  pseudo_label_kernel<<<grid, block, 0, stream>>>(logits, hard, conf, u, C, H, W, patch, threshold);
  return s4_check_launch("pseudo_label");
}

// ------------------------------------------------------------------------------------------
// masked CE + NCR, forward + gradient in one pass
// ------------------------------------------------------------------------------------------
// partial[blk*3 + {0,1,2}] = sum nll, sum ncr distance, #valid for that block
__global__ void __launch_bounds__(256)
ce_ncr_kernel(const float* __restrict__ zs, const float* __restrict__ zt,
              const long long* __restrict__ label, float* __restrict__ dz,
              float* __restrict__ partial, int C, size_t plane, size_t npix, float ce_scale,
              float ncr_scale, int ignore_index, const float* __restrict__ gscale) {
  __shared__ float red[32];
  if (gscale) {   // upstream gradients of (loss_ce, loss_ncr), device resident: no host sync
    ce_scale *= gscale[0];
    ncr_scale *= gscale[1];
  }
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float nll = 0.f, dist = 0.f, nvalid = 0.f;
  if (pix < npix) {
    const size_t b = pix / plane, r = pix % plane;
    const float* sp = zs + b * C * plane + r;
    float* gp = dz ? dz + b * C * plane + r : nullptr;
    const long long y = label[pix];
    const bool valid = (y != ignore_index) && y >= 0 && y < C;
    float v[S4_MAXC], g[S4_MAXC];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < S4_MAXC; ++c)
      if (c < C) { v[c] = __ldg(sp + c * plane); m = fmaxf(m, v[c]); g[c] = 0.f; }
    if (valid) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < S4_MAXC; ++c)
        if (c < C) s += expf(v[c] - m);
      const float logs = logf(s);
      const float inv = 1.f / s;
      nvalid = 1.f;
#pragma unroll
      for (int c = 0; c < S4_MAXC; ++c)
        if (c < C) {
          if (c == (int)y) nll = -(v[c] - m - logs);
          g[c] = ce_scale * (expf(v[c] - m) * inv - (c == (int)y ? 1.f : 0.f));
        }
      if (zt != nullptr) {
        // softmax over the C-1 negative classes, student and teacher
        const float* tp = zt + b * C * plane + r;
        float t[S4_MAXC];
        float ms = -INFINITY, mt = -INFINITY;
#pragma unroll
        for (int c = 0; c < S4_MAXC; ++c)
          if (c < C) {
            t[c] = __ldg(tp + c * plane);
            if (c != (int)y) { ms = fmaxf(ms, v[c]); mt = fmaxf(mt, t[c]); }
          }
        float ss = 0.f, st = 0.f;
#pragma unroll
        for (int c = 0; c < S4_MAXC; ++c)
          if (c < C && c != (int)y) { ss += expf(v[c] - ms); st += expf(t[c] - mt); }
        const float is = 1.f / ss, it = 1.f / st;
        float sq = 0.f;
#pragma unroll
        for (int c = 0; c < S4_MAXC; ++c)
          if (c < C && c != (int)y) {
            const float p = expf(v[c] - ms) * is;
            const float q = expf(t[c] - mt) * it;
            const float d = p - q + 1e-6f;     // torch PairwiseDistance eps, inside the norm
            sq += d * d;
            v[c] = p;                          // reuse registers: v <- p, t <- d
            t[c] = d;
          }
        dist = sqrtf(sq);
        const float ir = 1.f / dist;
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < S4_MAXC; ++c)
          if (c < C && c != (int)y) dot += t[c] * ir * v[c];
#pragma unroll
        for (int c = 0; c < S4_MAXC; ++c)
          if (c < C && c != (int)y) g[c] += ncr_scale * v[c] * (t[c] * ir - dot);
      }
    }
    if (gp) {
#pragma unroll
      for (int c = 0; c < S4_MAXC; ++c)
        if (c < C) gp[c * plane] = g[c];
    }
  }
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
// (device pointer to two floats, or null for (1,1)).
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
  ce_ncr_kernel<<<nblk, 256, 0, stream>>>(logits_s, logits_t, label, dlogits, (float*)workspace, C,
                                          (size_t)H * W, npix, ce_weight / P, ncr_weight / P,
                                          ignore_index, grad_scale);
  if (loss_out) s4_count_launches(1);
  if (loss_out)
    ce_ncr_finalize_kernel<<<1, 256, 0, stream>>>((const float*)workspace, nblk, loss_out,
                                                  ce_weight / P, ncr_weight / P);
  return s4_check_launch("ce_ncr");
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
