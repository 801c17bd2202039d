// Inference / validation path kernels: the pieces the reference
// runs after the decode head in encoder_decoder.py:1068-1232 (whole / slide inference, rescale to
// the original shape, softmax, flip, argmax) and the mIoU accumulation of
// mmseg/core/evaluation/metrics.py:26-131, on the device.
//
//   s4_resize_bilinear_nchw   mmseg/ops/wrappers.py:8-27 resize(..., mode='bilinear',
//                             align_corners=False) for arbitrary (non-integer) scale, NCHW fp32
//   s4_softmax_argmax_nchw    F.softmax(dim=1) [+ horizontal / vertical flip] and argmax(dim=1);
//                             ATen's arithmetic order (max, expf, sequential fp32 sum, IEEE
//                             division) so probabilities are reproducible; first index wins ties
//   s4_accumulate_nchw        preds += pad(crop_logits) and count += 1 of slide_inference :1089-1093
//   s4_intersect_union        the four class histograms of intersect_and_union (ignore_index
//                             masked), int64 counts
#include "common.cuh"

#define S4_INFER_MAXC 64

__global__ void __launch_bounds__(256)
resize_bilinear_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int IH, int IW,
                            int OH, int OW, float sy, float sx) {
  const size_t total = (size_t)planes * OH * OW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
    const size_t pl = i / ((size_t)OW * OH);
    // ATen area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false)
    float fy = sy * ((float)oy + 0.5f) - 0.5f;
    float fx = sx * ((float)ox + 0.5f) - 0.5f;
    if (fy < 0.f) fy = 0.f;
    if (fx < 0.f) fx = 0.f;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < IH - 1 ? 1 : 0), x1 = x0 + (x0 < IW - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* p = in + pl * (size_t)IH * IW;
    out[i] = hy * (hx * __ldg(p + (size_t)y0 * IW + x0) + lx * __ldg(p + (size_t)y0 * IW + x1)) +
             ly * (hx * __ldg(p + (size_t)y1 * IW + x0) + lx * __ldg(p + (size_t)y1 * IW + x1));
  }
}

extern "C" int s4_resize_bilinear_nchw(const float* in, float* out, int planes, int IH, int IW, int OH,
                                       int OW, cudaStream_t stream) {
  S4ProfScope prof_("resize_bilinear", 0.0, 1, stream);
  const size_t total = (size_t)planes * OH * OW;
  if (total == 0) return S4_OK;
  S4_REQUIRE(IH > 0 && IW > 0, "resize_bilinear: empty input");
  const int grid = (int)min((total + 255) / 256, (size_t)s4_num_sms() * 32);
  resize_bilinear_nchw_kernel<<<grid, 256, 0, stream>>>(in, out, planes, IH, IW, OH, OW, (float)IH / (float)OH,
                                                        (float)IW / (float)OW);
  return s4_check_launch("resize_bilinear");
}

// flip: 0 none, 1 horizontal (x), 2 vertical (y): prob / pred are written at the flipped position
__global__ void __launch_bounds__(256)
softmax_argmax_nchw_kernel(const float* __restrict__ z, float* __restrict__ prob, long long* __restrict__ pred,
                           int C, int H, int W, size_t npix, int flip) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const size_t plane = (size_t)H * W;
  const size_t b = pix / plane, r = pix % plane;
  const float* zp = z + b * C * plane + r;
  float v[S4_INFER_MAXC];
  float m = -INFINITY;
#pragma unroll
  for (int c = 0; c < S4_INFER_MAXC; ++c)
    if (c < C) {
      v[c] = __ldg(zp + c * plane);
      m = fmaxf(m, v[c]);
    }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < S4_INFER_MAXC; ++c)
    if (c < C) {
      v[c] = expf(v[c] - m);
      s += v[c];
    }
  int y = (int)(r / W), x = (int)(r % W);
  if (flip == 1) x = W - 1 - x;
  if (flip == 2) y = H - 1 - y;
  const size_t o = (size_t)y * W + x;
  float best = -1.f;
  int arg = 0;
#pragma unroll
  for (int c = 0; c < S4_INFER_MAXC; ++c)
    if (c < C) {
      const float pc = v[c] / s;
      if (prob) prob[(b * C + c) * plane + o] = pc;
      if (pc > best) { best = pc; arg = c; }      // first index wins ties
    }
  if (pred) pred[b * plane + o] = arg;
}

extern "C" int s4_softmax_argmax_nchw(const float* logits, float* prob, long long* pred, int B, int C, int H,
                                      int W, int flip, cudaStream_t stream) {
  S4ProfScope prof_("softmax_argmax", 0.0, 1, stream);
  S4_REQUIRE(C >= 1 && C <= S4_INFER_MAXC, "softmax_argmax: C=%d not in [1,%d]", C, S4_INFER_MAXC);
  S4_REQUIRE(flip >= 0 && flip <= 2, "softmax_argmax: flip must be 0, 1 or 2");
  const size_t npix = (size_t)B * H * W;
  if (npix == 0) return S4_OK;
  softmax_argmax_nchw_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, stream>>>(logits, prob, pred, C, H, W, npix, flip);
  return s4_check_launch("softmax_argmax");
}

// preds[:, :, y1:y1+ch, x1:x1+cw] += crop ; count[:, 0, y1:y1+ch, x1:x1+cw] += 1
__global__ void __launch_bounds__(256)
accumulate_crop_kernel(const float* __restrict__ crop, float* __restrict__ preds, float* __restrict__ count, int B,
                       int C, int H, int W, int y1, int x1, int ch, int cw) {
  const size_t total = (size_t)B * C * ch * cw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % cw), y = (int)((i / cw) % ch);
    const size_t bc = i / ((size_t)cw * ch);
    const int c = (int)(bc % C);
    const size_t b = bc / C;
    preds[(bc * H + (y1 + y)) * W + (x1 + x)] += crop[i];
    if (c == 0) count[(b * H + (y1 + y)) * W + (x1 + x)] += 1.f;
  }
}

extern "C" int s4_accumulate_crop(const float* crop, float* preds, float* count, int B, int C, int H, int W,
                                  int y1, int x1, int ch, int cw, cudaStream_t stream) {
  S4ProfScope prof_("accumulate_crop", 0.0, 1, stream);
  S4_REQUIRE(y1 >= 0 && x1 >= 0 && y1 + ch <= H && x1 + cw <= W, "accumulate_crop: window outside the image");
  const size_t total = (size_t)B * C * ch * cw;
  if (total == 0) return S4_OK;
  const int grid = (int)min((total + 255) / 256, (size_t)s4_num_sms() * 32);
  accumulate_crop_kernel<<<grid, 256, 0, stream>>>(crop, preds, count, B, C, H, W, y1, x1, ch, cw);
  return s4_check_launch("accumulate_crop");
}

// hist[0][c] = #(pred == label == c), hist[1][c] = #(pred == c), hist[2][c] = #(label == c) over the
// pixels with label != ignore_index; labels / predictions outside [0, C) are not counted (torch.histc
// with min=0, max=C-1 drops them too).  Block-local histograms in shared memory, one global atomic per
// (block, class).
__global__ void __launch_bounds__(256)
intersect_union_kernel(const long long* __restrict__ pred, const long long* __restrict__ label, size_t n, int C,
                       long long ignore_index, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[3 * 256];
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const long long l = label[i];
    if (l == ignore_index) continue;
    const long long p = pred[i];
    if (p >= 0 && p < C) {
      atomicAdd(&sh[C + (int)p], 1u);
      if (p == l) atomicAdd(&sh[(int)p], 1u);
    }
    if (l >= 0 && l < C) atomicAdd(&sh[2 * C + (int)l], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x)
    if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

extern "C" int s4_intersect_union(const long long* pred, const long long* label, long long n, int num_classes,
                                  long long ignore_index, long long* hist3, cudaStream_t stream) {
  S4ProfScope prof_("intersect_union", 0.0, 1, stream);
  S4_REQUIRE(num_classes >= 1 && num_classes <= 256, "intersect_union: num_classes=%d not in [1,256]", num_classes);
  if (n == 0) return S4_OK;
  const int grid = (int)min(((size_t)n + 255) / 256, (size_t)s4_num_sms() * 8);
  intersect_union_kernel<<<grid, 256, 0, stream>>>(pred, label, (size_t)n, num_classes, ignore_index,
                                                   (unsigned long long*)hist3);
  return s4_check_launch("intersect_union");
}
