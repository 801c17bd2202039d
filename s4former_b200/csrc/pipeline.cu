// GPU-side input pipeline of the weak / strong branches.
// Reference: configs/setr/*_MT_w_ours.py:42-126 (train / strong / weak pipelines),
// mmseg/datasets/pipelines/transforms.py:1165-1272 PhotoMetricDistortion, :572-604 Normalize
// (mmcv.imnormalize), :484-565 Pad, formatting.py:202-225 DefaultFormatBundle.
//
// One launch turns a batch of uint8 BGR crops (HWC, as RandomCrop / RandomFlip leave them) and uint8
// label crops into the float32 CHW network input and the int64 label map, for any number of
// BRANCHES per crop (a labeled crop feeds one branch; an unlabeled crop feeds the student and the
// teacher branch, each with its own distortion parameters) -- the host uploads 1 byte per pixel and
// channel instead of 4 bytes per branch, and the per-pixel arithmetic leaves the dataloader workers.
//
// Arithmetic follows the reference step by step, INCLUDING the uint8 re-quantisation between the
// distortion steps:
//   convert(img, alpha, beta) = uint8(clip(float32(img) * float32(alpha) + float32(beta), 0, 255))
//                               (truncation, transforms.py:1197-1201)
//   bgr2hsv / hsv2bgr         = OpenCV's 8-bit conversions (H in [0,180)): BGR->HSV is the integer
//                               table algorithm (bit-exact); HSV->BGR is the float sector formula with
//                               truncation, which is what OpenCV's SIMD row path computes -- OpenCV's own
//                               scalar tail rounds instead, so the reference is position dependent here
//                               (tests/test_pipeline_gpu.py: <= 1 LSB on < 0.05 % of the pixels).
//   normalize                 = float(double(float(x) - float(mean)) * (1 / double(std))), BGR -> RGB first
//                               (cv2.multiply keeps its scalar in double: pinned against OpenCV, bit-exact)
//   pad                       = 0.0 for the image (after normalisation), seg_pad_val for the labels
#include <algorithm>

#include "common.cuh"

struct S4PmdParams {     // per (crop, branch): PhotoMetricDistortion draws (host RNG, reference order)
  int do_brightness; float beta;
  int mode;                           // 1: contrast before saturation / hue, 0: after
  int do_contrast; float alpha_c;
  int do_saturation; float alpha_s;
  int do_hue; int hue_delta;
};

__device__ __forceinline__ int convert_u8(int x, float alpha, float beta) {
  float v = __fadd_rn(__fmul_rn((float)x, alpha), beta);     // two roundings, no FMA (numpy float32)
  v = fminf(fmaxf(v, 0.f), 255.f);
  return (int)v;                                               // astype(uint8): truncation
}

__device__ __forceinline__ int cv_round_pos(double x) { return (int)rint(x); }

__device__ __forceinline__ void bgr2hsv_u8(int b, int g, int r, int& h, int& s, int& v) {
  const int hsv_shift = 12;
  v = max(max(b, g), r);
  const int vmin = min(min(b, g), r);
  const int diff = v - vmin;
  const int sdiv = v ? cv_round_pos((double)(255 << hsv_shift) / (double)v) : 0;
  const int hdiv = diff ? cv_round_pos((double)(180 << hsv_shift) / (6.0 * (double)diff)) : 0;
  const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
  s = (diff * sdiv + (1 << (hsv_shift - 1))) >> hsv_shift;
  int hh = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
  hh = (hh * hdiv + (1 << (hsv_shift - 1))) >> hsv_shift;
  h = hh + (hh < 0 ? 180 : 0);
}

__device__ __forceinline__ void hsv2bgr_u8(int h, int s, int v, int& b, int& g, int& r) {
  const float sf = __fmul_rn((float)s, 1.0f / 255.0f), vf = __fmul_rn((float)v, 1.0f / 255.0f);
  float bb, gg, rr;
  if (s == 0) {
    bb = gg = rr = vf;
  } else {
    float hh = __fmul_rn((float)h, 6.0f / 180.0f);
    int sector = (int)floorf(hh);
    float f = __fsub_rn(hh, (float)sector);
    if ((unsigned)sector >= 6u) { sector = 0; f = 0.f; }
    float tab[4];
    tab[0] = vf;
    tab[1] = __fmul_rn(vf, __fsub_rn(1.f, sf));
    tab[2] = __fmul_rn(vf, __fsub_rn(1.f, __fmul_rn(sf, f)));
    tab[3] = __fmul_rn(vf, __fsub_rn(1.f, __fmul_rn(sf, __fsub_rn(1.f, f))));
    const int sd[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};
    bb = tab[sd[sector][0]]; gg = tab[sd[sector][1]]; rr = tab[sd[sector][2]];
  }
  b = (int)fminf(fmaxf(floorf(__fmul_rn(bb, 255.f)), 0.f), 255.f);
  g = (int)fminf(fmaxf(floorf(__fmul_rn(gg, 255.f)), 0.f), 255.f);
  r = (int)fminf(fmaxf(floorf(__fmul_rn(rr, 255.f)), 0.f), 255.f);
}

__device__ __forceinline__ int py_mod180(int x) {
  int m = x % 180;
  return m < 0 ? m + 180 : m;
}

// crops: per crop a device pointer to [h, w, 3] uint8 (BGR), labels [h, w] uint8 (may be null);
// branch i reads crop crop_of[i], writes image i of out_img [NB, 3, PH, PW] / out_lab [NB, 1, PH, PW].
__global__ void __launch_bounds__(256)
branch_pipeline_kernel(const unsigned char* const* __restrict__ crops, const unsigned char* const* __restrict__ labels,
                       const int* __restrict__ crop_hw, const int* __restrict__ crop_of,
                       const S4PmdParams* __restrict__ params, const float* __restrict__ mean_rgb,
                       const double* __restrict__ stdinv_rgb, int to_rgb, int seg_pad, float* __restrict__ out_img,
                       long long* __restrict__ out_lab, unsigned char* __restrict__ out_u8, int PH, int PW) {
  const int br = blockIdx.y;
  const int ci = crop_of[br];
  const int h = crop_hw[2 * ci], w = crop_hw[2 * ci + 1];
  const S4PmdParams p = params[br];
  const size_t plane = (size_t)PH * PW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / PW), x = (int)(i % PW);
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    long long lab = seg_pad;
    if (y < h && x < w) {
      const unsigned char* px = crops[ci] + ((size_t)y * w + x) * 3;
      int b = px[0], g = px[1], r = px[2];
      if (p.do_brightness) { b = convert_u8(b, 1.f, p.beta); g = convert_u8(g, 1.f, p.beta); r = convert_u8(r, 1.f, p.beta); }
      if (p.mode == 1 && p.do_contrast) {
        b = convert_u8(b, p.alpha_c, 0.f); g = convert_u8(g, p.alpha_c, 0.f); r = convert_u8(r, p.alpha_c, 0.f);
      }
      if (p.do_saturation) {
        int hh, ss, vv;
        bgr2hsv_u8(b, g, r, hh, ss, vv);
        ss = convert_u8(ss, p.alpha_s, 0.f);
        hsv2bgr_u8(hh, ss, vv, b, g, r);
      }
      if (p.do_hue) {
        int hh, ss, vv;
        bgr2hsv_u8(b, g, r, hh, ss, vv);
        hh = py_mod180(hh + p.hue_delta);
        hsv2bgr_u8(hh, ss, vv, b, g, r);
      }
      if (p.mode == 0 && p.do_contrast) {
        b = convert_u8(b, p.alpha_c, 0.f); g = convert_u8(g, p.alpha_c, 0.f); r = convert_u8(r, p.alpha_c, 0.f);
      }
      if (out_u8) {          // the distorted uint8 image (parity tests)
        unsigned char* q = out_u8 + ((size_t)br * plane + i) * 3;
        q[0] = (unsigned char)b; q[1] = (unsigned char)g; q[2] = (unsigned char)r;
      }
      const int c0 = to_rgb ? r : b, c2 = to_rgb ? b : r;
      o0 = (float)((double)__fsub_rn((float)c0, mean_rgb[0]) * stdinv_rgb[0]);
      o1 = (float)((double)__fsub_rn((float)g, mean_rgb[1]) * stdinv_rgb[1]);
      o2 = (float)((double)__fsub_rn((float)c2, mean_rgb[2]) * stdinv_rgb[2]);
      if (labels && labels[ci]) lab = labels[ci][(size_t)y * w + x];
    } else if (out_u8) {
      unsigned char* q = out_u8 + ((size_t)br * plane + i) * 3;
      q[0] = q[1] = q[2] = 0;
    }
    float* oi = out_img + (size_t)br * 3 * plane + i;
    oi[0] = o0; oi[plane] = o1; oi[2 * plane] = o2;
    if (out_lab) out_lab[(size_t)br * plane + i] = lab;
  }
}

extern "C" int s4_branch_pipeline(const void* const* crops, const void* const* labels, const int* crop_hw,
                                  const int* crop_of, const void* pmd_params, const float* mean, const double* stdinv,
                                  int to_rgb, int seg_pad_val, float* out_img, long long* out_label, void* out_u8,
                                  int n_branches, int pad_h, int pad_w, cudaStream_t stream) {
  S4ProfScope prof_("branch_pipeline", 0.0, 1, stream);
  if (n_branches == 0 || pad_h * pad_w == 0) return S4_OK;
  S4_REQUIRE(crops && crop_hw && crop_of && pmd_params && mean && stdinv && out_img, "branch_pipeline: null argument");
  const size_t plane = (size_t)pad_h * pad_w;
  dim3 grid((unsigned)std::min<size_t>((plane + 255) / 256, 1024), (unsigned)n_branches);
  branch_pipeline_kernel<<<grid, 256, 0, stream>>>(
      (const unsigned char* const*)crops, (const unsigned char* const*)labels, crop_hw, crop_of,
      (const S4PmdParams*)pmd_params, mean, stdinv, to_rgb, seg_pad_val, out_img, out_label,
      (unsigned char*)out_u8, pad_h, pad_w);
  return s4_check_launch("branch_pipeline");
}

extern "C" int s4_pmd_params_size() { return (int)sizeof(S4PmdParams); }
