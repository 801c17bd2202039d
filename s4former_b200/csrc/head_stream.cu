// Streaming (HBM-bound) kernels of the SETR-PUP head, second generation.
// Reference: mmseg/models/decode_heads/setr_up_head.py:49-77 (ConvModule conv -> SyncBN -> ReLU,
// then Upsample(bilinear, align_corners=False)), mmseg/ops/wrappers.py:30-51.
//
// The first-generation kernels (head.cu) own one (pixel, 16-byte channel vector) per thread and
// re-evaluate BatchNorm + ReLU and the bilinear weights for every neighbour they touch: ~500
// thread-instructions per 80 bytes of traffic, which is ISSUE-bound at ~2 TB/s (ncu: 20-25 %
// occupancy at 90-130 registers, DRAM 12-18 % busy).  Here a thread WALKS a short run of pixels
// along x with a sliding register window, and the bilinear filter is applied separably (vertical
// blend once per loaded column, horizontal blend per output), so every loaded element is
// normalised once and the instruction count per byte drops ~2.5x:
//
//   bn_relu_upsample_fwd_walk   out = bilinear_S(relu(y * scale + shift))              (NHWC bf16)
//   bn_relu_upsample_bwd_walk   dact = relu'(.) * bilinear_S^T(dout); BN-backward sums
//   bn_bwd_apply_x4             dy = A * dact + Bc * y + Cc, four vectors in flight per thread
//   pack_conv_weight2           both repacks with coalesced writes
//
// ATen semantics (upsample_bilinear2d, align_corners=False, integer scale S): source coordinate
// (o + 0.5) / S - 0.5 clamped at 0, second tap clamped at n - 1.  With clamped neighbour LOADS the
// interior weights reproduce the border cases exactly (a clamped row/column is a copy of the edge
// one), so no per-pixel weight evaluation is needed in the forward; the transpose (backward)
// folds the clamped taps' weights into the edge pixel explicitly.
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    f[2 * q] = __uint_as_float(w[q] << 16);
    f[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// VN bf16 elements (8 = 16 bytes, 4 = 8 bytes) <-> fp32 registers.  The walkers run with VN = 4
// where C allows: half the per-thread state (they are latency-bound at 128 registers / 25 %
// occupancy with VN = 8), still whole 32-byte sectors per 4 lanes.
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int VN> struct RawVec;
template <> struct RawVec<8> { uint4 v; };
template <> struct RawVec<4> { uint2 v; };
template <int VN>
__device__ __forceinline__ RawVec<VN> ldgv(const __nv_bfloat16* p);
template <>
__device__ __forceinline__ RawVec<8> ldgv<8>(const __nv_bfloat16* p) { RawVec<8> r; r.v = __ldg(reinterpret_cast<const uint4*>(p)); return r; }
template <>
__device__ __forceinline__ RawVec<4> ldgv<4>(const __nv_bfloat16* p) { RawVec<4> r; r.v = __ldg(reinterpret_cast<const uint2*>(p)); return r; }
__device__ __forceinline__ void unpackv(const RawVec<8>& r, float (&f)[8]) { unpack8(r.v, f); }
__device__ __forceinline__ void unpackv(const RawVec<4>& r, float (&f)[4]) {
  f[0] = __uint_as_float(r.v.x << 16); f[1] = __uint_as_float(r.v.x & 0xffff0000u);
  f[2] = __uint_as_float(r.v.y << 16); f[3] = __uint_as_float(r.v.y & 0xffff0000u);
}
__device__ __forceinline__ void storev(__nv_bfloat16* p, const float (&f)[8]) { *reinterpret_cast<uint4*>(p) = pack8(f); }
__device__ __forceinline__ void storev(__nv_bfloat16* p, const float (&f)[4]) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack2(f[0], f[1]), pack2(f[2], f[3]));
}

// ---------------------------------------------------------------------------------------------
// forward walker.  Unit of work: (image, input row iy, run of SEG input columns).  A group of
// CV = C/8 threads owns one unit (thread = 16-byte channel vector); a 256-thread block runs
// 256/CV units at a time.  The walker steps over PAIRS of adjacent input columns (c, c+1): the S
// output columns whose centres lie between the two input centres (S*c + S/2 + j, j < S) depend on
// those two columns only, so the register window is two columns of S vertical blends, every
// loaded element is normalised once, and each step writes S x S contiguous output pixels.
// Per step: 3 loads (rows iy-1, iy, iy+1 of column c+1, clamped), BN + ReLU, S vertical blends,
// S*S horizontal blends.  Pair c = -1 (first run) and c = W-1 (last run) produce the S/2 border
// columns on either side (clamped neighbour = the edge column itself).
// ---------------------------------------------------------------------------------------------
template <int S, int SEG, int VN>
__global__ void __launch_bounds__(256)
bn_relu_upsample_fwd_walk(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                          const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, int B, int H,
                          int W, int C, int n_units) {
  const int cv = C / VN, upb = 256 / cv;
  const int v = threadIdx.x % cv, ul = threadIdx.x / cv;
  float sc[VN], sh[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) { sc[e] = __ldg(scale + v * VN + e); sh[e] = __ldg(shift + v * VN + e); }
  const int segs = (W + SEG - 1) / SEG;
  const int OW = W * S;
  for (int unit = blockIdx.x * upb + ul; unit < n_units; unit += gridDim.x * upb) {
    const int seg = unit % segs;
    const int r = unit / segs;
    const int iy = r % H, b = r / H;
    const int x0 = seg * SEG;
    const __nv_bfloat16* img = x + (size_t)b * H * W * C + (size_t)v * VN;
    const __nv_bfloat16* row0 = img + (size_t)max(iy - 1, 0) * W * C;
    const __nv_bfloat16* row1 = img + (size_t)iy * W * C;
    const __nv_bfloat16* row2 = img + (size_t)min(iy + 1, H - 1) * W * C;
    __nv_bfloat16* obase = out + (((size_t)b * H * S + (size_t)iy * S) * OW) * C + (size_t)v * VN;
    float V[2][S][VN];     // vertical blends of two adjacent columns (ping-pong, statically indexed)
    auto column = [&](int ix, float (&Vc)[S][VN]) {
      const int cx = min(max(ix, 0), W - 1);
      float a0[VN], a1[VN], a2[VN];
      unpackv(ldgv<VN>(row0 + (size_t)cx * C), a0);
      unpackv(ldgv<VN>(row1 + (size_t)cx * C), a1);
      unpackv(ldgv<VN>(row2 + (size_t)cx * C), a2);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        a0[e] = fmaxf(fmaf(a0[e], sc[e], sh[e]), 0.f);
        a1[e] = fmaxf(fmaf(a1[e], sc[e], sh[e]), 0.f);
        a2[e] = fmaxf(fmaf(a2[e], sc[e], sh[e]), 0.f);
      }
#pragma unroll
      for (int ry = 0; ry < S; ++ry) {
        // source row = iy + (ry + 0.5)/S - 0.5: rows (iy-1, iy) for the upper half, (iy, iy+1) below
        const float ly = (ry + 0.5f) / S - 0.5f + (ry < S / 2 ? 1.f : 0.f);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const float up = ry < S / 2 ? a0[e] : a1[e], dn = ry < S / 2 ? a1[e] : a2[e];
          Vc[ry][e] = fmaf(ly, dn - up, up);
        }
      }
    };
    // the first run also owns pair (-1, 0): start one column earlier
    const int c0 = seg == 0 ? -1 : x0;
    column(c0, V[0]);
#pragma unroll
    for (int k = 0; k <= SEG; ++k) {
      const int c = c0 + k;                          // pair (c, c + 1)
      if (c >= min(x0 + SEG, W)) break;
      if (c + 3 < W) {                               // pull column c + 3 into L1 (two steps ahead)
        prefetch_l1(row0 + (size_t)(c + 3) * C);
        prefetch_l1(row1 + (size_t)(c + 3) * C);
        prefetch_l1(row2 + (size_t)(c + 3) * C);
      }
      column(c + 1, V[(k + 1) & 1]);
      const float (&Lc)[S][VN] = V[k & 1];
      const float (&Rc)[S][VN] = V[(k + 1) & 1];
#pragma unroll
      for (int j = 0; j < S; ++j) {
        const int ox = S * c + S / 2 + j;
        if (ox < 0 || ox >= OW) continue;
        const float lx = (j + 0.5f) / S;             // centre of the output column between c and c+1
#pragma unroll
        for (int ry = 0; ry < S; ++ry) {
          float o[VN];
#pragma unroll
          for (int e = 0; e < VN; ++e) o[e] = fmaf(lx, Rc[ry][e] - Lc[ry][e], Lc[ry][e]);
          storev(obase + ((size_t)ry * OW + ox) * C, o);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward walker (transpose of the above + ReLU mask + BatchNorm-backward sums), same pair walk:
// the S output columns between input columns c and c+1 are reduced vertically ONCE (2S row taps)
// and each reduced column T_j is consumed immediately by the two pixels it belongs to:
//     g[c]   += wr_j * T_j          g[c+1] += wl_j * T_j
// (interior: wl_j = (j + 0.5)/S, wr_j = 1 - wl_j; at the image border ATen's clamping puts the
// whole weight on the edge pixel -- evaluated by axis_w, which restates the forward's index math).
// A unit owns input pixels [x0, x0 + SEG) and walks pairs c = x0-1 .. x0+SEG-1.
// ---------------------------------------------------------------------------------------------
// weight of output index o for input index i along an axis of n inputs (0 when o is out of range)
template <int S>
__device__ __forceinline__ float axis_w(int o, int i, int n) {
  if (o < 0 || o >= n * S || i < 0 || i >= n) return 0.f;
  float src = ((float)o + 0.5f) / (float)S - 0.5f;
  if (src < 0.f) src = 0.f;
  const int i0 = (int)src;
  const int i1 = min(i0 + 1, n - 1);
  const float l = src - (float)i0;
  return (i == i0 ? 1.f - l : 0.f) + (i == i1 ? l : 0.f);
}

// One unit of the backward walk.  INTERIOR: every tap of the unit is inside the image (no clamps,
// no bounds tests, compile-time weights) -- all but the border rows / first and last run.
template <int S, int SEG, int VN, bool INTERIOR>
__device__ __forceinline__ void bwd_unit(const __nv_bfloat16* __restrict__ dimg, const __nv_bfloat16* __restrict__ xrow,
                                         __nv_bfloat16* __restrict__ drow, int iy, int x0, int H, int W, int C,
                                         const float (&sc)[VN], const float (&sf)[VN], const float (&mu)[VN],
                                         float (&as)[VN], float (&ad)[VN]) {
  constexpr int NT = 2 * S;
  constexpr int PD = 2;                            // prefetch distance in steps
  const int OH = H * S, OW = W * S;
  const int x1 = min(x0 + SEG, W);
  const int oy0 = S * iy - S / 2;
  float wy[NT];
  const __nv_bfloat16* rows[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    if (INTERIOR) {
      wy[t] = 1.f - fabsf((float)t + 0.5f - (float)S) / (float)S;
      rows[t] = dimg + (size_t)(oy0 + t) * OW * C;
    } else {
      wy[t] = axis_w<S>(oy0 + t, iy, H);
      rows[t] = dimg + (size_t)min(max(oy0 + t, 0), OH - 1) * OW * C;    // clamped rows carry weight 0
    }
  }
  float g[2][VN];        // gradient of pixels c (being finished) and c+1 (being started)
#pragma unroll
  for (int e = 0; e < VN; ++e) g[0][e] = 0.f;
#pragma unroll
  for (int k = 0; k <= SEG; ++k) {
    const int c = x0 - 1 + k;                      // pair (c, c + 1)
    if (!INTERIOR && c >= x1) break;
    float (&gc)[VN] = g[k & 1];
    float (&gn)[VN] = g[(k + 1) & 1];
#pragma unroll
    for (int e = 0; e < VN; ++e) gn[e] = 0.f;
    const bool inner = INTERIOR || (c >= 0 && c + 1 < W);
    if (INTERIOR && k + PD <= SEG) {
      // the walk is latency-bound (a step's 2S*S loads are consumed before the next step's are
      // issued, and registers are exhausted): pull the lines of step k + PD into L1 now
#pragma unroll
      for (int j = 0; j < S; ++j)
#pragma unroll
        for (int t = 0; t < NT; ++t) prefetch_l1(rows[t] + (size_t)(S * (c + PD) + S / 2 + j) * C);
    }
#pragma unroll
    for (int j = 0; j < S; ++j) {
      const int ox = S * c + S / 2 + j;
      if (!INTERIOR && (ox < 0 || ox >= OW)) continue;
      RawVec<VN> raw[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) raw[t] = ldgv<VN>(rows[t] + (size_t)ox * C);
      float T[VN];
#pragma unroll
      for (int e = 0; e < VN; ++e) T[e] = 0.f;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        float d[VN];
        unpackv(raw[t], d);
#pragma unroll
        for (int e = 0; e < VN; ++e) T[e] = fmaf(wy[t], d[e], T[e]);
      }
      const float wl = inner ? (j + 0.5f) / S : axis_w<S>(ox, c + 1, W);
      const float wr = inner ? 1.f - (j + 0.5f) / S : axis_w<S>(ox, c, W);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        gc[e] = fmaf(wr, T[e], gc[e]);
        gn[e] = fmaf(wl, T[e], gn[e]);
      }
    }
    if (k > 0) {     // pixel c is complete: ReLU mask, BatchNorm-backward sums, store
      float xe[VN], o[VN];
      unpackv(ldgv<VN>(xrow + (size_t)c * C), xe);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const float da = fmaf(xe[e], sc[e], sf[e]) > 0.f ? gc[e] : 0.f;
        o[e] = da;
        as[e] += da;
        ad[e] = fmaf(da, xe[e] - mu[e], ad[e]);    // invstd is applied at the fold
      }
      storev(drow + (size_t)c * C, o);
    }
  }
}

template <int S, int SEG, int VN>
__global__ void __launch_bounds__(256, VN == 8 ? 2 : 3)
bn_relu_upsample_bwd_walk(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ x,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd,
                          __nv_bfloat16* __restrict__ dact, float* __restrict__ dsum,
                          float* __restrict__ ddot, int B, int H, int W, int C, int n_units) {
  extern __shared__ float sh[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int cv = C / VN, upb = 256 / cv;
  const int v = threadIdx.x % cv, ul = threadIdx.x / cv;
  float sc[VN], sf[VN], mu[VN], as[VN], ad[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    sc[e] = __ldg(scale + v * VN + e); sf[e] = __ldg(shift + v * VN + e); mu[e] = __ldg(mean + v * VN + e);
    as[e] = 0.f; ad[e] = 0.f;
  }
  const int segs = (W + SEG - 1) / SEG;
  const int OH = H * S, OW = W * S;
  for (int unit = blockIdx.x * upb + ul; unit < n_units; unit += gridDim.x * upb) {
    const int seg = unit % segs;
    const int r = unit / segs;
    const int iy = r % H, b = r / H;
    const int x0 = seg * SEG;
    const __nv_bfloat16* dimg = dout + (size_t)b * OH * OW * C + (size_t)v * VN;
    const __nv_bfloat16* xrow = x + (((size_t)b * H + iy) * W) * C + (size_t)v * VN;
    __nv_bfloat16* drow = dact + (((size_t)b * H + iy) * W) * C + (size_t)v * VN;
    const bool interior = iy > 0 && iy < H - 1 && x0 > 0 && x0 + SEG < W;
    if (interior) bwd_unit<S, SEG, VN, true>(dimg, xrow, drow, iy, x0, H, W, C, sc, sf, mu, as, ad);
    else bwd_unit<S, SEG, VN, false>(dimg, xrow, drow, iy, x0, H, W, C, sc, sf, mu, as, ad);
  }
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    const int c = v * VN + e;
    atomicAdd(&sh[c], as[e]);
    atomicAdd(&sh[C + c], ad[e] * __ldg(invstd + c));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dsum + i, sh[i]);
    atomicAdd(ddot + i, sh[C + i]);
  }
}

// ---------------------------------------------------------------------------------------------
// backward STRIP kernel: the same arithmetic as the walker, but the dout rows are staged in shared
// memory by bulk async copies (cp.async.bulk + mbarrier) instead of per-thread loads.  The walker
// is latency-bound: a thread's 2S*S loads of a step must land before the next step's are issued,
// its registers are exhausted (128), and every dout row is fetched by the two input rows that
// share it (ncu: 25 % occupancy, 31 % issue slots, 2.6 TB/s).  Here a block owns a column strip of
// SEGW input pixels of one image and walks DOWN a chunk of input rows; the contiguous
// (S*SEGW + S)-pixel pieces of the dout rows stream through a ring of R = 3S slots (2S live rows +
// S rows in flight), so every dout row is read once per strip, S*piece bytes per block are always
// in flight without costing a register, and the compute reads conflict-free 16-byte vectors.
// Thread = (16-byte channel vector, pixel lane); a lane owns RL adjacent input pixels of the strip
// and applies the separable filter transpose as the walker does (vertical reduce per dout column,
// then two horizontal taps).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}

template <int S, int RL>
__global__ void __launch_bounds__(256)
bn_relu_upsample_bwd_strip(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ x,
                           const float* __restrict__ scale, const float* __restrict__ shift,
                           const float* __restrict__ mean, const float* __restrict__ invstd,
                           __nv_bfloat16* __restrict__ dact, float* __restrict__ dsum,
                           float* __restrict__ ddot, int B, int H, int W, int C, int strips, int rchunks,
                           int rows_per, uint32_t slot_bytes) {
  constexpr int R = 3 * S, NT = 2 * S;
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t raw = tc::smem_u32(smem_dyn);
  const uint32_t ring = (raw + 127u) & ~127u;
  const uint32_t bars = ring + R * slot_bytes;
  float* sh = reinterpret_cast<float*>(smem_dyn + (bars + 128u - raw));      // [2][C]
  const int cv = C >> 3, npl = 256 / cv, segw = RL * npl;
  const int v = threadIdx.x % cv, pl = threadIdx.x / cv;
  const int item = blockIdx.x;
  const int rc = item % rchunks, st = (item / rchunks) % strips, b = item / (rchunks * strips);
  const int ya = rc * rows_per, yb = min(ya + rows_per, H);
  const int n_rows = yb - ya;
  const int OH = H * S, OW = W * S;
  const int x0 = st * segw, x1 = min(x0 + segw, W);
  const int oxa_u = S * x0 - S / 2;                   // first dout column of the strip (may be -S/2)
  const int oxa = max(oxa_u, 0), oxb = min(S * x1 + S / 2, OW);
  const uint32_t piece_bytes = (uint32_t)(oxb - oxa) * (uint32_t)C * 2u;
  const uint32_t dst_off = (uint32_t)(oxa - oxa_u) * (uint32_t)C * 2u;
  const int oy_first = S * ya - S / 2;                // dout row of ring index q = 0
  const int total_q = S * n_rows + S;
  const __nv_bfloat16* dimg = dout + (size_t)b * OH * OW * C;
  auto issue = [&](int q) {
    const int oy = min(max(oy_first + q, 0), OH - 1);   // rows outside the image carry weight 0
    const int slot = q % R;
    tc::mbar_expect_tx(bars + 8u * slot, piece_bytes);
    bulk_load_1d(ring + slot * slot_bytes + dst_off, dimg + ((size_t)oy * OW + oxa) * C, piece_bytes,
                 bars + 8u * slot);
  };
  if (threadIdx.x == 0) {
    for (int i = 0; i < R; ++i) tc::mbar_init(bars + 8u * i, 1);
    tc::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  int next_q = 0;
  if (threadIdx.x == 0 && n_rows > 0)
    for (; next_q < min(total_q, R); ++next_q) issue(next_q);

  float sc[8], sf[8], mu[8], as[8], ad[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = __ldg(scale + v * 8 + e); sf[e] = __ldg(shift + v * 8 + e); mu[e] = __ldg(mean + v * 8 + e);
    as[e] = 0.f; ad[e] = 0.f;
  }
  const int c0 = x0 + RL * pl;                         // this lane's first input pixel
  const bool lane_on = c0 < x1;
  const uint32_t vec_off = (uint32_t)v * 16u;
  for (int i = 0; i < n_rows; ++i) {
    const int iy = ya + i;
    const size_t rowbase = (((size_t)b * H + iy) * W) * C + (size_t)v * 8;
    uint4 xr[RL];
#pragma unroll
    for (int r = 0; r < RL; ++r)
      if (c0 + r < x1) xr[r] = ldg16(x + rowbase + (size_t)(c0 + r) * C);
    float wy[NT];
    uint32_t rowaddr[NT];
    const bool yin = iy > 0 && iy < H - 1;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      wy[t] = yin ? 1.f - fabsf((float)t + 0.5f - (float)S) / (float)S : axis_w<S>(S * iy - S / 2 + t, iy, H);
      const int q = S * i + t;
      const int slot = q % R;
      rowaddr[t] = ring + slot * slot_bytes + vec_off;
      tc::mbar_wait(bars + 8u * slot, (uint32_t)(q / R) & 1u);
    }
    if (lane_on) {
      float g[RL][8];
#pragma unroll
      for (int r = 0; r < RL; ++r)
#pragma unroll
        for (int e = 0; e < 8; ++e) g[r][e] = 0.f;
#pragma unroll
      for (int pr = 0; pr <= RL; ++pr) {
        const int c = c0 - 1 + pr;                     // pair (c, c + 1)
        const bool inner = c >= 0 && c + 1 < W;
#pragma unroll
        for (int j = 0; j < S; ++j) {
          const int ox = S * c + S / 2 + j;
          if (ox < 0 || ox >= OW) continue;
          const uint32_t coff = (uint32_t)(ox - oxa_u) * (uint32_t)C * 2u;
          uint4 rawv[NT];
#pragma unroll
          for (int t = 0; t < NT; ++t) rawv[t] = lds16(rowaddr[t] + coff);
          float T[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) T[e] = 0.f;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            float d[8];
            unpack8(rawv[t], d);
#pragma unroll
            for (int e = 0; e < 8; ++e) T[e] = fmaf(wy[t], d[e], T[e]);
          }
          const float wl = inner ? (j + 0.5f) / S : axis_w<S>(ox, c + 1, W);
          const float wr = inner ? 1.f - (j + 0.5f) / S : axis_w<S>(ox, c, W);
          if (pr > 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) g[pr > 0 ? pr - 1 : 0][e] = fmaf(wr, T[e], g[pr > 0 ? pr - 1 : 0][e]);
          }
          if (pr < RL) {
#pragma unroll
            for (int e = 0; e < 8; ++e) g[pr < RL ? pr : 0][e] = fmaf(wl, T[e], g[pr < RL ? pr : 0][e]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < RL; ++r) {
        if (c0 + r < x1) {
          float xe[8], o[8];
          unpack8(xr[r], xe);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float da = fmaf(xe[e], sc[e], sf[e]) > 0.f ? g[r][e] : 0.f;
            o[e] = da;
            as[e] += da;
            ad[e] = fmaf(da, xe[e] - mu[e], ad[e]);
          }
          *reinterpret_cast<uint4*>(dact + rowbase + (size_t)(c0 + r) * C) = pack8(o);
        }
      }
    }
    __syncthreads();          // rows q < S (i + 1) are dead: their slots may be refilled
    if (threadIdx.x == 0)
      for (; next_q < total_q && next_q - R < S * (i + 1); ++next_q) issue(next_q);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = v * 8 + e;
    atomicAdd(&sh[c], as[e]);
    atomicAdd(&sh[C + c], ad[e] * __ldg(invstd + c));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dsum + i, sh[i]);
    atomicAdd(ddot + i, sh[C + i]);
  }
}

// ---------------------------------------------------------------------------------------------
// dy = gamma*invstd*(dact - dsum/n - xhat*ddot/n) = A*dact + Bc*y + Cc; four 16-byte vectors of
// each operand in flight per thread (the first-generation kernel had one: latency-bound).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_bwd_apply_x4(const __nv_bfloat16* __restrict__ dact, const __nv_bfloat16* __restrict__ x,
                const float* __restrict__ gamma, const float* __restrict__ mean,
                const float* __restrict__ invstd, const float* __restrict__ dsum,
                const float* __restrict__ ddot, float inv_n, __nv_bfloat16* __restrict__ dy, size_t rows,
                int C) {
  const int cv = C >> 3, rpb = 256 / cv;               // rows per block pass
  const int v = threadIdx.x % cv, rl = threadIdx.x / cv;
  float A[8], Bc[8], Cc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = v * 8 + e;
    const float is = __ldg(invstd + c), a = __ldg(gamma + c) * is;
    const float k = is * __ldg(ddot + c) * inv_n;          // xhat*ddot/n = (x - mean) * k
    A[e] = a;
    Bc[e] = -a * k;
    Cc[e] = a * (__ldg(mean + c) * k - __ldg(dsum + c) * inv_n);
  }
  const size_t stride = (size_t)gridDim.x * rpb;
  for (size_t r0 = (size_t)blockIdx.x * rpb + rl; r0 < rows; r0 += 4 * stride) {
    uint4 d[4], xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t r = r0 + u * stride;
      if (r < rows) {
        d[u] = ldg16(dact + r * C + (size_t)v * 8);
        xv[u] = ldg16(x + r * C + (size_t)v * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t r = r0 + u * stride;
      if (r < rows) {
        float df[8], xf[8], o[8];
        unpack8(d[u], df);
        unpack8(xv[u], xf);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(A[e], df[e], fmaf(Bc[e], xf[e], Cc[e]));
        *reinterpret_cast<uint4*>(dy + r * C + (size_t)v * 8) = pack8(o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// w [Cout][Cin][3][3] f32 -> fwd [Cout][tap*Cin + ci], dgrad [Cin][(8-tap)*Cout + co] (bf16).
// Two half-grids, each writing ITS output coalesced (the reads are 36-byte runs served by L2).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_conv_weight2(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd,
                  int Cin, int Cout) {
  const int total = Cout * Cin;
  const int half = gridDim.x / 2;
  if ((int)blockIdx.x < half) {            // forward layout: consecutive threads = consecutive ci
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += half * blockDim.x) {
      const int ci = i % Cin, co = i / Cin;
      const float* src = w + (size_t)i * 9;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
        wf[(size_t)co * 9 * Cin + tap * Cin + ci] = __float2bfloat16_rn(__ldg(src + tap));
    }
  } else {                                 // dgrad layout: consecutive threads = consecutive co
    for (int i = (blockIdx.x - half) * blockDim.x + threadIdx.x; i < total; i += (gridDim.x - half) * blockDim.x) {
      const int co = i % Cout, ci = i / Cout;
      const float* src = w + ((size_t)co * Cin + ci) * 9;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
        wd[(size_t)ci * 9 * Cout + (8 - tap) * Cout + co] = __float2bfloat16_rn(__ldg(src + tap));
    }
  }
}

}  // namespace

// ---- launchers (called from head.cu; return false when the shape is outside the fast path) ----
template <int S, int VN>
static void launch_fwd_walk(const void* x, const float* scale, const float* shift, void* out, int B, int H, int W,
                            int C, cudaStream_t st) {
  constexpr int SEG = 8;
  const int upb = 256 / (C / VN);
  const long long units = (long long)B * H * ((W + SEG - 1) / SEG);
  const int grid = (int)std::min<long long>((units + upb - 1) / upb, (long long)s4_num_sms() * 16);
  bn_relu_upsample_fwd_walk<S, SEG, VN><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, scale, shift,
                                                              (__nv_bfloat16*)out, B, H, W, C, (int)units);
}

// lanes per pixel must divide the block: VN = 4 (8-byte vectors) when C <= 1024, else 8
static int walk_vn(int C) {
  if (C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0) return 4;
  if (C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0) return 8;
  return 0;
}

bool s4_stream_upsample_fwd(const void* x, const float* scale, const float* shift, void* out, int B, int H,
                            int W, int C, int s, cudaStream_t st) {
  const int vn = walk_vn(C);
  if ((s != 2 && s != 4) || !vn) return false;
  if ((long long)B * H * W * s * s * (C / vn) >= (1ll << 31)) return false;
  if ((long long)B * H * ((W + 7) / 8) >= (1ll << 30)) return false;
  if (s == 2 && vn == 4) launch_fwd_walk<2, 4>(x, scale, shift, out, B, H, W, C, st);
  else if (s == 2) launch_fwd_walk<2, 8>(x, scale, shift, out, B, H, W, C, st);
  else if (vn == 4) launch_fwd_walk<4, 4>(x, scale, shift, out, B, H, W, C, st);
  else launch_fwd_walk<4, 8>(x, scale, shift, out, B, H, W, C, st);
  return true;
}

template <int S, int VN>
static void launch_bwd_walk(const void* dout, const void* x, const float* scale, const float* shift,
                            const float* mean, const float* invstd, void* dact, float* dsum, float* ddot,
                            int B, int H, int W, int C, cudaStream_t st) {
  constexpr int SEG = 8;
  const int upb = 256 / (C / VN);
  const long long units = (long long)B * H * ((W + SEG - 1) / SEG);
  // whole rounds of units per block (a block folds its BatchNorm sums once, at the end: blocks
  // with one round more than their neighbours made everybody wait at that barrier)
  const long long slots = (long long)s4_num_sms() * (VN == 8 ? 2 : 3) * upb;
  const long long rounds = (units + slots - 1) / slots;
  const int grid = (int)((units + rounds * upb - 1) / (rounds * upb));
  const size_t smem = 2 * (size_t)C * sizeof(float);
  bn_relu_upsample_bwd_walk<S, SEG, VN><<<grid, 256, smem, st>>>(
      (const __nv_bfloat16*)dout, (const __nv_bfloat16*)x, scale, shift, mean, invstd, (__nv_bfloat16*)dact,
      dsum, ddot, B, H, W, C, (int)units);
}

// strip kernel launch: returns false when the shape does not fit (the walker takes over)
template <int S, int RL>
static bool launch_bwd_strip(const void* dout, const void* x, const float* scale, const float* shift,
                             const float* mean, const float* invstd, void* dact, float* dsum, float* ddot,
                             int B, int H, int W, int C, cudaStream_t st) {
  const int cv = C / 8, npl = 256 / cv, segw = RL * npl;
  const size_t slot = (size_t)(S * segw + S) * C * 2;
  const size_t smem = 3 * S * slot + 128 + 128 + 2 * (size_t)C * sizeof(float);
  if (smem > 227 * 1024) return false;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(bn_relu_upsample_bwd_strip<S, RL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    attr_done = true;
  }
  const int strips = (W + segw - 1) / segw;
  int bps = (int)((227 * 1024) / (smem + 1024));
  if (bps > 3) bps = 3;
  if (bps < 1) bps = 1;
  const long long slots = (long long)s4_num_sms() * bps;
  int rchunks = (int)std::max<long long>(1, slots / ((long long)B * strips));
  if (rchunks > H) rchunks = H;
  int rows_per = (H + rchunks - 1) / rchunks;
  if (rows_per < 4) rows_per = std::min(4, H);       // halo rows: S per chunk
  rchunks = (H + rows_per - 1) / rows_per;
  const long long grid = (long long)B * strips * rchunks;
  if (grid >= (1ll << 31)) return false;
  bn_relu_upsample_bwd_strip<S, RL><<<(unsigned)grid, 256, smem, st>>>(
      (const __nv_bfloat16*)dout, (const __nv_bfloat16*)x, scale, shift, mean, invstd, (__nv_bfloat16*)dact,
      dsum, ddot, B, H, W, C, strips, rchunks, rows_per, (uint32_t)slot);
  return true;
}

bool s4_stream_upsample_bwd(const void* dout, const void* x, const float* scale, const float* shift,
                            const float* mean, const float* invstd, void* dact, float* dsum, float* ddot,
                            int B, int H, int W, int C, int s, cudaStream_t st) {
  // measured: the backward is faster with 16-byte vectors (156 vs 193 us at 8 x 128^2 x 256)
  const int vn = (C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0) ? 8 : walk_vn(C);
  if ((s != 2 && s != 4) || !vn) return false;
  if ((long long)B * H * W * s * s * (C / vn) >= (1ll << 31)) return false;
  if ((long long)B * H * ((W + 7) / 8) >= (1ll << 30)) return false;
  // bulk-copy strip kernel (C a multiple of 8 with C/8 dividing 256, 16-byte aligned rows)
  static const bool no_strip = getenv("S4_NO_BWD_STRIP") != nullptr;
  // (cp.async.bulk needs 16-byte aligned global addresses: an oddly offset view takes the walker)
  if (!no_strip && C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0 && H >= 2 && W >= 2 &&
      (((uintptr_t)dout) & 15) == 0) {
    const int npl = 256 / (C / 8);
    bool ok = false;
    if (s == 2 && W >= 4 * npl) ok = launch_bwd_strip<2, 2>(dout, x, scale, shift, mean, invstd, dact, dsum, ddot, B, H, W, C, st);
    else if (s == 2) ok = launch_bwd_strip<2, 1>(dout, x, scale, shift, mean, invstd, dact, dsum, ddot, B, H, W, C, st);
    else ok = launch_bwd_strip<4, 1>(dout, x, scale, shift, mean, invstd, dact, dsum, ddot, B, H, W, C, st);
    if (ok) return true;
  }
  if (s == 2 && vn == 4) launch_bwd_walk<2, 4>(dout, x, scale, shift, mean, invstd, dact, dsum, ddot, B, H, W, C, st);
  else if (s == 2) launch_bwd_walk<2, 8>(dout, x, scale, shift, mean, invstd, dact, dsum, ddot, B, H, W, C, st);
  else if (vn == 4) launch_bwd_walk<4, 4>(dout, x, scale, shift, mean, invstd, dact, dsum, ddot, B, H, W, C, st);
  else launch_bwd_walk<4, 8>(dout, x, scale, shift, mean, invstd, dact, dsum, ddot, B, H, W, C, st);
  return true;
}

bool s4_stream_bn_bwd_apply(const void* dact, const void* x, const float* gamma, const float* mean,
                            const float* invstd, const float* dsum, const float* ddot, double count,
                            void* dy, long long rows, int C, cudaStream_t st) {
  if (C % 8 || 256 % (C / 8) || C / 8 > 256) return false;
  const int rpb = 256 / (C / 8);
  const long long blocks = (rows + 4ll * rpb - 1) / (4ll * rpb);
  const int grid = (int)std::min<long long>(std::max<long long>(blocks, 1), (long long)s4_num_sms() * 8);
  bn_bwd_apply_x4<<<grid, 256, 0, st>>>((const __nv_bfloat16*)dact, (const __nv_bfloat16*)x, gamma, mean,
                                        invstd, dsum, ddot, (float)(1.0 / count), (__nv_bfloat16*)dy,
                                        (size_t)rows, C);
  return true;
}

bool s4_stream_pack_conv_weight(const float* w, void* wf, void* wd, int Cin, int Cout, cudaStream_t st) {
  if (!wf || !wd) return false;
  const int total = Cin * Cout;
  int half = std::min((total + 255) / 256, s4_num_sms() * 4);
  pack_conv_weight2<<<2 * half, 256, 0, st>>>(w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd, Cin, Cout);
  return true;
}
