// Last SETR-PUP stage, bf16 path: BatchNorm + ReLU + 1x1 `conv_seg` (decode_head.py:107,311-316
// commuted in front of the final bilinear upsample) and its backward.
//
// These contractions are tiny in FLOPs (C x NC per pixel, NC = 19/21 classes) and bound by the
// HBM traffic of the [pixels, C] conv output, so they are fused around that traffic instead of
// being shaped into big tensor-core GEMMs: every kernel streams the conv output ONCE, applies
// BN(+ReLU) in registers and runs the small contraction with warp-level bf16 MMAs
// (mma.sync m16n8k16, fp32 accumulate).
//
//   cls_fwd          z[p, j]   = bias[j] + sum_c w[j,c] relu(y[p,c] scale[c] + shift[c])
//   cls_bwd_reduce   dw[j,c]  += sum_p dz[p,j] act[p,c];  dbias[j] += sum_p dz[p,j]
//                    dsum[c]  += sum_p da[p,c];           ddot[c]  += sum_p da[p,c] xhat[p,c]
//                    with da = relu'(.) (dz w)  -- the two sums BatchNorm's backward needs
//   cls_bwd_apply    dy[p,c]   = gamma invstd (da - dsum/n - xhat ddot/n)   (da recomputed)
// plus the logits upsample pair (NHWC z <-> NCHW fp32 logits) with smem-staged rows.
#include "common.cuh"

namespace {

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool ok) {
  const int sz = ok ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------
// forward: one warp per 16 pixel rows, all C channels.  The contraction index is permuted so
// that each lane's A fragments are 8 CONSECUTIVE channels of its two rows (16-byte loads, four
// lanes = 64 contiguous bytes per row): channel(step s, mma m, slot) = 32 s + 8 t + 4 m + slot.
// ---------------------------------------------------------------------------------------------
template <int STEPS, int NT>   // C = 32 * STEPS, NC <= 8 * NT
__global__ void __launch_bounds__(256, STEPS <= 8 ? 2 : 1)
cls_fwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
               const float* __restrict__ shift, const float* __restrict__ w,
               const float* __restrict__ bias, float* __restrict__ z, long long rows, int NC) {
  constexpr int C = 32 * STEPS;
  extern __shared__ __align__(16) uint8_t smem[];
  float2* ss = reinterpret_cast<float2*>(smem);                       // [C] (scale, shift)
  uint4* bfr = reinterpret_cast<uint4*>(smem + C * sizeof(float2));   // [STEPS][NT][32]
  for (int i = threadIdx.x; i < C; i += blockDim.x) ss[i] = make_float2(scale[i], shift[i]);
  for (int i = threadIdx.x; i < STEPS * NT * 32; i += blockDim.x) {
    const int ln = i & 31, nt = (i >> 5) % NT, s = (i >> 5) / NT;
    const int n = nt * 8 + (ln >> 2), ch = 32 * s + 8 * (ln & 3);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = n < NC ? w[(size_t)n * C + ch + e] : 0.f;
    bfr[i] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long nblk = (rows + 15) / 16;
  // register double buffering: the loads of the warp's NEXT 16-row block are issued before the
  // current block's BN + ReLU + MMAs (v1 was latency-bound: 11 long-scoreboard stall cycles per
  // issue at 21 % issue utilisation)
  uint4 na[STEPS], nb[STEPS];
  auto fetch = [&](long long blk) {
    const long long ra = blk * 16 + g, rb = ra + 8;
    const bool oka = blk < nblk && ra < rows, okb = blk < nblk && rb < rows;
#pragma unroll
    for (int s = 0; s < STEPS; ++s) {
      na[s] = oka ? __ldg(reinterpret_cast<const uint4*>(y + ra * C + 32 * s + 8 * t)) : make_uint4(0, 0, 0, 0);
      nb[s] = okb ? __ldg(reinterpret_cast<const uint4*>(y + rb * C + 32 * s + 8 * t)) : make_uint4(0, 0, 0, 0);
    }
  };
  fetch(warp0);
  for (long long blk = warp0; blk < nblk; blk += nwarps) {
    const long long ra = blk * 16 + g, rb = ra + 8;
    const bool oka = ra < rows, okb = rb < rows;
    uint4 xa[STEPS], xb[STEPS];
#pragma unroll
    for (int s = 0; s < STEPS; ++s) { xa[s] = na[s]; xb[s] = nb[s]; }
    fetch(blk + nwarps);
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
    for (int s = 0; s < STEPS; ++s) {
      const float4* sp = reinterpret_cast<const float4*>(ss + 32 * s + 8 * t);
      float sc[8], sh[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 f = sp[q];
        sc[2 * q] = f.x; sh[2 * q] = f.y; sc[2 * q + 1] = f.z; sh[2 * q + 1] = f.w;
      }
      const uint32_t xw[4] = {xa[s].x, xa[s].y, xa[s].z, xa[s].w};
      const uint32_t yw[4] = {xb[s].x, xb[s].y, xb[s].z, xb[s].w};
      uint32_t pa[4], pb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 a = unpack2(xw[q]), b = unpack2(yw[q]);
        pa[q] = pack2(fmaxf(fmaf(a.x, sc[2 * q], sh[2 * q]), 0.f), fmaxf(fmaf(a.y, sc[2 * q + 1], sh[2 * q + 1]), 0.f));
        pb[q] = pack2(fmaxf(fmaf(b.x, sc[2 * q], sh[2 * q]), 0.f), fmaxf(fmaf(b.y, sc[2 * q + 1], sh[2 * q + 1]), 0.f));
      }
      const uint32_t a1[4] = {pa[0], pb[0], pa[1], pb[1]};
      const uint32_t a2[4] = {pa[2], pb[2], pa[3], pb[3]};
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint4 bf = bfr[(s * NT + nt) * 32 + lane];
        mma_bf16_16816(acc[nt], a1, bf.x, bf.y);
        mma_bf16_16816(acc[nt], a2, bf.z, bf.w);
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c0 = nt * 8 + 2 * t;
      if (c0 < NC) {
        const float b0 = __ldg(bias + c0);
        if (oka) z[ra * NC + c0] = acc[nt][0] + b0;
        if (okb) z[rb * NC + c0] = acc[nt][2] + b0;
      }
      if (c0 + 1 < NC) {
        const float b1 = __ldg(bias + c0 + 1);
        if (oka) z[ra * NC + c0 + 1] = acc[nt][1] + b1;
        if (okb) z[rb * NC + c0 + 1] = acc[nt][3] + b1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, pass 2 (apply): one warp per 16 rows, channels in groups of 64.  Inside a group the
// 8 n-tiles are interleaved so that lane t owns channels [8t, 8t+8) and [32+8t, 32+8t+8) of the
// group (two 16-byte chunks per row, 64 contiguous bytes per row per load instruction):
//   channel(nt, col j) = base(j/2) + 2 nt + (j & 1)   with base(t) = 8t for 2nt+e < 8 else 32+8t-8
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int grp_channel(int nt, int j) {
  const int idx = 2 * nt + (j & 1), tt = j >> 1;
  return idx < 8 ? 8 * tt + idx : 32 + 8 * tt + (idx - 8);
}

// v2: a warp iteration covers RB 16-row blocks so each per-channel constant load serves 2 RB
// outputs (v1: short-scoreboard stalls of 18 cycles per issue on three smem loads per output pair).
// (Folding gamma*invstd into the bf16 B fragments saves one more load but makes the recomputed
// da differ from the one cls_bwd_reduce summed; after BatchNorm's mean / xhat projection that
// inconsistency showed up as +45 % gradient error on the stage's conv weight -- not done.)
template <int GROUPS, int RB>   // C = 64 * GROUPS
__global__ void __launch_bounds__(256)
cls_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dz16, const __nv_bfloat16* __restrict__ y,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ mean, const float* __restrict__ invstd,
                     const float* __restrict__ gamma, const float* __restrict__ w,
                     const float* __restrict__ dsum, const float* __restrict__ ddot, float inv_n,
                     __nv_bfloat16* __restrict__ dy, long long rows, int NC) {
  constexpr int C = 64 * GROUPS;
  extern __shared__ __align__(16) uint8_t smem[];
  float4* t1 = reinterpret_cast<float4*>(smem);                       // [C] scale, shift, E, F
  float* t2 = reinterpret_cast<float*>(smem + C * sizeof(float4));    // [C] A = gamma * invstd
  uint2* bfr = reinterpret_cast<uint2*>(smem + C * (sizeof(float4) + sizeof(float)));   // [GROUPS][8][2][32]
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float is = invstd[c];
    const float E = is * ddot[c] * inv_n;                    // xhat*ddot/n = (y - mean) * E
    const float F = dsum[c] * inv_n - mean[c] * E;
    t1[c] = make_float4(scale[c], shift[c], E, F);
    t2[c] = gamma[c] * is;
  }
  for (int i = threadIdx.x; i < GROUPS * 8 * 2 * 32; i += blockDim.x) {
    const int ln = i & 31, ks = (i >> 5) & 1, nt = (i >> 6) & 7, gq = i >> 9;
    const int ch = 64 * gq + grp_channel(nt, ln >> 2);
    const int k = 16 * ks + 2 * (ln & 3);
    auto wv = [&](int kk) { return kk < NC ? w[(size_t)kk * C + ch] : 0.f; };
    bfr[i] = make_uint2(pack2(wv(k), wv(k + 1)), pack2(wv(k + 8), wv(k + 9)));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long nblk = (rows + 16 * RB - 1) / (16 * RB);
  for (long long blk = warp0; blk < nblk; blk += nwarps) {
    long long rr[RB][2];
    bool ok[RB][2];
    uint32_t a[RB][2][4];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      rr[rb][0] = (blk * RB + rb) * 16 + g;
      rr[rb][1] = rr[rb][0] + 8;
      ok[rb][0] = rr[rb][0] < rows;
      ok[rb][1] = rr[rb][1] < rows;
      const uint32_t* pa = reinterpret_cast<const uint32_t*>(dz16 + rr[rb][0] * 32);
      const uint32_t* pb = reinterpret_cast<const uint32_t*>(dz16 + rr[rb][1] * 32);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        a[rb][ks][0] = ok[rb][0] ? __ldg(pa + 8 * ks + t) : 0u;
        a[rb][ks][1] = ok[rb][1] ? __ldg(pb + 8 * ks + t) : 0u;
        a[rb][ks][2] = ok[rb][0] ? __ldg(pa + 8 * ks + 4 + t) : 0u;
        a[rb][ks][3] = ok[rb][1] ? __ldg(pb + 8 * ks + 4 + t) : 0u;
      }
    }
#pragma unroll 1
    for (int gq = 0; gq < GROUPS; ++gq) {
      const int cb = 64 * gq;
      uint4 yv[RB][2][2];     // [row block][row a/b][half of the 64-channel group]
#pragma unroll
      for (int rb = 0; rb < RB; ++rb)
#pragma unroll
        for (int ab = 0; ab < 2; ++ab)
#pragma unroll
          for (int hc = 0; hc < 2; ++hc)
            yv[rb][ab][hc] = ok[rb][ab] ? __ldg(reinterpret_cast<const uint4*>(y + rr[rb][ab] * C + cb + 32 * hc + 8 * t))
                                        : make_uint4(0, 0, 0, 0);
      float acc[RB][8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint2 b0 = bfr[((gq * 8 + nt) * 2 + 0) * 32 + lane];
        const uint2 b1 = bfr[((gq * 8 + nt) * 2 + 1) * 32 + lane];
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[rb][nt][e] = 0.f;
          mma_bf16_16816(acc[rb][nt], a[rb][0], b0.x, b0.y);
          mma_bf16_16816(acc[rb][nt], a[rb][1], b1.x, b1.y);
        }
      }
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {
        uint32_t oo[RB][2][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {   // channels cb + 32 hc + 8 t + 2q (+1)  <->  n-tile 4 hc + q
          const int nt = 4 * hc + q;
          const int ch = cb + 32 * hc + 8 * t + 2 * q;
          const float4 k0 = t1[ch], k1 = t1[ch + 1];
          const float2 A01 = *reinterpret_cast<const float2*>(t2 + ch);
          const float A0 = A01.x, A1 = A01.y;
#pragma unroll
          for (int rb = 0; rb < RB; ++rb) {
            const uint32_t wa = (&yv[rb][0][hc].x)[q], wb = (&yv[rb][1][hc].x)[q];
            const float2 va = unpack2(wa), vb = unpack2(wb);
            const float da00 = fmaf(va.x, k0.x, k0.y) > 0.f ? acc[rb][nt][0] : 0.f;
            const float da01 = fmaf(va.y, k1.x, k1.y) > 0.f ? acc[rb][nt][1] : 0.f;
            const float da10 = fmaf(vb.x, k0.x, k0.y) > 0.f ? acc[rb][nt][2] : 0.f;
            const float da11 = fmaf(vb.y, k1.x, k1.y) > 0.f ? acc[rb][nt][3] : 0.f;
            oo[rb][0][q] = pack2(A0 * (da00 - fmaf(va.x, k0.z, k0.w)), A1 * (da01 - fmaf(va.y, k1.z, k1.w)));
            oo[rb][1][q] = pack2(A0 * (da10 - fmaf(vb.x, k0.z, k0.w)), A1 * (da11 - fmaf(vb.y, k1.z, k1.w)));
          }
        }
#pragma unroll
        for (int rb = 0; rb < RB; ++rb)
#pragma unroll
          for (int ab = 0; ab < 2; ++ab)
            if (ok[rb][ab])
              *reinterpret_cast<uint4*>(dy + rr[rb][ab] * C + cb + 32 * hc + 8 * t) =
                  make_uint4(oo[rb][ab][0], oo[rb][ab][1], oo[rb][ab][2], oo[rb][ab][3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, pass 1 (reduce): block of C/64 warps, warp w owns channels [64w, 64w+64); the block
// walks 32-row tiles staged in smem with cp.async (double buffered); fragments via ldmatrix.
// ---------------------------------------------------------------------------------------------
constexpr int RT = 32;   // rows per tile

template <int WARPS>   // C = 64 * WARPS
__global__ void __launch_bounds__(WARPS * 32)
cls_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dz16, const __nv_bfloat16* __restrict__ y,
                      const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ mean, const float* __restrict__ invstd,
                      const float* __restrict__ w, float* __restrict__ dw, float* __restrict__ dbias,
                      float* __restrict__ dsum, float* __restrict__ ddot, long long rows, int NC,
                      long long tiles_per_block) {
  constexpr int C = 64 * WARPS;
  constexpr int YS = C * 2 + 16;    // row strides in bytes (padding keeps ldmatrix conflict free)
  constexpr int ZS = 64 + 16;
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* ytile = smem;                       // [2][RT][YS]
  uint8_t* ztile = smem + 2 * RT * YS;         // [2][RT][ZS]
  float4* tab = reinterpret_cast<float4*>(smem + 2 * RT * (YS + ZS));   // [C] scale, shift, mean, invstd
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  for (int c = tid; c < C; c += blockDim.x) tab[c] = make_float4(scale[c], shift[c], mean[c], invstd[c]);

  const long long tile0 = (long long)blockIdx.x * tiles_per_block;
  const long long ntiles_all = (rows + RT - 1) / RT;
  const long long tile1 = min(ntiles_all, tile0 + tiles_per_block);
  auto load_tile = [&](long long tile, int buf) {
    const long long r0 = tile * RT;
    for (int i = tid; i < RT * (C / 8); i += WARPS * 32) {
      const int r = i / (C / 8), ck = i % (C / 8);
      const bool ok = r0 + r < rows;
      cp_async16(smem_addr(ytile + (buf * RT + r) * YS + ck * 16), y + (ok ? (r0 + r) : 0) * C + ck * 8, ok);
    }
    for (int i = tid; i < RT * 4; i += WARPS * 32) {
      const int r = i >> 2, ck = i & 3;
      const bool ok = r0 + r < rows;
      cp_async16(smem_addr(ztile + (buf * RT + r) * ZS + ck * 16), dz16 + (ok ? (r0 + r) : 0) * 32 + ck * 8, ok);
    }
    cp_async_commit();
  };
  // W fragments for g = dz @ W over this warp's channels: (k = class, n = channel)
  uint32_t wf[8][2][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int ch = 64 * warp + 8 * nt + g, k = 16 * ks + 2 * t;
      auto wv = [&](int kk) { return kk < NC ? w[(size_t)kk * C + ch] : 0.f; };
      wf[nt][ks][0] = pack2(wv(k), wv(k + 1));
      wf[nt][ks][1] = pack2(wv(k + 8), wv(k + 9));
    }
  // BN constants of the channels this lane sees in the dW B-fragments: channel 64w + 8nt + g
  float bsc[8], bsh[8];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    bsc[nt] = scale[64 * warp + 8 * nt + g];
    bsh[nt] = shift[64 * warp + 8 * nt + g];
  }
  float dwacc[2][8][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dwacc[m][nt][e] = 0.f;
  float ssum[8][2], sdot[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { ssum[nt][0] = ssum[nt][1] = sdot[nt][0] = sdot[nt][1] = 0.f; }
  float sbias = 0.f;

  if (tile0 < tile1) load_tile(tile0, 0);
  __syncthreads();   // tab visible
  for (long long tile = tile0; tile < tile1; ++tile) {
    const int buf = (int)((tile - tile0) & 1);
    if (tile + 1 < tile1) {
      load_tile(tile + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint8_t* yt = ytile + buf * RT * YS;
    const uint8_t* zt = ztile + buf * RT * ZS;
    if (warp == 0) {   // dbias: lane j sums class column j of the tile
      const __nv_bfloat16* zc = reinterpret_cast<const __nv_bfloat16*>(zt) + lane;
#pragma unroll 8
      for (int r = 0; r < RT; ++r) sbias += __bfloat162float(zc[r * (ZS / 2)]);
    }
#pragma unroll
    for (int sub = 0; sub < RT / 16; ++sub) {
      const int r0 = sub * 16;
      // ---- g = dz @ W  (A: rows x classes, straight ldmatrix) ----
      uint32_t az[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        ldmatrix_x4(az[ks], smem_addr(zt + (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ZS + (16 * ks + (lane >> 4) * 8) * 2));
      // ---- dz^T fragments for dW (A: classes x rows, transposed ldmatrix) ----
      uint32_t azt[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
        ldmatrix_x4_trans(azt[m], smem_addr(zt + (r0 + (lane & 7) + (lane >> 4) * 8) * ZS +
                                            (16 * m + ((lane >> 3) & 1) * 8) * 2));
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        // act fragments of n-tiles 2np, 2np+1 (B: rows x channels, transposed ldmatrix) + BN + ReLU
        uint32_t yb[4];
        ldmatrix_x4_trans(yb, smem_addr(yt + (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * YS +
                                        (64 * warp + 16 * np + (lane >> 4) * 8) * 2));
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int nt = 2 * np + q;
          uint32_t bact[2];
#pragma unroll
          for (int hk = 0; hk < 2; ++hk) {
            const float2 v = unpack2(yb[2 * q + hk]);
            bact[hk] = pack2(fmaxf(fmaf(v.x, bsc[nt], bsh[nt]), 0.f), fmaxf(fmaf(v.y, bsc[nt], bsh[nt]), 0.f));
          }
          mma_bf16_16816(dwacc[0][nt], azt[0], bact[0], bact[1]);
          mma_bf16_16816(dwacc[1][nt], azt[1], bact[0], bact[1]);
          // g for this n-tile and the BatchNorm-backward sums
          float gacc[4] = {0.f, 0.f, 0.f, 0.f};
          mma_bf16_16816(gacc, az[0], wf[nt][0][0], wf[nt][0][1]);
          mma_bf16_16816(gacc, az[1], wf[nt][1][0], wf[nt][1][1]);
          const int ch = 64 * warp + 8 * nt + 2 * t;
          const float4 k0 = tab[ch], k1 = tab[ch + 1];
          const float2 va = unpack2(*reinterpret_cast<const uint32_t*>(yt + (r0 + g) * YS + ch * 2));
          const float2 vb = unpack2(*reinterpret_cast<const uint32_t*>(yt + (r0 + g + 8) * YS + ch * 2));
          const float d00 = fmaf(va.x, k0.x, k0.y) > 0.f ? gacc[0] : 0.f;
          const float d01 = fmaf(va.y, k1.x, k1.y) > 0.f ? gacc[1] : 0.f;
          const float d10 = fmaf(vb.x, k0.x, k0.y) > 0.f ? gacc[2] : 0.f;
          const float d11 = fmaf(vb.y, k1.x, k1.y) > 0.f ? gacc[3] : 0.f;
          ssum[nt][0] += d00 + d10;
          ssum[nt][1] += d01 + d11;
          sdot[nt][0] += (d00 * (va.x - k0.z) + d10 * (vb.x - k0.z)) * k0.w;
          sdot[nt][1] += (d01 * (va.y - k1.z) + d11 * (vb.y - k1.z)) * k1.w;
        }
      }
    }
    __syncthreads();   // everyone is done with `buf` before it is refilled
  }
  // ---- flush: BatchNorm sums (reduce over the 8 row-lanes), dW, dbias ----
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float a = ssum[nt][e], b = sdot[nt][e];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (g == 0) {
        const int ch = 64 * warp + 8 * nt + 2 * t + e;
        atomicAdd(dsum + ch, a);
        atomicAdd(ddot + ch, b);
      }
    }
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int ch = 64 * warp + 8 * nt + 2 * t;
      const int j0 = 16 * m + g, j1 = j0 + 8;
      if (j0 < NC) {
        atomicAdd(dw + (size_t)j0 * C + ch, dwacc[m][nt][0]);
        atomicAdd(dw + (size_t)j0 * C + ch + 1, dwacc[m][nt][1]);
      }
      if (j1 < NC) {
        atomicAdd(dw + (size_t)j1 * C + ch, dwacc[m][nt][2]);
        atomicAdd(dw + (size_t)j1 * C + ch + 1, dwacc[m][nt][3]);
      }
    }
  if (warp == 0 && lane < NC) atomicAdd(dbias + lane, sbias);
}

// ---------------------------------------------------------------------------------------------
// logits upsample (bilinear, align_corners=False, integer scale s): z [B,H,W,NC] fp32 ->
// logits [B,NC,H*s,W*s] fp32 NCHW.  One block per (image, output row): the two source rows are
// staged in smem, writes are contiguous along x.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bl_src(int o, int s, int n_in, int& i0, int& i1, float& l) {
  float src = ((float)o + 0.5f) / (float)s - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  i1 = min(i0 + 1, n_in - 1);
  l = src - (float)i0;
}
__device__ __forceinline__ float bl_w(int o, int i, int s, int n_in) {
  int i0, i1;
  float l;
  bl_src(o, s, n_in, i0, i1, l);
  return (i == i0 ? 1.f - l : 0.f) + (i == i1 ? l : 0.f);
}

__global__ void __launch_bounds__(256)
cls_upsample_fwd_kernel(const float* __restrict__ z, float* __restrict__ out, int H, int W, int NC, int s) {
  extern __shared__ __align__(16) float zs[];   // [2][W*NC] source rows, then [OW] (x0 | x1<<16) and [OW] lx
  const int OH = H * s, OW = W * s;
  const int oy = blockIdx.x % OH, b = blockIdx.x / OH;
  int y0, y1;
  float ly;
  bl_src(oy, s, H, y0, y1, ly);
  const int n = W * NC;
  int* xi = reinterpret_cast<int*>(zs + 2 * n);
  float* xl = zs + 2 * n + OW;
  const float* r0 = z + ((size_t)b * H + y0) * n;
  const float* r1 = z + ((size_t)b * H + y1) * n;
  // vertical blend once per source element: zs[i] = (1-ly) z[y0] + ly z[y1]
  for (int i = threadIdx.x; i < n; i += blockDim.x) zs[i] = (1.f - ly) * __ldg(r0 + i) + ly * __ldg(r1 + i);
  for (int ox = threadIdx.x; ox < OW; ox += blockDim.x) {
    int x0, x1;
    float lx;
    bl_src(ox, s, W, x0, x1, lx);
    xi[ox] = x0 | (x1 << 16);
    xl[ox] = lx;
  }
  __syncthreads();
  float* ob = out + (size_t)b * NC * OH * OW + (size_t)oy * OW;
  const bool vec4 = (OW & 3) == 0 && (n & 1) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (int j = threadIdx.x / 32; j < NC; j += blockDim.x / 32) {     // one warp per class row
    float* orow = ob + (size_t)j * OH * OW;
    if (vec4) {
      // four consecutive outputs per lane: one 16-byte store instead of four 4-byte ones
      for (int ox = (threadIdx.x & 31) * 4; ox < OW; ox += 128) {
        const int4 pk = *reinterpret_cast<const int4*>(xi + ox);
        const float4 lx = *reinterpret_cast<const float4*>(xl + ox);
        float4 o;
        o.x = (1.f - lx.x) * zs[(pk.x & 0xffff) * NC + j] + lx.x * zs[(pk.x >> 16) * NC + j];
        o.y = (1.f - lx.y) * zs[(pk.y & 0xffff) * NC + j] + lx.y * zs[(pk.y >> 16) * NC + j];
        o.z = (1.f - lx.z) * zs[(pk.z & 0xffff) * NC + j] + lx.z * zs[(pk.z >> 16) * NC + j];
        o.w = (1.f - lx.w) * zs[(pk.w & 0xffff) * NC + j] + lx.w * zs[(pk.w >> 16) * NC + j];
        *reinterpret_cast<float4*>(orow + ox) = o;
      }
    } else {
      for (int ox = threadIdx.x & 31; ox < OW; ox += 32) {
        const int pk = xi[ox];
        const float lx = xl[ox];
        const float v0 = zs[(pk & 0xffff) * NC + j], v1 = zs[(pk >> 16) * NC + j];
        orow[ox] = (1.f - lx) * v0 + lx * v1;
      }
    }
  }
}

// v2: one block per (image, strip of P source rows) = S*P output rows.  Every source row is read
// from global memory once per strip (cp.async into a 3-row smem ring, the next row in flight under
// the current row's outputs) instead of once per output row by its own block; v1 ran at 2.0 TB/s
// with 15 long-scoreboard stall cycles per issue (load phase and store phase serialised per block).
template <int S>
__global__ void __launch_bounds__(256)
cls_upsample_fwd_strip_kernel(const float* __restrict__ z, float* __restrict__ out, int H, int W, int NC, int P) {
  extern __shared__ __align__(16) float zs[];   // [3][n] ring, [n] blended row, [OW] (x0 | x1<<16), [OW] lx
  const int OH = H * S, OW = W * S;
  const int n = W * NC;                         // floats per source row (host checks n % 4 == 0)
  float* vb = zs + 3 * n;
  int* xi = reinterpret_cast<int*>(vb + n);
  float* xl = vb + n + OW;
  const int strips = (H + P - 1) / P;
  const int strip = blockIdx.x % strips, b = blockIdx.x / strips;
  const int r0 = strip * P, r1 = min(r0 + P, H);
  const float* zb = z + (size_t)b * H * n;
  auto fetch = [&](int r) {                     // source row r (clamped) -> ring slot (r + 1) % 3
    const float* src = zb + (size_t)min(max(r, 0), H - 1) * n;
    float* dst = zs + ((r + 1) % 3) * n;
    for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) cp_async16(smem_addr(dst + i), src + i, true);
    cp_async_commit();
  };
  fetch(r0 - 1);
  fetch(r0);
  for (int ox = threadIdx.x; ox < OW; ox += blockDim.x) {
    int x0, x1;
    float lx;
    bl_src(ox, S, W, x0, x1, lx);
    xi[ox] = x0 | (x1 << 16);
    xl[ox] = lx;
  }
  float* ob = out + (size_t)b * NC * OH * OW;
  for (int r = r0; r < r1; ++r) {
    fetch(r + 1);                               // rows r-1, r resident or in flight; r+1 prefetched
    cp_async_wait<1>();
    __syncthreads();
    const float* up = zs + ((r + 0) % 3) * n;   // slot of row r-1
    const float* mid = zs + ((r + 1) % 3) * n;  // row r
    const float* dn = zs + ((r + 2) % 3) * n;   // row r+1 (complete only for the lower half: waited below)
#pragma unroll
    for (int j = 0; j < S; ++j) {
      const int oy = S * r + j;
      // output row oy: source = r + (j + 0.5)/S - 0.5 -> rows (r-1, r) for j < S/2, (r, r+1) below
      const float ly = (j + 0.5f) / S - 0.5f + (j < S / 2 ? 1.f : 0.f);
      if (j == S / 2) {                         // the lower half needs row r+1
        cp_async_wait<0>();
        __syncthreads();
      }
      const float* a0 = j < S / 2 ? up : mid;
      const float* a1 = j < S / 2 ? mid : dn;
      for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) {
        const float4 p = *reinterpret_cast<const float4*>(a0 + i), q = *reinterpret_cast<const float4*>(a1 + i);
        float4 o;
        o.x = fmaf(ly, q.x - p.x, p.x); o.y = fmaf(ly, q.y - p.y, p.y);
        o.z = fmaf(ly, q.z - p.z, p.z); o.w = fmaf(ly, q.w - p.w, p.w);
        *reinterpret_cast<float4*>(vb + i) = o;
      }
      __syncthreads();
      for (int cls = threadIdx.x / 32; cls < NC; cls += blockDim.x / 32) {     // one warp per class row
        float* orow = ob + ((size_t)cls * OH + oy) * OW;
        for (int ox = (threadIdx.x & 31) * 4; ox < OW; ox += 128) {
          const int4 pk = *reinterpret_cast<const int4*>(xi + ox);
          const float4 lx = *reinterpret_cast<const float4*>(xl + ox);
          float4 o;
          o.x = (1.f - lx.x) * vb[(pk.x & 0xffff) * NC + cls] + lx.x * vb[(pk.x >> 16) * NC + cls];
          o.y = (1.f - lx.y) * vb[(pk.y & 0xffff) * NC + cls] + lx.y * vb[(pk.y >> 16) * NC + cls];
          o.z = (1.f - lx.z) * vb[(pk.z & 0xffff) * NC + cls] + lx.z * vb[(pk.z >> 16) * NC + cls];
          o.w = (1.f - lx.w) * vb[(pk.w & 0xffff) * NC + cls] + lx.w * vb[(pk.w >> 16) * NC + cls];
          *reinterpret_cast<float4*>(orow + ox) = o;
        }
      }
      __syncthreads();                          // vb is rewritten by the next output row
    }
  }
  cp_async_wait<0>();
}

// transpose of the above into the padded bf16 layout the backward MMAs read: dz16 [B*H*W, 32].
// One block per (image, source row).  Exactly 2S output rows / columns carry weight for a source
// row / column (S*i - S/2 .. S*i + 3S/2 - 1).  Phase 1 streams the 2S output rows of every class
// with independent 16-byte loads (all taps of an item in flight) into a per-class row of vertical
// partial sums in smem; phase 2 applies the 2S horizontal taps from a per-block weight table.
template <int S>
__global__ void __launch_bounds__(256)
cls_upsample_bwd_kernel(const float* __restrict__ dout, __nv_bfloat16* __restrict__ dz16, int H, int W,
                        int NC) {
  constexpr int NT = 2 * S;
  extern __shared__ __align__(16) float ts[];   // [NC][OW + 4] partial sums, [W][NT] horizontal weights, [NT] vertical
  const int OH = H * S, OW = W * S;
  const int TS = OW + 4;                        // row pitch: phase 2 reads a column across classes
  const int iy = blockIdx.x % H, b = blockIdx.x / H;
  float* wxs = ts + NC * TS;
  float* wys = wxs + W * NT;
  for (int i = threadIdx.x; i < W * NT; i += blockDim.x) {
    const int ix = i / NT, ox = S * ix - S / 2 + i % NT;
    wxs[i] = (ox >= 0 && ox < OW) ? bl_w(ox, ix, S, W) : 0.f;
  }
  const int oy0 = S * iy - S / 2;
  if (threadIdx.x < NT) {
    const int oy = oy0 + threadIdx.x;
    wys[threadIdx.x] = (oy >= 0 && oy < OH) ? bl_w(oy, iy, S, H) : 0.f;
  }
  __syncthreads();
  float wy[NT];
  size_t roff[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    wy[t] = wys[t];
    roff[t] = (size_t)min(max(oy0 + t, 0), OH - 1) * OW;      // clamped rows carry zero weight
  }
  const float* db = dout + (size_t)b * NC * OH * OW;
  const int ow4 = OW / 4;
  for (int i = threadIdx.x; i < NC * ow4; i += blockDim.x) {
    const int j = i / ow4, o4 = i % ow4;
    const float* plane = db + (size_t)j * OH * OW + o4 * 4;
    float4 d[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) d[t] = __ldg(reinterpret_cast<const float4*>(plane + roff[t]));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      acc.x = fmaf(wy[t], d[t].x, acc.x);
      acc.y = fmaf(wy[t], d[t].y, acc.y);
      acc.z = fmaf(wy[t], d[t].z, acc.z);
      acc.w = fmaf(wy[t], d[t].w, acc.w);
    }
    *reinterpret_cast<float4*>(ts + j * TS + o4 * 4) = acc;
  }
  __syncthreads();
  __nv_bfloat16* orow = dz16 + ((size_t)b * H + iy) * W * 32;
  for (int i = threadIdx.x; i < W * 32; i += blockDim.x) {
    const int j = i & 31, ix = i >> 5;
    float gsum = 0.f;
    if (j < NC) {
      const int o0 = S * ix - S / 2;
#pragma unroll
      for (int tp = 0; tp < NT; ++tp) {
        const int ox = min(max(o0 + tp, 0), OW - 1);
        gsum = fmaf(wxs[ix * NT + tp], ts[j * TS + ox], gsum);
      }
    }
    orow[i] = __float2bfloat16_rn(gsum);
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    s4_set_error("%s: cudaFuncSetAttribute(%zu) failed: %s", what, bytes, cudaGetErrorString(e));
    return S4_ERR_CUDA;
  }
  return S4_OK;
}

}  // namespace

bool s4_cls_tc_supported(int C, int NC, int dtype) {
  return dtype == S4_BF16 && NC >= 1 && NC <= 24 && (C == 64 || C == 128 || C == 256 || C == 512);
}

int s4_cls_fwd_tc(const void* y, const float* scale, const float* shift, const float* w, const float* bias,
                  float* z, long long rows, int C, int NC, cudaStream_t st) {
  const size_t smem = (size_t)C * 8 + (size_t)(C / 32) * 3 * 32 * 16;
  const int grid = (int)std::min<long long>((rows + 127) / 128, (long long)s4_num_sms() * 4);
  int rc;
#define S4_CLS_FWD(STEPS)                                                                         \
  {                                                                                               \
    if ((rc = set_smem(cls_fwd_kernel<STEPS, 3>, smem, "cls_fwd"))) return rc;                    \
    cls_fwd_kernel<STEPS, 3><<<grid, 256, smem, st>>>((const __nv_bfloat16*)y, scale, shift, w,   \
                                                      bias, z, rows, NC);                         \
  }
  if (C == 64) S4_CLS_FWD(2)
  else if (C == 128) S4_CLS_FWD(4)
  else if (C == 256) S4_CLS_FWD(8)
  else S4_CLS_FWD(16)
#undef S4_CLS_FWD
  return s4_check_launch("cls_fwd");
}

int s4_cls_bwd_reduce_tc(const void* dz16, const void* y, const float* scale, const float* shift,
                         const float* mean, const float* invstd, const float* w, float* dw, float* dbias,
                         float* dsum, float* ddot, long long rows, int C, int NC, cudaStream_t st) {
  const int warps = C / 64;
  const size_t smem = (size_t)2 * RT * ((size_t)C * 2 + 16 + 80) + (size_t)C * 16;
  const long long ntiles = (rows + RT - 1) / RT;
  const int per_sm = warps <= 2 ? 4 : (warps <= 4 ? 2 : 1);
  long long blocks = std::min<long long>(ntiles, (long long)s4_num_sms() * per_sm);
  const long long tpb = (ntiles + blocks - 1) / blocks;
  blocks = (ntiles + tpb - 1) / tpb;
  int rc;
#define S4_CLS_RED(WARPS)                                                                          \
  {                                                                                                \
    if ((rc = set_smem(cls_bwd_reduce_kernel<WARPS>, smem, "cls_bwd_reduce"))) return rc;          \
    cls_bwd_reduce_kernel<WARPS><<<(unsigned)blocks, WARPS * 32, smem, st>>>(                      \
        (const __nv_bfloat16*)dz16, (const __nv_bfloat16*)y, scale, shift, mean, invstd, w, dw,    \
        dbias, dsum, ddot, rows, NC, tpb);                                                         \
  }
  if (warps == 1) S4_CLS_RED(1)
  else if (warps == 2) S4_CLS_RED(2)
  else if (warps == 4) S4_CLS_RED(4)
  else S4_CLS_RED(8)
#undef S4_CLS_RED
  return s4_check_launch("cls_bwd_reduce");
}

int s4_cls_bwd_apply_tc(const void* dz16, const void* y, const float* scale, const float* shift,
                        const float* mean, const float* invstd, const float* gamma, const float* w,
                        const float* dsum, const float* ddot, double count, void* dy, long long rows,
                        int C, int NC, cudaStream_t st) {
  const size_t smem = (size_t)C * 20 + (size_t)(C / 64) * 8 * 2 * 32 * 8;
  const int grid = (int)std::min<long long>((rows + 255) / 256, (long long)s4_num_sms() * 4);
  const float inv_n = (float)(1.0 / count);
  int rc;
#define S4_CLS_APP(GROUPS)                                                                         \
  {                                                                                                \
    if ((rc = set_smem(cls_bwd_apply_kernel<GROUPS, 2>, smem, "cls_bwd_apply"))) return rc;        \
    cls_bwd_apply_kernel<GROUPS, 2><<<grid, 256, smem, st>>>(                                       \
        (const __nv_bfloat16*)dz16, (const __nv_bfloat16*)y, scale, shift, mean, invstd, gamma, w, \
        dsum, ddot, inv_n, (__nv_bfloat16*)dy, rows, NC);                                          \
  }
  if (C == 64) S4_CLS_APP(1)
  else if (C == 128) S4_CLS_APP(2)
  else if (C == 256) S4_CLS_APP(4)
  else S4_CLS_APP(8)
#undef S4_CLS_APP
  return s4_check_launch("cls_bwd_apply");
}

int s4_cls_upsample_fwd(const float* z, float* logits, int B, int H, int W, int NC, int s, cudaStream_t st) {
  int rc;
  const size_t n = (size_t)W * NC;
  const size_t smem2 = 4 * n * 4 + (size_t)2 * W * s * 4;
  if ((s == 2 || s == 4) && n % 4 == 0 && (W * s) % 4 == 0 && smem2 <= 200 * 1024 &&
      (((uintptr_t)z | (uintptr_t)logits) & 15) == 0) {
    // strip height: >= 2 blocks per SM in flight (two fit by shared memory), at most 8 rows
    // strip height: ONE wave of blocks (two fit per SM by shared memory), as many as possible
    const int P = (int)std::max<long long>(2, ((long long)B * H + 2 * s4_num_sms() - 1) / (2 * s4_num_sms()));
    const int grid = B * ((H + P - 1) / P);
    if (s == 2) {
      if ((rc = set_smem(cls_upsample_fwd_strip_kernel<2>, smem2, "cls_upsample_fwd"))) return rc;
      cls_upsample_fwd_strip_kernel<2><<<grid, 256, smem2, st>>>(z, logits, H, W, NC, P);
    } else {
      if ((rc = set_smem(cls_upsample_fwd_strip_kernel<4>, smem2, "cls_upsample_fwd"))) return rc;
      cls_upsample_fwd_strip_kernel<4><<<grid, 256, smem2, st>>>(z, logits, H, W, NC, P);
    }
    return s4_check_launch("cls_upsample_fwd");
  }
  const size_t smem = (size_t)2 * W * NC * 4 + (size_t)2 * W * s * 4;
  if ((rc = set_smem(cls_upsample_fwd_kernel, smem, "cls_upsample_fwd"))) return rc;
  cls_upsample_fwd_kernel<<<B * H * s, 256, smem, st>>>(z, logits, H, W, NC, s);
  return s4_check_launch("cls_upsample_fwd");
}

int s4_cls_upsample_bwd(const float* dlogits, void* dz16, int B, int H, int W, int NC, int s, cudaStream_t st) {
  if (s != 2 && s != 4) {
    s4_set_error("cls_upsample_bwd: scale %d not supported (2 or 4)", s);
    return S4_ERR_UNSUPPORTED;
  }
  if ((((uintptr_t)dlogits) & 15) != 0) {
    s4_set_error("cls_upsample_bwd: dlogits must be 16-byte aligned");
    return S4_ERR_ARG;
  }
  const size_t smem = (size_t)NC * (W * s + 4) * 4 + (size_t)W * 2 * s * 4 + (size_t)2 * s * 4;
  int rc;
  if (s == 2) {
    if ((rc = set_smem(cls_upsample_bwd_kernel<2>, smem, "cls_upsample_bwd"))) return rc;
    cls_upsample_bwd_kernel<2><<<B * H, 256, smem, st>>>(dlogits, (__nv_bfloat16*)dz16, H, W, NC);
  } else {
    if ((rc = set_smem(cls_upsample_bwd_kernel<4>, smem, "cls_upsample_bwd"))) return rc;
    cls_upsample_bwd_kernel<4><<<B * H, 256, smem, st>>>(dlogits, (__nv_bfloat16*)dz16, H, W, NC);
  }
  return s4_check_launch("cls_upsample_bwd");
}
