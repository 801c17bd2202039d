// Fused flash-style attention on tcgen05 / TMEM / TMA with S4Former's patch-adaptive (PASA)
// bias applied in-tile (sm_100a).
//
// Reference: mmcv MultiheadAttention -> nn.MultiheadAttention as driven from
// mmseg/models/backbones/vit.py:119 with the additive float mask of vit.py:519-535,
//     bias[b,h,q,k] = w * gate[b,q] * u0[b,k]           (rank 1; same for all heads / layers)
// which is added to the scaled logits inside the softmax warps: no L x L tensor ever exists.
//
// FORWARD (attn_fwd_kernel): persistent, two co-resident CTAs per SM walk (batch, head, 128-query
// tile) work items; TMA K/V rings, S = Q K^T and O += P V on tcgen05 with S, O and P in separate
// TMEM columns, two softmax threads per query row, online softmax with lazy rescaling, the bias
// added in registers.  BACKWARD (attn_bwd_kernel): one CTA per (batch, head, 128-key tile) walking
// the query tiles, all five GEMMs on tcgen05, dQ reduced across key tiles with bulk fp32 adds.
// Details at each kernel.  Both were tuned with the device-side event trace below
// (s4_attention_set_trace / tools/attn_trace.py).
#include <algorithm>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int HD = 64;
constexpr int TILE_BYTES = 128 * HD * 2;   // 16 KB: 128 rows x 64 bf16, 128B-swizzled
constexpr int FWD_THREADS = 384;
constexpr int S_COL = 0;                    // TMEM columns: S fp32 [0,128)
constexpr int O_COL = 128;                  // O fp32 [128,192)
constexpr int P_COL = 192;                  // P bf16x2 [192,256)
constexpr int TMEM_COLS = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 domain: P may grow up to 2^8 before O is rescaled

// Optional device-side event trace (debug / profiling builds of the kernels only: the production
// instantiations compile the hooks away).  Each traced role owns TRACE_MAX (id, clock64) pairs.
constexpr int TRACE_MAX = 512;
struct TraceCfg {
  unsigned long long* buf;   // [roles][TRACE_MAX][2]
  int block;                 // the CTA that records
};
template <bool TR>
struct Tracer {
  unsigned long long* b;
  int n;
  __device__ __forceinline__ Tracer(const TraceCfg& c, int role) : b(nullptr), n(0) {
    if (TR && c.buf && (int)blockIdx.x == c.block) b = c.buf + (size_t)role * TRACE_MAX * 2;
  }
  __device__ __forceinline__ void ev(int id) {
    if (TR && b && n < TRACE_MAX) {
      b[2 * n] = (unsigned long long)id;
      b[2 * n + 1] = (unsigned long long)clock64();
      ++n;
    }
  }
};

struct FwdParams {
  void* out;          // [B, L, H*64] bf16
  float* lse;         // [B, H, L] natural-log LSE of the biased, scaled logits
  const float* u0;    // [B, L] or null
  const float* gate;  // [B, L] or null
  float w;            // bias weight
  float scale;        // 1/sqrt(hd)
  int B, H, L;
  int q_tiles;        // ceil(L / 128)
  int n_full, rem, tail_n;   // key tiles: n_full x 128 + one tail of tail_n (rem valid) keys
  // L = 128 k + 1 (a ViT's cls token + a square patch grid): the one extra key is not worth a
  // tile of its own (a 16-key tail tile costs a full tile's chain of barriers): its logit is a
  // 64-element dot product per query row on the CUDA cores and it is folded into (m, l, O) in the
  // item's epilogue.  `qkv` is the plain pointer behind the tensor map.
  int peel;
  const void* qkv;
  TraceCfg trace;
};

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// FORWARD.  PERSISTENT: 2 x #SM CTAs (two co-resident per SM, so one CTA's softmax overlaps the
// other's MMAs) walk the (batch, head, 128-query tile) work items.  384 threads:
//   warp 0      TMA producer: Q tile once, then K and V tiles (128 keys x 64) through two
//               2-stage rings (cp.async.bulk.tensor.4d, 128B swizzle, zero fill past L)
//   warp 1      MMA issuer + TMEM owner (whole warp in uniform control flow, elected lane issues):
//               S = Q K^T (128 x n x 64, SS) into TMEM cols [0,128);
//               O += P V (128 x 64 x n, P read from TMEM cols [192,256), V from smem) into [128,192)
//   warps 4-11  softmax, TWO threads per query row: warps 4-7 own key columns [0,64) of the tile,
//               warps 8-11 columns [64,128) (a warp reaches the TMEM lanes 32*(warp%4)..+31, so warps
//               w and w+4 share rows).  Halving the per-thread work halves the dependent chain
//               tcgen05.ld -> max -> 64 x ex2 -> tcgen05.st that bounds a tile, and doubles the
//               warps the MUFU / TMEM-load units can be kept busy from.  The two halves exchange
//               their row maxima through shared memory (one 64-thread named barrier per tile), so
//               both use the same reference maximum; online softmax in the log2 domain with LAZY
//               rescaling (O is only rescaled when the running max grows by more than 2^8).
// P lives in its own TMEM columns, so S(j+1) = Q K_{j+1}^T is issued as soon as the softmax warps
// hold S_j in registers (s_free) and runs under tile j's exponentials; O += P_j V_j follows P_j
// (p_full) and releases the P columns / O through pv_done.
// Register budget is moved from warps 0-3 to the softmax warps with setmaxnreg.
// Keys: full 128-key tiles, then either one compact tail tile of round_up(L % 128, 16) keys or -
// for L = 128 k + 1 (cls token + square patch grid: 1025, 2305) - NO tail tile: the one extra key
// is a 64-element dot product per row on the CUDA cores, folded into (m, l, O) in the epilogue.
template <bool TR>
__global__ void __launch_bounds__(FWD_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sK = base + TILE_BYTES;          // 2 stages
  const uint32_t sV = base + 3 * TILE_BYTES;      // 2 stages
  const uint32_t bar = base + 5 * TILE_BYTES;
  const uint32_t q_full = bar;
  auto k_full = [&](int s) { return bar + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bar + 8u * (3 + s); };
  auto v_full = [&](int s) { return bar + 8u * (5 + s); };
  auto v_empty = [&](int s) { return bar + 8u * (7 + s); };
  const uint32_t s_full = bar + 8u * 9;
  const uint32_t p_full = bar + 8u * 10;
  const uint32_t o_final = bar + 8u * 11;
  const uint32_t s_free = bar + 8u * 12;      // every softmax warp has S_j in registers
  const uint32_t pv_done = bar + 8u * 13;     // O += P_j V_j retired (P columns / O reusable)
  const uint32_t q_empty = bar + 8u * 14;     // every S = Q K^T of the work item has retired
  const uint32_t tmem_slot = bar + 8u * 15;
  uint8_t* gen = smem_raw + (base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + 5 * TILE_BYTES + 8 * 15);
  float* xch = reinterpret_cast<float*>(gen + 5 * TILE_BYTES + 128);   // [2 parity][2 halves][128 rows]
  float* xch_l = xch + 512;                                            // [2 halves][128 rows]: row sums
  float* u0s = xch + 768;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_full + (p.tail_n ? 1 : 0);
  // PERSISTENT: the CTA walks work items (batch, head, query tile) blockIdx.x, +gridDim.x, ...;
  // TMEM, barriers and the instruction cache stay warm, and the producer prefetches the next
  // item's Q / K / V while the current item drains.  All barrier parities run on counters that
  // continue across items (`t0` = tiles done before the item, `it` = items done).
  const int n_items = p.B * p.H * p.q_tiles;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tm_qkv);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(k_full(s), 1);
      tc::mbar_init(k_empty(s), 1);
      tc::mbar_init(v_full(s), 1);
      tc::mbar_init(v_empty(s), 1);
    }
    tc::mbar_init(s_full, 1);
    tc::mbar_init(p_full, 8);
    tc::mbar_init(o_final, 1);
    tc::mbar_init(s_free, 8);
    tc::mbar_init(pv_done, 1);
    tc::mbar_init(q_empty, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp < 4) {
    // register budget: the CTA is launched with 80 registers per thread (launch bounds 384 x 2);
    // 128 threads x (80 - 32) released here = 256 threads x (104 - 80) claimed by the softmax warps
    tc::reg_dec<32>();
    if (warp == 0 && lane == 0) {
      // ================================ TMA producer ========================================
      int it = 0, t0 = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it, t0 += n_tiles) {
        const int qt = item % p.q_tiles;
        const int bh = item / p.q_tiles;
        const int h = bh % p.H, b = bh / p.H;
        if (it > 0) tc::mbar_wait(q_empty, (uint32_t)(it - 1) & 1u);
        tc::mbar_expect_tx(q_full, TILE_BYTES);
        tc::tma_load_4d(sQ, &tm_qkv, q_full, 0, qt * BQ, h, b);
        for (int j = 0; j < n_tiles; ++j) {
          const int t = t0 + j;
          const int s = t & 1;
          const uint32_t ph = (uint32_t)(t >> 1) & 1u;
          tc::mbar_wait(k_empty(s), ph ^ 1u);
          tc::mbar_expect_tx(k_full(s), TILE_BYTES);
          tc::tma_load_4d(sK + s * TILE_BYTES, &tm_qkv, k_full(s), 0, j * BKV, p.H + h, b);
          tc::mbar_wait(v_empty(s), ph ^ 1u);
          tc::mbar_expect_tx(v_full(s), TILE_BYTES);
          tc::tma_load_4d(sV + s * TILE_BYTES, &tm_qkv, v_full(s), 0, j * BKV, 2 * p.H + h, b);
        }
      }
    } else if (warp == 1) {
      // ====================== MMA issuer (whole warp in uniform control flow) ================
      // All 32 lanes run the loop and the waits, so every descriptor is a warp-uniform value held
      // in uniform registers; the elected lane issues.  (Issuing from a `lane == 0` branch makes
      // ptxas wrap each tcgen05.mma in a serialising elect loop: ~120 clk per MMA, which made
      // this thread, not the tensor pipe or the softmax warps, the critical path of a tile.)
      const bool leader = tc::elect_one();
      Tracer<TR> tr(p.trace, 0);
      if (!leader) tr.b = nullptr;
      constexpr uint32_t idesc_pv = tc::make_idesc_bf16(BQ, HD, 0, 1);
      constexpr uint32_t idesc_full = tc::make_idesc_bf16(BQ, BKV, 0, 0);
      const uint32_t idesc_tail = tc::make_idesc_bf16(BQ, p.tail_n ? p.tail_n : BKV, 0, 0);
      // descriptor = [hi: SBO 1024 B | version 1 | SWIZZLE_128B] [lo: LBO >> 4 << 16 | addr >> 4]
      constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t q_lo = ((sQ >> 4) & 0x3FFFu) | (1u << 16);
      auto issue_qk = [&](int j, int t, bool last) {
        const int s = t & 1;
        tc::mbar_wait(k_full(s), (uint32_t)(t >> 1) & 1u);
        tc::fence_after_sync();
        const uint32_t idesc_s = (j < p.n_full) ? idesc_full : idesc_tail;
        const uint32_t k_lo = (((sK + s * TILE_BYTES) >> 4) & 0x3FFFu) | (1u << 16);
        if (leader) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            tc::mma_f16_ss(tmem + S_COL, ((uint64_t)desc_hi << 32) | (uint64_t)(q_lo + k * 2),
                           ((uint64_t)desc_hi << 32) | (uint64_t)(k_lo + k * 2), idesc_s, k > 0 ? 1u : 0u);
          tc::mma_commit(k_empty(s));
          tc::mma_commit(s_full);
          if (last) tc::mma_commit(q_empty);     // the Q tile may be overwritten by the next item's
        }
        __syncwarp();
        tr.ev(2);
      };
      int it = 0, t0 = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it, t0 += n_tiles) {
      tc::mbar_wait(q_full, (uint32_t)it & 1u);
      if (t0 > 0) {          // the previous item's last S tile must have been read out
        tc::mbar_wait(s_free, (uint32_t)(t0 - 1) & 1u);
        tc::fence_after_sync();
      }
      issue_qk(0, t0, n_tiles == 1);
      for (int j = 0; j < n_tiles; ++j) {
        const int t = t0 + j;
        const int s = t & 1;
        const bool full = j < p.n_full;
        // S_{j+1} = Q K_{j+1}^T is issued as soon as the softmax warps hold S_j in registers, i.e.
        // it runs on the tensor pipe WHILE they exponentiate tile j (P has its own TMEM columns)
        if (j + 1 < n_tiles) {
          tc::mbar_wait(s_free, (uint32_t)t & 1u);
          tc::fence_after_sync();
          tr.ev(1);
          issue_qk(j + 1, t + 1, j + 2 == n_tiles);
        }
        tc::mbar_wait(p_full, (uint32_t)t & 1u);          // P_j written
        tc::mbar_wait(v_full(s), (uint32_t)(t >> 1) & 1u);
        tc::fence_after_sync();
        tr.ev(3);
        // V tile [128 keys][64 dims] is the MN-major B operand: LBO 16384, one 16-key step = 2048 B
        const uint32_t v_lo = (((sV + s * TILE_BYTES) >> 4) & 0x3FFFu) | ((16384u >> 4) << 16);
        if (leader) {
          if (full) {
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k)
              tc::mma_f16_ts(tmem + O_COL, tmem + P_COL + k * 8,
                             ((uint64_t)desc_hi << 32) | (uint64_t)(v_lo + k * (2048u >> 4)), idesc_pv,
                             (j > 0 || k > 0) ? 1u : 0u);
          } else {
            for (int k = 0; k < p.tail_n / 16; ++k)
              tc::mma_f16_ts(tmem + O_COL, tmem + P_COL + k * 8,
                             ((uint64_t)desc_hi << 32) | (uint64_t)(v_lo + k * (2048u >> 4)), idesc_pv,
                             (j > 0 || k > 0) ? 1u : 0u);
          }
          tc::mma_commit(v_empty(s));
          tc::mma_commit(pv_done);
          if (j == n_tiles - 1) tc::mma_commit(o_final);
        }
        __syncwarp();
        tr.ev(4);
      }
      }
    }
  } else {
    // ================================== softmax warps =======================================
    tc::reg_inc<104>();
    Tracer<TR> tr(p.trace, (lane == 0 && (warp == 4 || warp == 8)) ? (warp == 4 ? 1 : 2) : 0);
    if (!(lane == 0 && (warp == 4 || warp == 8))) tr.b = nullptr;
    const int half = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    const int t256 = threadIdx.x - 128;
    const int pair_bar = 2 + quad;              // warps (4+quad, 8+quad): the two halves of 32 rows
    const float c1 = p.scale * LOG2E;
    const int col0 = half * 64;                 // this thread's key columns of every tile
    bool has_bias = false;
    float wgl = 0.f;
    // the tail tile (L % 128 keys, padded to x16) runs through a compact rolled path that re-reads
    // S from TMEM chunk by chunk: it executes once per item, so its code must stay small (a fully
    // unrolled, predicated copy of the main path cost ~5700 clk per CTA in instruction fetch)
    auto tail_t16 = [&](int j, int c16, int valid, float (&t)[16]) {
      uint32_t rr[16];
      tc::tmem_ld16(lane_base + S_COL + col0 + c16 * 16, rr);
      tc::tmem_ld_wait();
      const float* ut = u0s + j * BKV + col0 + c16 * 16;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float v = __uint_as_float(rr[i]) * c1;
        if (has_bias) v = fmaf(wgl, ut[i], v);
        t[i] = (col0 + c16 * 16 + i < valid) ? v : -INFINITY;
      }
    };
    int it = 0, t0 = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it, t0 += n_tiles) {
    const int qt = item % p.q_tiles;
    const int bh = item / p.q_tiles;
    const int h = bh % p.H, b = bh / p.H;
    const int q = qt * BQ + row;
    const bool q_ok = q < p.L;
    if (p.peel && lane < 2) {
      // pull the peeled key's K row (and this half's V slice) towards the SM now: they are
      // consumed after the Q tile has landed / in the epilogue
      const __nv_bfloat16* kvx = reinterpret_cast<const __nv_bfloat16*>(p.qkv) +
                                 ((size_t)(b * p.L + p.L - 1) * 3 * p.H + (lane == 0 ? p.H : 2 * p.H) + h) * HD;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(kvx));
    }
    has_bias = p.u0 != nullptr;
    if (has_bias) {
      // an all-zero u0 row means "no bias for this image" (the batched student pass mixes biased
      // and unbiased images): detect it while staging the row and take the cheaper path
      const float* ub = p.u0 + (size_t)b * p.L;
      if (it > 0) tc::named_bar_sync(1, 256);   // every softmax thread is done with the old row
      uint32_t nz = 0;
      for (int i = t256; i < n_tiles * BKV; i += 256) {
        const float v = (i < p.L) ? ub[i] : 0.f;
        u0s[i] = v;
        nz |= (v != 0.f) ? 1u : 0u;
      }
      // the peeled key L-1 is not part of the staged row but counts for "has a bias"
      if (p.peel && t256 == 0) nz |= (ub[p.L - 1] != 0.f) ? 1u : 0u;
      uint32_t any;
      asm volatile(
          "{\n"
          ".reg .pred p, q;\n"
          "setp.ne.u32 q, %1, 0;\n"
          "bar.red.or.pred p, 1, 256, q;\n"
          "selp.u32 %0, 1, 0, p;\n"
          "}\n"
          : "=r"(any)
          : "r"(nz)
          : "memory");
      has_bias = any != 0;
    }
    wgl = 0.f;
    if (has_bias) wgl = p.w * LOG2E * ((p.gate && q_ok) ? p.gate[(size_t)b * p.L + q] : 1.f);
    float m_ref = -INFINITY, l_sum = 0.f;
    float t_x = 0.f;          // peeled key L-1: its biased logit in the log2 domain
    tr.ev(20);
    if (p.peel) {
      // (n_full >= 2: the producer cannot overwrite the Q tile before this warp has released
      // S_0 further down, so the rows read here are this item's)
      const int xk = p.L - 1;
      const uint4* kx = reinterpret_cast<const uint4*>(
          reinterpret_cast<const __nv_bfloat16*>(p.qkv) + ((size_t)(b * p.L + xk) * 3 * p.H + p.H + h) * HD);
      // all eight 16-byte pieces of the key row are requested before the first is used (the row was
      // prefetched towards L1 at the top of the item; a rolled loop exposed one L2 latency per piece)
      uint4 kv[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) kv[c] = __ldg(kx + c);
      tc::mbar_wait(q_full, (uint32_t)it & 1u);
      tr.ev(21);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 qv;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(qv.x), "=r"(qv.y), "=r"(qv.z), "=r"(qv.w)
                     : "r"(sQ + row * 128 + ((((uint32_t)c) ^ (uint32_t)(row & 7)) << 4)));
        const __nv_bfloat162* qa = reinterpret_cast<const __nv_bfloat162*>(&qv);
        const __nv_bfloat162* ka = reinterpret_cast<const __nv_bfloat162*>(&kv[c]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 qf = __bfloat1622float2(qa[e]), kf = __bfloat1622float2(ka[e]);
          acc = fmaf(qf.x, kf.x, acc);
          acc = fmaf(qf.y, kf.y, acc);
        }
      }
      t_x = acc * c1;
      if (has_bias) t_x = fmaf(wgl, p.u0[(size_t)b * p.L + xk], t_x);
      tr.ev(22);
    }
    for (int j = 0; j < n_tiles; ++j) {
      const bool full = j < p.n_full;
      const int n = full ? BKV : p.tail_n;
      const int valid = full ? BKV : p.rem;
      const int mine = max(0, min(64, n - col0));          // columns of this half that exist (x16)
      const int t = t0 + j;
      tc::mbar_wait(s_full, (uint32_t)t & 1u);
      tc::fence_after_sync();
      tr.ev(10);
      uint32_t r[64];
      float mt = -INFINITY;
      const bool rawdom = full && !has_bias;     // r[] holds unscaled q.k
      if (full) {
        tc::tmem_ld32(lane_base + S_COL + col0, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tc::tmem_ld32(lane_base + S_COL + col0 + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        tc::tmem_ld_wait();
        // S_j now lives in registers: let the MMA warp overwrite the S columns with S_{j+1}
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(s_free);
        tr.ev(11);
        // ---- logits in the log2 domain + row max (8 independent chains) ----
        const float* ut = u0s + j * BKV + col0;
        float mx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
        if (has_bias) {
#pragma unroll
          for (int c4 = 0; c4 < 16; ++c4) {
            const float4 u4 = *reinterpret_cast<const float4*>(ut + c4 * 4);
            const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = c4 * 4 + i;
              const float t = fmaf(__uint_as_float(r[c]), c1, wgl * uu[i]);
              r[c] = __float_as_uint(t);
              mx[c & 7] = fmaxf(mx[c & 7], t);
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c) mx[c & 7] = fmaxf(mx[c & 7], __uint_as_float(r[c]));
        }
        mt = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])),
                   fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
        if (rawdom) mt *= c1;
      } else {
        for (int c16 = 0; c16 < (mine >> 4); ++c16) {
          float t[16];
          tail_t16(j, c16, valid, t);
#pragma unroll
          for (int i = 0; i < 16; ++i) mt = fmaxf(mt, t[i]);
        }
        tr.ev(11);
      }
      // ---- the two halves of a row agree on the tile maximum ----
      float* xs = xch + (t & 1) * 256;
      xs[half * 128 + row] = mt;
      tc::named_bar_sync(pair_bar, 64);
      tr.ev(12);
      mt = fmaxf(mt, xs[(half ^ 1) * 128 + row]);
      // O += P_{j-1} V_{j-1} must have retired before O is rescaled or the P columns are rewritten
      // (it was issued a whole S-load / max phase ago: this wait is normally already satisfied)
      if (j > 0) {
        tc::mbar_wait(pv_done, (uint32_t)(t - 1) & 1u);
        tc::fence_after_sync();
      }
      tr.ev(13);
      const float m_new = fmaxf(m_ref, mt);
      const bool resc = m_new > m_ref + RESCALE_THRESHOLD;
      float alpha = 1.f;
      if (resc) {
        alpha = tc::ex2(m_ref - m_new);   // 0 on the first tile (m_ref = -inf)
        m_ref = m_new;
        l_sum *= alpha;
      }
      if (j > 0 && __any_sync(0xffffffffu, resc)) {
        // PV(j-1) has retired and PV(j) waits for this tile's P: O is quiescent; each half owns 32 columns
        uint32_t o[32];
        tc::tmem_ld32(lane_base + O_COL + half * 32, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tc::tmem_st32(lane_base + O_COL + half * 32, o);
      }
      // ---- P = 2^(t - m_ref) -> bf16 pairs into the P columns; 4 independent sum chains ----
      const float neg_m = -m_ref;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      if (full) {
        const float mul = rawdom ? c1 : 1.f;
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float p0 = tc::ex2(fmaf(__uint_as_float(r[c16 * 16 + 2 * i]), mul, neg_m));
            const float p1 = tc::ex2(fmaf(__uint_as_float(r[c16 * 16 + 2 * i + 1]), mul, neg_m));
            ls[i & 3] += p0;
            ls[(i + 2) & 3] += p1;
            pk[i] = pack_bf16(p0, p1);
          }
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                       ::"r"(lane_base + P_COL + half * 32 + c16 * 8), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                         "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                       : "memory");
        }
      } else {
        for (int c16 = 0; c16 < (mine >> 4); ++c16) {
          float t[16];
          tail_t16(j, c16, valid, t);
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float p0 = tc::ex2(t[2 * i] + neg_m);
            const float p1 = tc::ex2(t[2 * i + 1] + neg_m);
            ls[i & 3] += p0 + p1;
            pk[i] = pack_bf16(p0, p1);
          }
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                       ::"r"(lane_base + P_COL + half * 32 + c16 * 8), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                         "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                       : "memory");
        }
        // second pass done: the S columns may now be overwritten (by the next item's first tile)
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(s_free);
      }
      l_sum += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      tr.ev(14);
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(p_full);
      tr.ev(15);
    }
    // ---- epilogue: O / l -> bf16, lse; each half writes its 32 of the 64 head dims ----
    uint4 vxv[4];
    if (p.peel) {
      // requested now, consumed after the last PV has retired
      const uint4* vx = reinterpret_cast<const uint4*>(
          reinterpret_cast<const __nv_bfloat16*>(p.qkv) +
          ((size_t)(b * p.L + p.L - 1) * 3 * p.H + 2 * p.H + h) * HD + half * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) vxv[c] = __ldg(vx + c);
    }
    xch_l[half * 128 + row] = l_sum;
    tc::named_bar_sync(pair_bar, 64);
    l_sum += xch_l[(half ^ 1) * 128 + row];
    tr.ev(16);
    tc::mbar_wait(o_final, (uint32_t)it & 1u);
    tc::fence_after_sync();
    tr.ev(17);
    float alpha_x = 1.f, p_x = 0.f;
    if (p.peel) {
      const float m_new = fmaxf(m_ref, t_x);
      alpha_x = tc::ex2(m_ref - m_new);
      p_x = tc::ex2(t_x - m_new);
      l_sum = fmaf(l_sum, alpha_x, p_x);
      m_ref = m_new;
    }
    const float inv_l = 1.f / l_sum;
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) +
                          ((size_t)b * p.L + (q_ok ? q : 0)) * (size_t)(p.H * HD) + h * HD + half * 32;
    {
      uint32_t o[32];
      tc::tmem_ld32(lane_base + O_COL + half * 32, o);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[8 * i + e]);
        if (p.peel) {
          // O <- O alpha + p_x V[L-1]
          const __nv_bfloat162* va = reinterpret_cast<const __nv_bfloat162*>(&vxv[i]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 vf = __bfloat1622float2(va[e]);
            f[2 * e] = fmaf(f[2 * e], alpha_x, p_x * vf.x);
            f[2 * e + 1] = fmaf(f[2 * e + 1], alpha_x, p_x * vf.y);
          }
        }
        if (q_ok) {
          uint4 v;
          v.x = pack_bf16(f[0] * inv_l, f[1] * inv_l);
          v.y = pack_bf16(f[2] * inv_l, f[3] * inv_l);
          v.z = pack_bf16(f[4] * inv_l, f[5] * inv_l);
          v.w = pack_bf16(f[6] * inv_l, f[7] * inv_l);
          *reinterpret_cast<uint4*>(orow + i * 8) = v;
        }
      }
    }
    // every warp's O reads are complete before its p_full arrival for the next item's first tile,
    // which is what lets the MMA warp overwrite O
    tc::fence_before_sync();
    tr.ev(18);
    if (q_ok && half == 0) p.lse[((size_t)b * p.H + h) * p.L + q] = (m_ref + log2f(l_sum)) * LN2;
    }   // work items
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, TMEM_COLS);
  }
}


// =============================================================================================
// BACKWARD.  One CTA per (batch, head, 128-key tile), 1 CTA / SM, 512 threads; the CTA walks the
// query tiles.  Everything is computed TRANSPOSED (keys on the TMEM lanes) so that P^T and dS^T
// are directly the TMEM A-operands of the dV / dK MMAs:
//   S^T  = K_j Q_i^T        (128 keys x n queries, K = 64)      TMEM cols [0,128)
//   dP^T = V_j dO_i^T                                          TMEM cols [128,256)
//   P^T  = 2^(S^T c + bias - lse),  dS^T = P^T (dP^T - D)       (softmax warps; bf16 over S / dP)
//   dV_j += P^T dO_i,  dK_j += dS^T Q_i   (A from TMEM)         TMEM cols [256,320), [320,384)
//   dQ_i  = dS K_j     (A = dS^T staged in smem, MN-major)      TMEM cols [384,448)
// dQ tiles are reduced over the key tiles with cp.reduce.async.bulk (fp32 add in L2) into a
// [B,H,q_tiles,128,64] accumulator that a small kernel converts to bf16.
// Warps: 0 TMA producer, 1 MMA issuer (uniform control flow, elected lane), 4-7 / 8-11 softmax for
// query half 0 / 1 (software-pipelined against each other by the issue order; TMEM loads of
// chunk c+1 in flight under chunk c), 12-15 dQ epilogue.  K_j / V_j are staged once into TMEM as
// the A operands of S^T / dP^T; the dS^T staging tile for dQ is double-buffered.
// =============================================================================================
constexpr int BWD_THREADS = 512;
constexpr int DP_COL = 128, DV_COL = 256, DK_COL = 320, DQ_COL = 384;
// K_j and V_j are the A operands of EVERY S^T / dP^T MMA of the CTA: they are copied once into the
// last 64 TMEM columns (bf16 pairs, 32 columns each), so those MMAs read only the 2 KB B operand
// from shared memory (an SS 128x64x16 MMA moves 6 KB per 32-clk slot: shared-memory bound)
constexpr int KT_COL = 448, VT_COL = 480;

struct BwdParams {
  void* dqkv;            // [B, L, 3*H*64] bf16 (dK, dV written here; dQ by the convert kernel)
  float* dq_accum;       // [B, H, q_tiles, 128, 64] fp32, zeroed
  const float* lse;      // [B, H, L]
  const float* delta;    // [B, H, L]  rowsum(dO * O)
  const float* u0;
  const float* gate;
  float w, scale;
  int B, H, L, q_tiles;
  int kv_tiles;          // 128-key tiles walked by CTAs (q_tiles, or q_tiles - 1 when key L-1 is peeled)
  TraceCfg trace;
};

__device__ __forceinline__ void bulk_reduce_add_f32(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}

__device__ __forceinline__ int bwd_half_n(int L, int i, int h) {
  int qn = L - i * 128;
  qn = qn > 128 ? 128 : ((qn + 15) & ~15);
  if (h == 0) return qn < 64 ? qn : 64;
  return qn > 64 ? qn - 64 : 16;    // at least a 16-wide (all masked) MMA keeps both halves in step
}

template <bool TR>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sK = base;
  const uint32_t sV = base + TILE_BYTES;
  const uint32_t sQ = base + 2 * TILE_BYTES;      // 2 stages
  const uint32_t sdO = base + 4 * TILE_BYTES;     // 2 stages
  const uint32_t sdS = base + 6 * TILE_BYTES;     // 2 x 32 KB: [tile parity][2 query halves][128 keys][64 q] bf16
  const uint32_t sdQ = base + 10 * TILE_BYTES;    // 32 KB fp32 staging
  const uint32_t bar = base + 12 * TILE_BYTES;
  const uint32_t kv_full = bar;
  auto qd_full = [&](int s) { return bar + 8u * (1 + s); };
  auto qd_empty = [&](int s) { return bar + 8u * (3 + s); };
  auto sdp_full = [&](int h) { return bar + 8u * (5 + h); };
  auto pds_full = [&](int h) { return bar + 8u * (7 + h); };
  // dQ_i = dS_i K retired: one barrier per dS staging buffer (tile parity), so "tile i-2 done" can
  // never be confused with "tile i-1 done"
  auto dq_done = [&](int par) { return bar + 8u * (9 + par); };
  const uint32_t dq_empty = bar + 8u * 11;
  const uint32_t dvk_full = bar + 8u * 12;
  const uint32_t kvt_full = bar + 8u * 13;    // K_j / V_j copied into TMEM
  const uint32_t tmem_slot = bar + 8u * 14;
  uint8_t* gen = smem_raw + (base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + 12 * TILE_BYTES + 8 * 14);
  const int lpad = p.q_tiles * 128;
  float* lse2s = reinterpret_cast<float*>(gen + 12 * TILE_BYTES + 128);
  float* dlts = lse2s + lpad;
  float* wgls = dlts + lpad;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv_tiles = p.kv_tiles;
  const int jt = blockIdx.x % kv_tiles;
  const int bh = blockIdx.x / kv_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int nq = p.q_tiles;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tm_qkv);
    tc::prefetch_tmap(&tm_do);
    tc::mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(qd_full(s), 1);
      tc::mbar_init(qd_empty(s), 1);
      tc::mbar_init(sdp_full(s), 1);
      tc::mbar_init(pds_full(s), 4);
    }
    tc::mbar_init(dq_done(0), 1);
    tc::mbar_init(dq_done(1), 1);
    tc::mbar_init(dq_empty, 4);
    tc::mbar_init(dvk_full, 1);
    tc::mbar_init(kvt_full, 8);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ================================ TMA producer ========================================
      tc::mbar_expect_tx(kv_full, 2 * TILE_BYTES);
      tc::tma_load_4d(sK, &tm_qkv, kv_full, 0, jt * 128, p.H + h, b);
      tc::tma_load_4d(sV, &tm_qkv, kv_full, 0, jt * 128, 2 * p.H + h, b);
      for (int i = 0; i < nq; ++i) {
        const int st = i & 1;
        tc::mbar_wait(qd_empty(st), (((uint32_t)i >> 1) & 1u) ^ 1u);
        tc::mbar_expect_tx(qd_full(st), 2 * TILE_BYTES);
        tc::tma_load_4d(sQ + st * TILE_BYTES, &tm_qkv, qd_full(st), 0, i * 128, h, b);
        tc::tma_load_4d(sdO + st * TILE_BYTES, &tm_do, qd_full(st), 0, i * 128, h, b);
      }
    }
  } else if (warp == 1) {
    // ====================== MMA issuer (whole warp in uniform control flow) ==================
    // All 32 lanes run the loop and the waits; every descriptor is a warp-uniform value in uniform
    // registers and the elected lane issues (see the forward kernel for why not `lane == 0`).
    {
      const bool leader = tc::elect_one();
      Tracer<TR> tr(p.trace, 0);
      if (!leader) tr.b = nullptr;
      constexpr uint32_t idesc_dvk = tc::make_idesc_bf16(128, HD, 0, 1);
      constexpr uint32_t idesc_dq = tc::make_idesc_bf16(128, HD, 1, 1);
      // descriptor = [hi: SBO 1024 B | version 1 | SWIZZLE_128B] [lo: LBO >> 4 << 16 | addr >> 4]
      constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t KM = 1u << 16;                    // K-major operand: LBO field 1
      constexpr uint32_t MN = (16384u >> 4) << 16;         // MN-major operand: LBO 16384
      auto dsc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
      const uint32_t k_lo = (sK >> 4) & 0x3FFFu, ds_lo = (sdS >> 4) & 0x3FFFu;
      uint32_t acc_dvk = 0;
      auto issue_sdp = [&](int i, int hh) {
        const int st = i & 1;
        const uint32_t idesc = tc::make_idesc_bf16(128, bwd_half_n(p.L, i, hh), 0, 0);
        const uint32_t q_lo = ((sQ + st * TILE_BYTES + hh * 8192) >> 4) & 0x3FFFu;
        const uint32_t do_lo = ((sdO + st * TILE_BYTES + hh * 8192) >> 4) & 0x3FFFu;
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::mma_f16_ts(tmem + hh * 64, tmem + KT_COL + k * 8, dsc((q_lo + k * 2) | KM), idesc,
                           k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::mma_f16_ts(tmem + DP_COL + hh * 64, tmem + VT_COL + k * 8, dsc((do_lo + k * 2) | KM), idesc,
                           k > 0 ? 1u : 0u);
          tc::mma_commit(sdp_full(hh));
        }
        __syncwarp();
      };
      auto issue_dvk = [&](int i, int hh, uint32_t acc) {
        const int st = i & 1;
        const int steps = bwd_half_n(p.L, i, hh) >> 4;
        const uint32_t q_lo = (((sQ + st * TILE_BYTES + hh * 8192) >> 4) & 0x3FFFu) | MN;
        const uint32_t do_lo = (((sdO + st * TILE_BYTES + hh * 8192) >> 4) & 0x3FFFu) | MN;
        if (leader) {
          if (steps == 4) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
              tc::mma_f16_ts(tmem + DV_COL, tmem + hh * 64 + s * 8, dsc(do_lo + s * 128), idesc_dvk,
                             (acc | (uint32_t)s) ? 1u : 0u);
#pragma unroll
            for (int s = 0; s < 4; ++s)
              tc::mma_f16_ts(tmem + DK_COL, tmem + DP_COL + hh * 64 + s * 8, dsc(q_lo + s * 128), idesc_dvk,
                             (acc | (uint32_t)s) ? 1u : 0u);
          } else {
            for (int s = 0; s < steps; ++s)
              tc::mma_f16_ts(tmem + DV_COL, tmem + hh * 64 + s * 8, dsc(do_lo + s * 128), idesc_dvk,
                             (acc | (uint32_t)s) ? 1u : 0u);
            for (int s = 0; s < steps; ++s)
              tc::mma_f16_ts(tmem + DK_COL, tmem + DP_COL + hh * 64 + s * 8, dsc(q_lo + s * 128), idesc_dvk,
                             (acc | (uint32_t)s) ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      tc::mbar_wait(kv_full, 0);
      tc::mbar_wait(kvt_full, 0);
      tc::mbar_wait(qd_full(0), 0);
      tc::fence_after_sync();
      issue_sdp(0, 0);
      issue_sdp(0, 1);
      for (int i = 0; i < nq; ++i) {
        const int st = i & 1;
        const uint32_t ph = (uint32_t)i & 1u;
        // Q_{i+1} / dO_{i+1} were requested a whole tile ago: waiting for them here (idle time of
        // this warp) keeps the wait off the dV/dK -> S^T/dP^T critical path below
        if (i + 1 < nq) tc::mbar_wait(qd_full(st ^ 1), ((uint32_t)(i + 1) >> 1) & 1u);
        tc::mbar_wait(pds_full(0), ph);
        tc::fence_after_sync();
        tr.ev(1);
        issue_dvk(i, 0, acc_dvk);
        tr.ev(2);
        if (i + 1 < nq) {
          tr.ev(3);
          issue_sdp(i + 1, 0);
          tr.ev(4);
        }
        tc::mbar_wait(pds_full(1), ph);
        tc::fence_after_sync();
        tr.ev(5);
        issue_dvk(i, 1, 1u);
        acc_dvk = 1;
        if (leader) tc::mma_commit(qd_empty(st));
        __syncwarp();
        tr.ev(6);
        // S^T / dP^T of the next tile's second half go first: warpgroup 1 restarts without waiting
        // for dQ_i (the dS staging buffer is double-buffered, nobody is blocked on dQ_i)
        if (i + 1 < nq) issue_sdp(i + 1, 1);
        tr.ev(9);
        tc::mbar_wait(dq_empty, ph ^ 1u);
        tc::fence_after_sync();
        tr.ev(7);
        if (leader) {
          const uint32_t dsb = ds_lo + (uint32_t)st * (32768u >> 4);
#pragma unroll
          for (int s = 0; s < 8; ++s)
            tc::mma_f16_ss(tmem + DQ_COL, dsc((dsb + s * 128) | MN), dsc((k_lo + s * 128) | MN), idesc_dq,
                           s > 0 ? 1u : 0u);
          tc::mma_commit(dq_done(st));
        }
        __syncwarp();
        tr.ev(8);
      }
      if (leader) tc::mma_commit(dvk_full);
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 12) {
    // ================================== softmax warps =======================================
    const int hh = (warp - 4) >> 2;             // query half handled by this warpgroup
    const int quad = warp & 3;
    Tracer<TR> tr(p.trace, 1 + hh);
    if (!(lane == 0 && quad == 0)) tr.b = nullptr;
    const int row = quad * 32 + lane;           // key row (TMEM lane)
    const int key = jt * 128 + row;
    const bool key_ok = key < p.L;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    {
      const int t256 = threadIdx.x - 128;
      const float* lse_b = p.lse + ((size_t)b * p.H + h) * p.L;
      const float* dl_b = p.delta + ((size_t)b * p.H + h) * p.L;
      for (int i = t256; i < lpad; i += 256) {
        const bool ok = i < p.L;
        lse2s[i] = ok ? lse_b[i] * LOG2E : INFINITY;
        dlts[i] = ok ? dl_b[i] : 0.f;
        float g = 0.f;
        if (p.u0 && ok) g = p.w * LOG2E * (p.gate ? p.gate[(size_t)b * p.L + i] : 1.f);
        wgls[i] = g;
      }
      tc::named_bar_sync(1, 256);
    }
    {
      // this thread's key row of K_j (warpgroup 0) / V_j (warpgroup 1): 128 B of the 128B-swizzled
      // TMA tile -> 32 TMEM columns of bf16 pairs (the A-operand layout of a kind::f16 MMA)
      tc::mbar_wait(kv_full, 0);
      const uint32_t src = (hh == 0 ? sK : sV) + row * 128;
      uint32_t kv[32];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(kv[4 * c]), "=r"(kv[4 * c + 1]), "=r"(kv[4 * c + 2]), "=r"(kv[4 * c + 3])
                     : "r"(src + ((((uint32_t)c) ^ (uint32_t)(row & 7)) << 4)));
      tc::tmem_st32(lane_base + (hh == 0 ? KT_COL : VT_COL), kv);
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(kvt_full);
    }
    const float u0k = (p.u0 && key_ok) ? p.u0[(size_t)b * p.L + key] : 0.f;
    const float c1 = p.scale * LOG2E;
    const uint32_t ds_row0 = sdS + hh * 16384 + row * 128;
    const uint32_t swz = (uint32_t)(row & 7);
    for (int i = 0; i < nq; ++i) {
      const int n = bwd_half_n(p.L, i, hh);
      const uint32_t ds_row = ds_row0 + (uint32_t)(i & 1) * 32768u;
      tc::mbar_wait(sdp_full(hh), (uint32_t)i & 1u);
      tc::fence_after_sync();
      tr.ev(10);
      const int qbase = i * 128 + hh * 64;
      // one 16-query chunk: P^T = 2^(S^T c + bias - lse), dS^T = P^T (dP^T - D) -> bf16 over S / dP in
      // TMEM (the A operands of the dV / dK MMAs) and dS^T into the swizzled staging tile (dQ MMA)
      auto chunk = [&](int c16, const uint32_t (&sv)[16], const uint32_t (&dv)[16]) {
        uint32_t pp[8], dsp[8];
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          const float4 a4 = *reinterpret_cast<const float4*>(lse2s + qbase + c16 * 16 + v4 * 4);
          const float4 b4 = *reinterpret_cast<const float4*>(dlts + qbase + c16 * 16 + v4 * 4);
          const float4 c4 = *reinterpret_cast<const float4*>(wgls + qbase + c16 * 16 + v4 * 4);
          const float ls[4] = {a4.x, a4.y, a4.z, a4.w};
          const float dd[4] = {b4.x, b4.y, b4.z, b4.w};
          const float wg[4] = {c4.x, c4.y, c4.z, c4.w};
          float pv[4], dsv[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const int c = 4 * v4 + x;
            const float t = fmaf(__uint_as_float(sv[c]), c1, fmaf(wg[x], u0k, -ls[x]));
            const float pe = key_ok ? tc::ex2(t) : 0.f;
            pv[x] = pe;
            dsv[x] = pe * (__uint_as_float(dv[c]) - dd[x]);
          }
          pp[2 * v4] = pack_bf16(pv[0], pv[1]);
          pp[2 * v4 + 1] = pack_bf16(pv[2], pv[3]);
          dsp[2 * v4] = pack_bf16(dsv[0], dsv[1]);
          dsp[2 * v4 + 1] = pack_bf16(dsv[2], dsv[3]);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(lane_base + hh * 64 + c16 * 8), "r"(pp[0]), "r"(pp[1]), "r"(pp[2]), "r"(pp[3]),
                       "r"(pp[4]), "r"(pp[5]), "r"(pp[6]), "r"(pp[7]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(lane_base + DP_COL + hh * 64 + c16 * 8), "r"(dsp[0]), "r"(dsp[1]), "r"(dsp[2]),
                       "r"(dsp[3]), "r"(dsp[4]), "r"(dsp[5]), "r"(dsp[6]), "r"(dsp[7]) : "memory");
        if (c16 == 0 && i > 1) {   // dQ_{i-2} must have consumed this dS staging buffer
          tr.ev(11);
          tc::mbar_wait(dq_done(i & 1), (uint32_t)((i >> 1) - 1) & 1u);
          tr.ev(12);
        }
        // dS^T row (this key) for queries [c16*16, +16): two 16-byte chunks of the swizzled row
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_row + ((((uint32_t)(2 * c16)) ^ swz) << 4)),
                     "r"(dsp[0]), "r"(dsp[1]), "r"(dsp[2]), "r"(dsp[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_row + ((((uint32_t)(2 * c16 + 1)) ^ swz) << 4)),
                     "r"(dsp[4]), "r"(dsp[5]), "r"(dsp[6]), "r"(dsp[7]) : "memory");
      };
      if (n == 64) {
        // full half tile: the TMEM loads of chunk c+1 are in flight while chunk c is computed
        uint32_t sa[16], da[16], sb[16], db[16];
        tc::tmem_ld16(lane_base + hh * 64, sa);
        tc::tmem_ld16(lane_base + DP_COL + hh * 64, da);
        tc::tmem_ld_wait();
        tc::tmem_ld16(lane_base + hh * 64 + 16, sb);
        tc::tmem_ld16(lane_base + DP_COL + hh * 64 + 16, db);
        chunk(0, sa, da);
        tc::tmem_ld_wait();
        tc::tmem_ld16(lane_base + hh * 64 + 32, sa);
        tc::tmem_ld16(lane_base + DP_COL + hh * 64 + 32, da);
        chunk(1, sb, db);
        tc::tmem_ld_wait();
        tc::tmem_ld16(lane_base + hh * 64 + 48, sb);
        tc::tmem_ld16(lane_base + DP_COL + hh * 64 + 48, db);
        chunk(2, sa, da);
        tc::tmem_ld_wait();
        chunk(3, sb, db);
      } else {
        for (int c16 = 0; c16 < (n >> 4); ++c16) {
          uint32_t sv[16], dv[16];
          tc::tmem_ld16(lane_base + hh * 64 + c16 * 16, sv);
          tc::tmem_ld16(lane_base + DP_COL + hh * 64 + c16 * 16, dv);
          tc::tmem_ld_wait();
          chunk(c16, sv, dv);
        }
      }
      tr.ev(13);
      tc::tmem_st_wait();
      tc::fence_proxy_async();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(pds_full(hh));
      tr.ev(14);
    }
    // ---- final: dV (warpgroup 0) / dK (warpgroup 1) -> bf16 rows of dqkv ----
    tc::mbar_wait(dvk_full, 0);
    tc::fence_after_sync();
    const int D3 = 3 * p.H * HD;
    const float osc = hh == 0 ? 1.f : p.scale;
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.dqkv) + ((size_t)b * p.L + (key_ok ? key : 0)) * D3 +
                          (hh == 0 ? 2 * p.H + h : p.H + h) * HD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tc::tmem_ld32(lane_base + (hh == 0 ? DV_COL : DK_COL) + c * 32, o);
      tc::tmem_ld_wait();
      if (key_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(o[8 * i + 0]) * osc, __uint_as_float(o[8 * i + 1]) * osc);
          v.y = pack_bf16(__uint_as_float(o[8 * i + 2]) * osc, __uint_as_float(o[8 * i + 3]) * osc);
          v.z = pack_bf16(__uint_as_float(o[8 * i + 4]) * osc, __uint_as_float(o[8 * i + 5]) * osc);
          v.w = pack_bf16(__uint_as_float(o[8 * i + 6]) * osc, __uint_as_float(o[8 * i + 7]) * osc);
          *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = v;
        }
      }
    }
  } else if (warp >= 12) {
    // ================================== dQ epilogue =========================================
    const int quad = warp & 3;
    Tracer<TR> tr(p.trace, 3);
    if (!(lane == 0 && quad == 0)) tr.b = nullptr;
    const int row = quad * 32 + lane;           // query row of the tile
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t srow = sdQ + row * 256;
    const uint32_t swz = (uint32_t)(row & 15);
    float* acc_base = p.dq_accum + ((size_t)(b * p.H + h) * p.q_tiles) * (128 * 64);
    for (int i = 0; i < nq; ++i) {
      tc::mbar_wait(dq_done(i & 1), (uint32_t)(i >> 1) & 1u);
      tc::fence_after_sync();
      tr.ev(20);
      uint32_t o0[32], o1[32];
      tc::tmem_ld32(lane_base + DQ_COL, o0);
      tc::tmem_ld32(lane_base + DQ_COL + 32, o1);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(dq_empty);
      // the previous tile's bulk reduce must have finished reading the staging buffer
      if (threadIdx.x == 12 * 32) tc::bulk_wait_read<0>();
      tc::named_bar_sync(2, 128);
      tr.ev(21);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((((uint32_t)c) ^ swz) << 4)),
                     "r"(o0[4 * c]), "r"(o0[4 * c + 1]), "r"(o0[4 * c + 2]), "r"(o0[4 * c + 3]) : "memory");
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((((uint32_t)(c + 8)) ^ swz) << 4)),
                     "r"(o1[4 * c]), "r"(o1[4 * c + 1]), "r"(o1[4 * c + 2]), "r"(o1[4 * c + 3]) : "memory");
      tc::fence_proxy_async();
      tc::named_bar_sync(2, 128);
      if (threadIdx.x == 12 * 32) {
        bulk_reduce_add_f32(acc_base + (size_t)i * (128 * 64), sdQ, 128 * 64 * 4);
        tc::bulk_commit();
      }
      tr.ev(22);
    }
    if (threadIdx.x == 12 * 32) tc::bulk_wait<0>();
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, 512);
  }
}

// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]
__global__ void __launch_bounds__(256)
attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                      float* __restrict__ delta, int B, int H, int L) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, q, h)
  const long long total = (long long)B * L * H;
  if (idx >= total) return;
  const int hh = (int)(idx % H);
  const long long bq = idx / H;
  const int q = (int)(bq % L);
  const int b = (int)(bq / L);
  const uint4* a = reinterpret_cast<const uint4*>(dout + (size_t)bq * H * HD + hh * HD);
  const uint4* o = reinterpret_cast<const uint4*>(out + (size_t)bq * H * HD + hh * HD);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 x = a[i], y = o[i];
    const __nv_bfloat162* xs = reinterpret_cast<const __nv_bfloat162*>(&x);
    const __nv_bfloat162* ys = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fx = __bfloat1622float2(xs[k]), fy = __bfloat1622float2(ys[k]);
      acc = fmaf(fx.x, fy.x, acc);
      acc = fmaf(fx.y, fy.y, acc);
    }
  }
  delta[((size_t)b * H + hh) * L + q] = acc;
}

// PEELED KEY (L = 128 k + 1: a ViT's cls token + square patch grid).  Key x = L-1 would be a
// 128-key tile of its own with ONE valid key: a CTA walking all query tiles with full-size MMAs for
// 1/128 of their rows (1 of 9 CTAs at L = 1025: 11 % of the kernel).  Its contribution is a rank-1
// term instead, computed on the CUDA cores in the SAME pass that forms delta = rowsum(dO * O):
//   p[q]  = 2^(c1 Q_q.K_x + w gate_q u0_x log2e - lse_q log2e)   ds[q] = p[q] (dO_q.V_x - delta_q)
//   dV_x  = sum_q p[q] dO_q     dK_x = scale sum_q ds[q] Q_q     dQ_q += ds[q] K_x
// p / ds are rounded to bf16 before use, as the tensor-core path rounds P^T / dS^T.
// A WARP owns one 512-byte column group of the rows (lane l = the 16-byte chunk 32 k + l: 8 head dims
// of head (32 k + l) / 8) and walks DP_ROWS consecutive query rows of one image four at a time, all
// twelve loads of the four rows in flight before the first is consumed (the first version walked
// row by row with 7 warps per SM resident: latency-bound at 19 % of the DRAM rate).  Every load is
// a fully coalesced row segment, the per-head dot products are xor-reductions over aligned 8-lane
// groups, and the sums over queries stay in the lane's registers (its dims never change).  The
// block's row groups are folded through shared memory into the block's slice of
// peel_part[b, qgroup, h, {dV, dK}, 64] (plain stores: deterministic, nothing to zero); ds goes to
// ds_peel[b,h,q] for the dQ term.  The dq convert kernel adds ds_peel * K_x to dQ, sums the slices
// and writes the dK_x / dV_x rows.
constexpr int DP_ROWS_MIN = 8;   // query rows per warp (8 or 16), four at a time (their loads all in flight)
constexpr int DP_RW = 2;         // row groups per block: a block = (chunk groups) x DP_RW warps

__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ void bf16x8_to_f32(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __bfloat1622float2(h2[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}

template <int NCH, int MINB>     // chunk groups (warps across a row): ceil(H * 8 / 32)
__global__ void __launch_bounds__(NCH * DP_RW * 32, MINB)
attn_bwd_delta_peel_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                           const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ lse,
                           const float* __restrict__ u0, const float* __restrict__ gate, float w, float scale,
                           float* __restrict__ delta, float* __restrict__ ds_peel, float* __restrict__ peel_part,
                           int B, int H, int L, int peel, int DP_ROWS) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int k = wib % NCH, rg = wib / NCH;
  const int qgroups = (L + DP_ROWS * DP_RW - 1) / (DP_ROWS * DP_RW);
  const int b = blockIdx.x / qgroups;
  const int q0 = ((blockIdx.x % qgroups) * DP_RW + rg) * DP_ROWS;
  const int D = H * HD, D3 = 3 * D;
  const int nchunks = H * 8;
  const int x = L - 1;
  const int c = lane + 32 * k;
  const bool cok = c < nchunks;
  const int hh = cok ? (c >> 3) : 0;
  float kx[8], vx[8], accv[8], acck[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { kx[e] = 0.f; vx[e] = 0.f; accv[e] = 0.f; acck[e] = 0.f; }
  if (peel && cok) {
    const __nv_bfloat16* xr = qkv + ((size_t)b * L + x) * D3;
    bf16x8_to_f32(__ldg(reinterpret_cast<const uint4*>(xr + D + c * 8)), kx);
    bf16x8_to_f32(__ldg(reinterpret_cast<const uint4*>(xr + 2 * D + c * 8)), vx);
  }
  const float c1 = scale * LOG2E;
  const float u0x = (peel && u0) ? u0[(size_t)b * L + x] : 0.f;
  const size_t bh = (size_t)b * H + hh;
  for (int qi = 0; qi < DP_ROWS; qi += 4) {
    if (q0 + qi >= L) break;                          // warp-uniform
    // every load of four rows is issued before the first is consumed
    uint4 rd[4], ro[4], rq[4];
    float ls[4], wg[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + qi + u;
      const bool ok = cok && q < L;
      const size_t bq = (size_t)b * L + (ok ? q : 0);
      rd[u] = make_uint4(0, 0, 0, 0); ro[u] = rd[u]; rq[u] = rd[u];
      ls[u] = 0.f; wg[u] = 0.f;
      if (ok) {
        rd[u] = __ldg(reinterpret_cast<const uint4*>(dout + bq * D + c * 8));
        ro[u] = __ldg(reinterpret_cast<const uint4*>(out + bq * D + c * 8));
        if (peel) {
          rq[u] = __ldg(reinterpret_cast<const uint4*>(qkv + bq * D3 + c * 8));
          ls[u] = __ldg(lse + bh * L + q) * LOG2E;
          if (u0) wg[u] = w * LOG2E * (gate ? __ldg(gate + bq) : 1.f);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + qi + u;
      if (q >= L) break;                              // warp-uniform
      float dof[8], of[8], qf[8];
      bf16x8_to_f32(rd[u], dof);
      bf16x8_to_f32(ro[u], of);
      float dl = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) dl = fmaf(dof[e], of[e], dl);
      dl = group8_sum(dl);
      if (cok && (lane & 7) == 0) delta[bh * L + q] = dl;
      if (peel) {
        bf16x8_to_f32(rq[u], qf);
        float sd = 0.f, dp = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          sd = fmaf(qf[e], kx[e], sd);
          dp = fmaf(dof[e], vx[e], dp);
        }
        sd = group8_sum(sd);
        dp = group8_sum(dp);
        if (cok) {
          const float pe = tc::ex2(fmaf(sd, c1, fmaf(wg[u], u0x, -ls[u])));
          const float pr = __bfloat162float(__float2bfloat16_rn(pe));
          const float dsr = __bfloat162float(__float2bfloat16_rn(pe * (dp - dl)));
          if ((lane & 7) == 0) ds_peel[bh * L + q] = dsr;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            accv[e] = fmaf(pr, dof[e], accv[e]);
            acck[e] = fmaf(dsr, qf[e], acck[e]);
          }
        }
      }
    }
  }
  if (peel) {
    // the block's row groups are folded through shared memory and the block's partial sums go to
    // its own slice of peel_part (no atomics: the dq convert kernel adds the slices of an image)
    extern __shared__ float red[];              // [DP_RW][H * 128]
    if (cok) {
      float* dst = red + (size_t)rg * H * 128 + (c >> 3) * 128 + (c & 7) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dst[e] = accv[e];
        dst[64 + e] = acck[e];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H * 128; i += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < DP_RW; ++w2) t += red[(size_t)w2 * H * 128 + i];
      peel_part[(size_t)blockIdx.x * H * 128 + i] = t;
    }
  }
}

// dq_accum [B,H,q_tiles,128,64] fp32 (16-byte chunks XOR-swizzled by row) -> dqkv[:, :, h*64 + d] * scale
// (+ the peeled key's ds_peel[b,h,q] * K_x[d] when ds_peel != null)
__global__ void __launch_bounds__(256)
attn_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, int B, int H,
                           int L, int q_tiles, float scale, const float* __restrict__ ds_peel,
                           const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ peel_part,
                           int qgroups, int fold_blocks) {
  if ((int)blockIdx.x < fold_blocks) {
    // the peeled key's own gradient rows dV_x / dK_x (scaled): sum the delta/peel kernel's per-block
    // partials [qgroups][H][{dV, dK}][64] of the image.  These blocks come FIRST in the grid so the
    // short serial sums run under the rest of the kernel.
    const int i = blockIdx.x * 256 + threadIdx.x;           // (b, h, {dV, dK}, d)
    if (i >= B * H * 128) return;
    const int j = i & 127, hh = (i >> 7) % H, b = (i >> 7) / H;
    const float* pp = peel_part + ((size_t)b * qgroups * H + hh) * 128 + j;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    int g = 0;
    for (; g + 4 <= qgroups; g += 4) {
      t0 += pp[(size_t)(g + 0) * H * 128];
      t1 += pp[(size_t)(g + 1) * H * 128];
      t2 += pp[(size_t)(g + 2) * H * 128];
      t3 += pp[(size_t)(g + 3) * H * 128];
    }
    for (; g < qgroups; ++g) t0 += pp[(size_t)g * H * 128];
    const float t = (t0 + t1) + (t2 + t3);
    __nv_bfloat16* row = dqkv + ((size_t)b * L + (L - 1)) * 3 * H * HD;
    if (j < 64) row[(2 * H + hh) * HD + j] = __float2bfloat16_rn(t);
    else row[(H + hh) * HD + (j - 64)] = __float2bfloat16_rn(t * scale);
    return;
  }
  const long long idx = (long long)(blockIdx.x - fold_blocks) * blockDim.x + threadIdx.x;   // (b, q, h, 8-column group)
  const long long total = (long long)B * L * H * 8;
  if (idx >= total) return;
  const int g8 = (int)(idx & 7);
  const int hh = (int)((idx >> 3) % H);
  const long long bq = (idx >> 3) / H;
  const int q = (int)(bq % L);
  const int b = (int)(bq / L);
  const int r = q & 127;
  const float* rowp = acc + (((size_t)(b * H + hh) * q_tiles + (q >> 7)) * 128 + r) * 64;
  float4 lo = *reinterpret_cast<const float4*>(rowp + (((2 * g8) ^ (r & 15)) << 2));
  float4 hi = *reinterpret_cast<const float4*>(rowp + (((2 * g8 + 1) ^ (r & 15)) << 2));
  if (ds_peel) {
    const float dsq = ds_peel[((size_t)b * H + hh) * L + q];
    const uint4 kraw = __ldg(reinterpret_cast<const uint4*>(qkv + ((size_t)b * L + (L - 1)) * 3 * H * HD + (H + hh) * HD + g8 * 8));
    const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kraw);
    const float2 k0 = __bfloat1622float2(k2[0]), k1 = __bfloat1622float2(k2[1]);
    const float2 k2f = __bfloat1622float2(k2[2]), k3 = __bfloat1622float2(k2[3]);
    lo.x = fmaf(dsq, k0.x, lo.x); lo.y = fmaf(dsq, k0.y, lo.y); lo.z = fmaf(dsq, k1.x, lo.z); lo.w = fmaf(dsq, k1.y, lo.w);
    hi.x = fmaf(dsq, k2f.x, hi.x); hi.y = fmaf(dsq, k2f.y, hi.y); hi.z = fmaf(dsq, k3.x, hi.z); hi.w = fmaf(dsq, k3.y, hi.w);
  }
  uint4 v;
  v.x = pack_bf16(lo.x * scale, lo.y * scale);
  v.y = pack_bf16(lo.z * scale, lo.w * scale);
  v.z = pack_bf16(hi.x * scale, hi.y * scale);
  v.w = pack_bf16(hi.z * scale, hi.w * scale);
  *reinterpret_cast<uint4*>(dqkv + (size_t)bq * 3 * H * HD + hh * HD + g8 * 8) = v;
}

TraceCfg g_trace{nullptr, 0};

int env_composed() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("S4_ATTN_COMPOSED");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

}  // namespace

// Debug hook: record a device-side event trace of CTA `block` of the following attention launches
// into `dev_buf` (roles x TRACE_MAX x 2 uint64, zero-filled by the caller); nullptr switches it off.
extern "C" int s4_attention_set_trace(void* dev_buf, int block) {
  g_trace.buf = (unsigned long long*)dev_buf;
  g_trace.block = block;
  return TRACE_MAX;
}

bool s4_attention_tc_fwd_supported(int B, int H, int L, int hd, int dtype) {
  if (env_composed()) return false;
  return dtype == S4_BF16 && hd == HD && L >= 1 && B >= 1 && H >= 1 && (long long)B * H * ((L + BQ - 1) / BQ) < (1ll << 31);
}

bool s4_attention_tc_bwd_supported(int B, int H, int L, int hd, int dtype) {
  return s4_attention_tc_fwd_supported(B, H, L, hd, dtype);
}

// workspace of the fused backward: delta [B,H,L] + dq_accum [B,H,q_tiles,128,64], fp32
size_t s4_attention_tc_bwd_workspace(int B, int H, int L) {
  const size_t qt = (size_t)(L + 127) / 128;
  const size_t qgroups = ((size_t)L + DP_ROWS_MIN * DP_RW - 1) / (DP_ROWS_MIN * DP_RW);
  return 2 * (((size_t)B * H * L * 4 + 255) / 256 * 256) + (size_t)B * H * qt * 128 * 64 * 4 +
         (size_t)B * qgroups * H * 128 * 4;
}

int s4_attention_tc_fwd(const void* qkv, const float* u0, const float* gate, float w, void* out,
                        float* lse, int B, int H, int L, int hd, cudaStream_t st) {
  if ((((uintptr_t)qkv) & 15) || (((uintptr_t)out) & 15)) {
    s4_set_error("attention_tc_fwd: qkv/out must be 16-byte aligned");
    return S4_ERR_ARG;
  }
  const int D = H * HD;
  CUtensorMap tm;
  const uint64_t dims[4] = {(uint64_t)HD, (uint64_t)L, (uint64_t)(3 * H), (uint64_t)B};
  const uint64_t str[3] = {(uint64_t)(3 * D), (uint64_t)HD, (uint64_t)L * 3 * D};
  const uint32_t box[4] = {64, 128, 1, 1};
  int rc = s4_make_tmap_bf16(&tm, qkv, dims, str, box);
  if (rc) return rc;
  FwdParams p{};
  p.out = out; p.lse = lse; p.u0 = u0; p.gate = gate; p.w = w;
  p.scale = 1.0f / sqrtf((float)hd);
  p.B = B; p.H = H; p.L = L;
  p.q_tiles = (L + BQ - 1) / BQ;
  p.n_full = L / BKV;
  p.rem = L % BKV;
  p.qkv = qkv;
  p.peel = (p.rem == 1 && p.n_full >= 2) ? 1 : 0;
  if (p.peel) p.rem = 0;
  p.tail_n = (p.rem + 15) & ~15;
  const int n_tiles = p.n_full + (p.tail_n ? 1 : 0);
  const size_t smem = 1024 + 5 * TILE_BYTES + 128 + 3072 + (size_t)n_tiles * BKV * 4;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      s4_set_error("attention_tc_fwd: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return S4_ERR_CUDA;
    }
    smem_set = smem;
  }
  const double flops = 4.0 * B * H * (double)L * L * HD;
  S4ProfScope prof("attn_fwd_tc", flops, 0, st);
  p.trace = g_trace;
  const int n_items = B * H * p.q_tiles;
  const int grid = std::min(n_items, 2 * s4_num_sms());     // persistent: two co-resident CTAs per SM
  if (g_trace.buf) attn_fwd_kernel<true><<<grid, FWD_THREADS, smem, st>>>(tm, p);
  else attn_fwd_kernel<false><<<grid, FWD_THREADS, smem, st>>>(tm, p);
  return s4_check_launch("attn_fwd_tc");
}

int s4_attention_tc_bwd(const void* dout, const void* qkv, const void* out, const float* lse,
                        const float* u0, const float* gate, float w, void* dqkv, void* ws,
                        size_t ws_bytes, int B, int H, int L, int hd, cudaStream_t st) {
  if ((((uintptr_t)qkv) & 15) || (((uintptr_t)dout) & 15) || (((uintptr_t)out) & 15) || (((uintptr_t)dqkv) & 15) ||
      (((uintptr_t)ws) & 255)) {
    s4_set_error("attention_tc_bwd: pointers must be 16-byte (workspace 256-byte) aligned");
    return S4_ERR_ARG;
  }
  if (ws_bytes < s4_attention_tc_bwd_workspace(B, H, L)) {
    s4_set_error("attention_tc_bwd: workspace too small");
    return S4_ERR_ARG;
  }
  const int D = H * HD;
  const int q_tiles = (L + 127) / 128;
  float* delta = (float*)ws;
  const size_t delta_bytes = ((size_t)B * H * L * 4 + 255) / 256 * 256;
  float* ds_peel = (float*)((char*)ws + delta_bytes);
  float* dq_accum = (float*)((char*)ws + 2 * delta_bytes);
  const bool peel = (L % 128 == 1) && L > 128 && H <= 16;
  const size_t acc_bytes = (size_t)B * H * q_tiles * 128 * 64 * 4;
  CUtensorMap tq, tdo;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)HD, (uint64_t)L, (uint64_t)(3 * H), (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)(3 * D), (uint64_t)HD, (uint64_t)L * 3 * D};
    const uint32_t box[4] = {64, 128, 1, 1};
    if ((rc = s4_make_tmap_bf16(&tq, qkv, dims, str, box))) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)HD, (uint64_t)L, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)D, (uint64_t)HD, (uint64_t)L * D};
    const uint32_t box[4] = {64, 128, 1, 1};
    if ((rc = s4_make_tmap_bf16(&tdo, dout, dims, str, box))) return rc;
  }
  BwdParams p{};
  p.dqkv = dqkv; p.dq_accum = dq_accum; p.lse = lse; p.delta = delta; p.u0 = u0; p.gate = gate;
  p.w = w; p.scale = 1.0f / sqrtf((float)hd);
  p.B = B; p.H = H; p.L = L; p.q_tiles = q_tiles;
  p.kv_tiles = peel ? q_tiles - 1 : q_tiles;
  const size_t smem = 1024 + 12 * TILE_BYTES + 128 + (size_t)3 * q_tiles * 128 * 4;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      s4_set_error("attention_tc_bwd: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return S4_ERR_CUDA;
    }
    smem_set = smem;
  }
  S4ProfScope prof("attn_bwd_tc", 8.0 * B * H * (double)L * L * HD, 0, st);
  float* peel_part = (float*)((char*)dq_accum + acc_bytes);     // [B,qgroups,H,2,64], fully written
  static const int dp_rows = (getenv("S4_DP_ROWS") && atoi(getenv("S4_DP_ROWS")) == 8) ? 8 : 16;
  static const int dp_minb = (getenv("S4_DP_MINB") && atoi(getenv("S4_DP_MINB")) == 3) ? 3 : 2;
  const int qgroups = (L + dp_rows * DP_RW - 1) / (dp_rows * DP_RW);
  cudaError_t e = cudaMemsetAsync(dq_accum, 0, acc_bytes, st);
  if (e != cudaSuccess) {
    s4_set_error("attention_tc_bwd: memset failed: %s", cudaGetErrorString(e));
    return S4_ERR_CUDA;
  }
  if (H <= 16) {
    const int nch = (H * 8 + 31) / 32;
#define S4_DP(N, MB)                                                                                     \
  attn_bwd_delta_peel_kernel<N, MB><<<B * qgroups, N * DP_RW * 32, peel ? (size_t)DP_RW * H * 128 * 4 : 0, st>>>( \
      (const __nv_bfloat16*)dout, (const __nv_bfloat16*)out, (const __nv_bfloat16*)qkv, lse, u0, gate, w, \
      p.scale, delta, ds_peel, peel_part, B, H, L, peel ? 1 : 0, dp_rows)
    if (nch <= 1) S4_DP(1, 2); else if (nch == 2) S4_DP(2, 2);
    else if (nch == 3) { if (dp_minb == 3) S4_DP(3, 3); else S4_DP(3, 2); }
    else S4_DP(4, 2);
#undef S4_DP
    if ((rc = s4_check_launch("attn_bwd_delta"))) return rc;
  } else {
    const long long total = (long long)B * L * H;
    attn_bwd_delta_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        (const __nv_bfloat16*)dout, (const __nv_bfloat16*)out, delta, B, H, L);
    if ((rc = s4_check_launch("attn_bwd_delta"))) return rc;
  }
  p.trace = g_trace;
  if (g_trace.buf) attn_bwd_kernel<true><<<B * H * p.kv_tiles, BWD_THREADS, smem, st>>>(tq, tdo, p);
  else attn_bwd_kernel<false><<<B * H * p.kv_tiles, BWD_THREADS, smem, st>>>(tq, tdo, p);
  if ((rc = s4_check_launch("attn_bwd_tc"))) return rc;
  {
    const long long total = (long long)B * L * H * 8;
    const int fold_blocks = peel ? (B * H * 128 + 255) / 256 : 0;
    attn_bwd_dq_convert_kernel<<<(unsigned)((total + 255) / 256) + fold_blocks, 256, 0, st>>>(
        dq_accum, (__nv_bfloat16*)dqkv, B, H, L, q_tiles, p.scale, peel ? ds_peel : nullptr,
        (const __nv_bfloat16*)qkv, peel_part, qgroups, fold_blocks);
    rc = s4_check_launch("attn_bwd_dq_convert");
  }
  return rc;
}
