// tcgen05 / TMEM / TMA GEMM for sm_100a: the tensor-core path behind s4_gemm, s4_conv3x3_* .
//
// One persistent, warp-specialised kernel (grid = #SMs, 1 CTA/SM), in two flavours:
//   CT = 1  every CTA owns a 128 x BN tile (tcgen05.mma.cta_group::1)
//   CT = 2  a CTA PAIR (cluster of 2 = one TPC) owns a 256 x BN tile: each CTA stages its own 128
//           rows of A and HALF of the B tile, the leader CTA issues tcgen05.mma.cta_group::2 for
//           both, each CTA drains its own 128 x BN accumulator.  The L2 -> SM operand traffic per
//           MMA cycle drops from (16 + BN/8) KB to (16 + BN/16) KB per k-block, which is what
//           bounds these GEMMs (the fabric delivers ~43 B/clk/SM with all SMs pulling; profiles/).
// Roles per CTA:
//   warp 0      TMA producer   (cp.async.bulk.tensor.4d, 128B swizzle, mbarrier complete_tx)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16)
//   warp 2      TMEM allocator (2 accumulator stages of BN fp32 columns)
//   warps 4-11  epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), so the
// epilogue of tile i overlaps the main loop of tile i+1.
//
// Operand forms (bf16, fp32 accumulate):
//   A: K-major [M][K] | MN-major [K][M] | conv window (4-D NHWC map, coordinates shifted per
//      filter tap; TMA zero-fills the halo => implicit GEMM with no im2col buffer)
//   B: K-major [N][K] | MN-major [K][N] | conv window (wgrad: K = pixels)
// Epilogue: alpha, bias[n], gelu'(aux), pre-activation copy, GELU(erf), residual, accumulate,
//           bf16 / fp32 store, or fp32 atomic accumulate (split-K weight gradients).
//
// Replaces the cuBLAS / cuDNN calls PyTorch makes for the reference's nn.Linear /
// nn.MultiheadAttention / Conv2d layers (vit.py:113-127, embed.py:145-153,
// setr_up_head.py:57-64).
#include <mutex>

#include "common.cuh"
#include "gemm_params.h"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                       // 64 bf16 = 128 B = one swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;

enum { OP_KMAJOR = 0, OP_MNMAJOR = 1, OP_CONV_K = 2, OP_CONV_MN = 3 };

struct TcParams {
  int tiles_m, tiles_n, nb, nb2, splits;
  int M, N;
  int kblocks, kb_per_split;
  int a_mode, b_mode;
  // conv geometry (a_mode == OP_CONV_K or b_mode == OP_CONV_MN)
  int cH, cW, cTW, cTH, cblocks;   // cblocks = Cin / 64 (fwd); tile = cTH x cTW pixels
  int rows_valid;                  // rows of the 128-row tile that are real (conv fwd)
  int row_pitch;                   // global rows advanced per m-tile (BM, or rows_valid for conv)
  int a_nobatch;                   // A tensor map has no batch dims (coordinates forced to 0)
  // conv wgrad with W not expressible in 64-pixel k-blocks (48, 96, ... : 768^2 crops): a k-block
  // carries k_rows < 64 pixels (TMA boxes of k_rows rows); rows [k_rows, 64) of every operand
  // sub-tile are zeroed ONCE at kernel start and never written again, so the K = 64 MMAs see zeros.
  int k_rows;                      // 0 / 64 = full k-blocks
  // epilogue
  void* c;
  const float* bias;
  const __nv_bfloat16* aux;
  const __nv_bfloat16* res;
  __nv_bfloat16* pre;
  long long c_sm, c_sn, c_b1, c_b2;
  float alpha;
  int act, accumulate, c_f32, atomic;
  // TMA-staged epilogue (coalesced stores through smem): tmC stores / reduce-adds C, tmX is the
  // one extra [M,N] bf16 operand: 1 = pre-activation store, 2 = residual load, 3 = GELU' input load
  int tma_epi, x_mode;
  // fused column statistics of the OUTPUT (TMA-staged epilogue only): colsum[n] += sum_m C[m,n],
  // colsq[n] += sum_m C[m,n]^2 (fp32 atomics).  Used for bias gradients (the dgrad GEMM that
  // produces dY also emits colsum(dY)) and BatchNorm batch statistics (conv forward).
  float* colsum;
  float* colsq;
  // dynamic tile scheduler: sched[0] = next tile index, sched[1] = CTA groups that have finished
  // (both self-resetting: the last group to leave zeroes them); null = static round-robin
  int* sched;
};

constexpr int EPI_WARP_BYTES = 8192;     // per epilogue warp: OUT[2] + X[2] boxes of 2 KB
constexpr int EPI_BYTES = NUM_EPI_WARPS * EPI_WARP_BYTES;
constexpr int BAR_BYTES = 1024;

constexpr int SMEM_LIMIT = 227 * 1024;

template <int BN, int CT>
struct Cfg {
  static constexpr int BN_CTA = BN / CT;                 // B rows staged by one CTA
  static constexpr int B_STAGE_BYTES = BN_CTA * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int FIXED_BYTES = 1024 /*align*/ + BAR_BYTES + EPI_BYTES;
  static constexpr int STAGES_FIT = (SMEM_LIMIT - FIXED_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
  static constexpr int TMEM_COLS = 2 * BN > 256 ? 512 : (2 * BN > 128 ? 256 : 128);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED_BYTES;
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(BN % (64 * CT) == 0 || (BN % (16 * CT) == 0), "tile width");
};

__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const float* v) {
  uint4 o;
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c);
  o.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float* v) {
  const uint4 o = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

// Column sums of a 32 (lanes = rows) x 32 (registers = columns) tile held one row per lane:
// reduce-scatter butterfly, 31 shuffles; lane l returns the sum of column l.
__device__ __forceinline__ float warp_colsum32(const float (&s)[32], int lane) {
  float a[16], b[8], c[4], d[2];
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2, u1 = lane & 1;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float mine = u16 ? s[16 + i] : s[i], theirs = u16 ? s[i] : s[16 + i];
    a[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float mine = u8 ? a[8 + i] : a[i], theirs = u8 ? a[i] : a[8 + i];
    b[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float mine = u4 ? b[4 + i] : b[i], theirs = u4 ? b[i] : b[4 + i];
    c[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float mine = u2 ? c[2 + i] : c[i], theirs = u2 ? c[i] : c[2 + i];
    d[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, 2);
  }
  const float mine = u1 ? d[1] : d[0], theirs = u1 ? d[0] : d[1];
  return mine + __shfl_xor_sync(0xffffffffu, theirs, 1);
}

struct TileCoord {
  int m0, n0, z1, z2, kb0, kb1;
};

// p.tiles_m counts CTA-GROUP tiles (128*CT rows); `rank` is the CTA's rank in its pair
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int tile, int BN, int CT, int rank) {
  TileCoord t;
  const int split = tile % p.splits;
  int r = tile / p.splits;
  const int nt = r % p.tiles_n;
  r /= p.tiles_n;
  const int mt = r % p.tiles_m;
  const int z = r / p.tiles_m;
  t.m0 = (mt * CT + rank) * BM;
  t.n0 = nt * BN;
  t.z1 = z / p.nb2;
  t.z2 = z % p.nb2;
  t.kb0 = split * p.kb_per_split;
  t.kb1 = min(p.kblocks, t.kb0 + p.kb_per_split);
  return t;
}


// ---- dynamic tile queue -----------------------------------------------------------------------
// Static round-robin tile assignment makes a persistent grid as slow as its slowest CTA: when
// another kernel holds some SMs (NCCL's all-reduce CTAs during the data-parallel backward), the
// CTAs that are not yet resident start a full tile-share late and the launch takes ~2x.  With
// p.sched the LEADER producer takes tiles from a global counter (the atomic for the next tile is
// issued when a tile's loads start, so its latency hides under them) and publishes them through a
// 4-slot shared-memory queue that the other roles (and, for a CTA pair, the peer CTA: the index
// travels by st.async with complete_tx on the peer's queue barrier) consume in order.  CTAs that
// become resident late find the counter exhausted and leave.
constexpr int TQ = 4;
__device__ __forceinline__ uint32_t peer_addr(uint32_t a) {      // same offset in CTA rank 1 of the pair
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 1;" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ void st_async_u32(uint32_t raddr, uint32_t v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];"
               ::"r"(raddr), "r"(v), "r"(rbar) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// A role's view of the tile sequence (whole warp, uniform control flow).
struct TileSeq {
  uint32_t q_base;      // bar_base + 640: full[TQ] (8 B each), empty[TQ], values[TQ] (4 B each)
  int n;                // tiles taken so far
  int step, total;
  bool dyn, pair, peer, rearm;
  __device__ __forceinline__ uint32_t full(int s) const { return q_base + 8u * s; }
  __device__ __forceinline__ uint32_t empty(int s) const { return q_base + 32u + 8u * s; }
  __device__ __forceinline__ uint32_t val(int s) const { return q_base + 64u + 4u * s; }
  // consumer: the n-th tile of this CTA group (static: first + n * step)
  __device__ __forceinline__ int take(int first) {
    if (!dyn) return first + (n++) * step;
    const int s = n & (TQ - 1);
    tc::mbar_wait(full(s), (uint32_t)(n / TQ) & 1u);
    // (broadcast from lane 0: tells the compiler the value is warp-uniform, so the coordinates
    // derived from it stay in uniform registers)
    const int tile = __shfl_sync(0xffffffffu, (int)lds_u32(val(s)), 0);
    if ((threadIdx.x & 31) == 0) {
      if (peer) {
        tc::mbar_arrive_leader(empty(s));
        if (rearm) tc::mbar_expect_tx(full(s), 4);     // the peer producer arms the slot's next use
      } else {
        tc::mbar_arrive(empty(s));
      }
    }
    ++n;
    return tile;
  }
  // leader producer: hand tile number `tile` (index n of the sequence) to everybody else
  __device__ __forceinline__ void publish(int tile) {
    const int s = n & (TQ - 1);
    tc::mbar_wait(empty(s), ((uint32_t)(n / TQ) & 1u) ^ 1u);
    if ((threadIdx.x & 31) == 0) {
      sts_u32(val(s), (uint32_t)tile);
      tc::mbar_arrive(full(s));
      if (pair) st_async_u32(peer_addr(val(s)), (uint32_t)tile, peer_addr(full(s)));
    }
    __syncwarp();
    ++n;
  }
};

// STATS: the epilogue also emits column sums (/ sums of squares) of the output.  A separate
// instantiation: the two reduce-scatter butterflies add ~250 instructions per 32-column chunk,
// and compiled into every GEMM they slowed the plain epilogues down (register count 118 -> 151,
// a longer chunk loop body) by more than the fusion saved.
template <int BN, int CT, bool STATS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmX,
               const TcParams p) {
  using C = Cfg<BN, CT>;
  constexpr bool PAIR = CT == 2;
  const int rank = PAIR ? (int)tc::cluster_ctarank() : 0;
  const int tile_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;            // 1024-B aligned stage ring
  const uint32_t bar_base = base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  const uint32_t epi_base = bar_base + BAR_BYTES;          // 1024-B aligned staging boxes
  auto xbar = [&](int ew, int s) { return bar_base + 512u + 8u * (ew * 2 + s); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n * p.nb * p.splits;
  const bool dyn = p.sched != nullptr;
  // the first tile of the group is requested before anything else so that the round trip runs under
  // the barrier / TMEM set-up
  int first_dyn = 0;
  if (dyn && warp == 0 && lane == 0 && rank == 0) first_dyn = atomicAdd(p.sched, 1);
  TileSeq seq;
  seq.q_base = bar_base + 640u;
  seq.n = 0;
  seq.step = tile_step;
  seq.total = total_tiles;
  seq.dyn = dyn;
  seq.pair = PAIR;
  seq.peer = PAIR && rank == 1;
  seq.rearm = false;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    if (p.tma_epi) {
      tc::prefetch_tmap(&tmC);
      if (p.x_mode) tc::prefetch_tmap(&tmX);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      tc::mbar_init(full_bar(s), 1);
      tc::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(tfull_bar(s), 1);
      tc::mbar_init(tempty_bar(s), NUM_EPI_WARPS * CT);   // leader: both CTAs' epilogue warps
    }
    for (int w = 0; w < NUM_EPI_WARPS; ++w) {
      tc::mbar_init(xbar(w, 0), 1);
      tc::mbar_init(xbar(w, 1), 1);
    }
    for (int s = 0; s < TQ; ++s) {
      tc::mbar_init(seq.full(s), 1);
      // consumers of a published tile: MMA warp + epilogue warps of the leader, producer + epilogue
      // warps of the peer
      tc::mbar_init(seq.empty(s), (1 + NUM_EPI_WARPS) * CT);
    }
    tc::fence_barrier_init();
    if (dyn && PAIR && rank == 1)
      for (int s = 0; s < TQ; ++s) tc::mbar_expect_tx(seq.full(s), 4);    // armed for the first use
  }
  if (warp == 2) {
    if (PAIR) tc::tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    else tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  if (p.k_rows > 0 && p.k_rows < BK) {
    uint4* ring = reinterpret_cast<uint4*>(smem_raw + (base - raw));
    const int n16 = C::STAGES * C::STAGE_BYTES / 16;
    for (int i = threadIdx.x; i < n16; i += NUM_THREADS) ring[i] = make_uint4(0u, 0u, 0u, 0u);
    tc::fence_proxy_async();        // generic-proxy zeros ordered before the TMA / MMA accesses
  }
  tc::fence_before_sync();
  if (PAIR) tc::cluster_sync();     // peer barriers initialised before any remote arrive / TMA signal
  // (CTA barrier in both modes: compute-sanitizer's racecheck does not credit barrier.cluster for the
  // tcgen05.alloc -> shared-memory slot -> read hand-over below and reported it as a hazard)
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================ TMA producer ==========================================
    // The whole warp walks the tiles in uniform control flow (coordinates live in uniform
    // registers); the elected lane issues the TMA instructions.  The per-k-block body is kept to a
    // handful of instructions (it must stay well under the 4 x 128-cycle MMA time of a k-block):
    // all coordinate arithmetic is incremental, no divisions inside the loop.
    {
      const bool leader = tc::elect_one();
      int stage = 0;
      uint32_t phase = 0;
      const int krows = (p.k_rows > 0 && p.k_rows < BK) ? p.k_rows : BK;   // pixels per k-block (wgrad)
      const uint32_t a_bytes = (p.a_mode == OP_CONV_K) ? (uint32_t)(p.cTW * p.cTH * BK * 2)
                                                        : (uint32_t)(A_STAGE_BYTES / BK * krows);
      // pair mode: the leader's barrier counts the bytes landing in BOTH CTAs
      const uint32_t tx_bytes = (a_bytes + (uint32_t)(C::B_STAGE_BYTES / BK * krows)) * CT;
      auto load = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
        if (leader) {
          if (PAIR) tc::tma_load_4d_pair(dst, m, bar, c0, c1, c2, c3);
          else tc::tma_load_4d(dst, m, bar, c0, c1, c2, c3);
        }
      };
      const int nb_off = rank * C::BN_CTA;          // this CTA's slice of the B tile
      const bool publisher = dyn && rank == 0;
      seq.rearm = true;
      int tile;
      if (publisher) {
        tile = __shfl_sync(0xffffffffu, first_dyn, 0);
        seq.publish(tile);
      } else {
        tile = seq.take(tile_first);
      }
      for (; tile < total_tiles;) {
        // the request for the next tile is in flight while this one is loaded
        int next_req = 0;
        if (publisher && lane == 0) next_req = atomicAdd(p.sched, 1);
        auto advance = [&]() {
          if (publisher) {
            tile = __shfl_sync(0xffffffffu, next_req, 0);
            seq.publish(tile);
          } else {
            tile = seq.take(tile_first);
          }
        };
        const TileCoord t = decode_tile(p, tile, BN, CT, rank);
        if (p.a_mode == OP_KMAJOR && p.b_mode == OP_KMAJOR) {
          // ---- plain GEMM, both operands K-major (forward / dgrad linear layers) ----
          int k0 = t.kb0 * BK;
          for (int kb = t.kb0; kb < t.kb1; ++kb, k0 += BK) {
            tc::mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = base + stage * C::STAGE_BYTES;
            if (rank == 0 && leader) tc::mbar_expect_tx(full_bar(stage), tx_bytes);
            load(sa, &tmA, full_bar(stage), k0, t.m0, t.z2, t.z1);
            load(sa + A_STAGE_BYTES, &tmB, full_bar(stage), k0, t.n0 + nb_off, t.z2, t.z1);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
          }
          advance();
          continue;
        }
        // conv forward: the m-tile is a cTH x cTW pixel window of image cb
        int cb = 0, cy0 = 0, cx0 = 0, tap = 0, cblk = 0;
        if (p.a_mode == OP_CONV_K) {
          const int tiles_x = p.cW / p.cTW;
          const int tiles_per_img = (p.cH / p.cTH) * tiles_x;
          const int mt = t.m0 / BM;
          cb = mt / tiles_per_img;
          const int r = mt % tiles_per_img;
          cy0 = (r / tiles_x) * p.cTH;
          cx0 = (r % tiles_x) * p.cTW;
          tap = t.kb0 / p.cblocks;
          cblk = t.kb0 % p.cblocks;
        }
        int tx = tap % 3 - 1, ty = tap / 3 - 1;
        // conv wgrad: k-block = 64 consecutive pixels (a row segment or whole rows) of image wb
        int wb = 0, wy = 0, wx = 0, wdx = 0, wdy = 0;
        if (p.b_mode == OP_CONV_MN) {
          const int pix = t.kb0 * krows;
          const int hw = p.cH * p.cW;
          wb = pix / hw;
          const int r = pix % hw;
          wy = r / p.cW;
          wx = r % p.cW;
          wdx = t.z2 % 3 - 1;
          wdy = t.z2 / 3 - 1;
        }
        const int az2 = p.a_nobatch ? 0 : t.z2, az1 = p.a_nobatch ? 0 : t.z1;
        int k0 = t.kb0 * krows;
        for (int kb = t.kb0; kb < t.kb1; ++kb, k0 += krows) {
          tc::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
          if (rank == 0 && leader) tc::mbar_expect_tx(full_bar(stage), tx_bytes);
          // ---- A ----
          if (p.a_mode == OP_KMAJOR) {
            load(sa, &tmA, full_bar(stage), k0, t.m0, t.z2, t.z1);
          } else if (p.a_mode == OP_MNMAJOR) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              load(sa + j * (BK * 128), &tmA, full_bar(stage), t.m0 + 64 * j, k0, az2, az1);
          } else {  // OP_CONV_K: k-block = (tap, 64-channel chunk)
            load(sa, &tmA, full_bar(stage), cblk * 64, cx0 + tx, cy0 + ty, cb);
            if (++cblk == p.cblocks) {
              cblk = 0;
              if (++tx == 2) { tx = -1; ++ty; }
            }
          }
          // ---- B ----
          if (p.b_mode == OP_KMAJOR) {
            load(sb, &tmB, full_bar(stage), k0, t.n0 + nb_off, t.z2, t.z1);
          } else if (p.b_mode == OP_MNMAJOR) {
#pragma unroll
            for (int j = 0; j < C::BN_CTA / 64; ++j)
              load(sb + j * (BK * 128), &tmB, full_bar(stage), t.n0 + nb_off + 64 * j, k0, t.z2, t.z1);
          } else {  // OP_CONV_MN (wgrad)
#pragma unroll
            for (int j = 0; j < C::BN_CTA / 64; ++j)
              load(sb + j * (BK * 128), &tmB, full_bar(stage), t.n0 + nb_off + 64 * j, wx + wdx,
                   wy + wdy, wb);
            wx += p.cTW;                      // cTW x cTH = 64 pixels per k-block
            if (wx >= p.cW) {
              wx = 0;
              wy += p.cTH;
              if (wy >= p.cH) { wy = 0; ++wb; }
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        advance();
      }
      if (publisher && lane == 0) {
        // this group has taken its terminating index: the last group to get here re-arms the counters
        const int groups = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
        __threadfence();
        if (atomicAdd(p.sched + 1, 1) == groups - 1) {
          p.sched[0] = 0;
          p.sched[1] = 0;
          __threadfence();
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ============================================
    // The WHOLE warp runs the loop and the waits in uniform control flow, so tile coordinates and
    // descriptors are warp-uniform values in uniform registers; the elected lane issues the
    // tcgen05.mma / commit instructions.  (Issuing from a `lane == 0` branch makes ptxas wrap every
    // tcgen05.mma in a serialising elect loop with five R2UR.BROADCASTs: ~100 clk per MMA, as
    // long as a 128x128x16 MMA itself.)
    if (rank == 0) {
      const bool leader = tc::elect_one();
      const int a_mn = (p.a_mode == OP_MNMAJOR) ? 1 : 0;
      const int b_mn = (p.b_mode == OP_MNMAJOR || p.b_mode == OP_CONV_MN) ? 1 : 0;
      const uint32_t idesc = tc::make_idesc_bf16(BM * CT, BN, a_mn, b_mn);
      auto commit = [&](uint32_t bar) {
        if (PAIR) tc::mma_commit_pair(bar);
        else tc::mma_commit(bar);
      };
      // descriptor = [hi: SBO 1024 B | version 1 | SWIZZLE_128B] [lo: LBO << 16 | addr >> 4]
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo_fixed = (a_mn ? (uint32_t)((BK * 128) >> 4) : 1u) << 16;
      const uint32_t b_lo_fixed = (b_mn ? (uint32_t)((BK * 128) >> 4) : 1u) << 16;
      const uint32_t a_kstep = a_mn ? (2048u >> 4) : (32u >> 4);   // per 16-wide k step
      const uint32_t b_kstep = b_mn ? (2048u >> 4) : (32u >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = seq.take(tile_first); tile < total_tiles; tile = seq.take(tile_first)) {
        const TileCoord t = decode_tile(p, tile, BN, CT, rank);
        tc::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        uint32_t accum = 0;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          tc::mbar_wait(full_bar(stage), phase);
          tc::fence_after_sync();
          const uint32_t sa = base + stage * C::STAGE_BYTES;
          const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | a_lo_fixed;
          const uint32_t b_lo = (((sa + A_STAGE_BYTES) >> 4) & 0x3FFFu) | b_lo_fixed;
          if (leader) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + k * a_kstep);
              const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + k * b_kstep);
              if (PAIR) tc::mma_f16_ss_pair(d_tmem, ad, bd, idesc, (k > 0) ? 1u : accum);
              else tc::mma_f16_ss(d_tmem, ad, bd, idesc, (k > 0) ? 1u : accum);
            }
            commit(empty_bar(stage));                       // frees the smem slot when MMAs retire
            if (kb == t.kb1 - 1) commit(tfull_bar(acc));    // accumulator ready
          }
          __syncwarp();
          accum = 1;
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4 && p.tma_epi) {
    // ===================== epilogue, TMA-staged (coalesced through smem) ===================
    // Each warp owns 32 accumulator rows (its TMEM lane quadrant) and half of the tile's
    // 32-column chunks.  Per chunk: tcgen05.ld -> fused math in registers -> the lane's row is
    // written into a 64B-swizzled 32x32 box in smem -> one lane issues the TMA store (or
    // reduce-add).  Residual / GELU' inputs arrive the same way in the other direction,
    // prefetched one chunk ahead.  TMA clips rows >= M and columns >= N.
    // the elected lane (always the same one: the warp is converged at every use) issues the TMA /
    // bulk-group instructions; coordinates are warp-uniform, so no per-instruction elect loop
    const bool el = tc::elect_one();
    const int ew = warp - 4;
    const int quad = warp & 3;
    const int half = ew >> 2;
    constexpr int CHUNKS = BN / 32;
    constexpr int CH_PER_WARP = (CHUNKS + 1) / 2;
    const uint32_t stg = epi_base + ew * EPI_WARP_BYTES;
    auto out_buf = [&](int i) { return stg + (uint32_t)(i & 1) * 2048u; };
    auto x_buf = [&](int i) { return stg + 4096u + (uint32_t)(i & 1) * 2048u; };
    const bool xload = p.x_mode >= 2;
    const uint32_t swz = (uint32_t)((lane >> 1) & 3);
    const uint32_t row_off = (uint32_t)lane * 64u;
    uint32_t xphase0 = 0, xphase1 = 0;
    int xi = 0, si = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = seq.take(tile_first); tile < total_tiles; tile = seq.take(tile_first)) {
      const TileCoord t = decode_tile(p, tile, BN, CT, rank);
      const int c_begin = half * CH_PER_WARP;
      const int c_end = min(CHUNKS, c_begin + CH_PER_WARP);
      const int row0 = (t.m0 / BM) * p.row_pitch + quad * 32;
      const bool rows_ok = quad * 32 < p.rows_valid && row0 < p.M;
      if (xload && rows_ok && el && t.n0 + c_begin * 32 < p.N && c_begin < c_end) {
        tc::mbar_expect_tx(xbar(ew, xi & 1), 2048);
        tc::tma_load_4d(x_buf(xi & 1), &tmX, xbar(ew, xi & 1), t.n0 + c_begin * 32, row0, t.z2, t.z1);
      }
      tc::mbar_wait(tfull_bar(acc), acc_phase);
      tc::fence_after_sync();
#pragma unroll 1
      for (int chunk = c_begin; chunk < c_end; ++chunk) {
        const int col0 = t.n0 + chunk * 32;
        uint32_t r[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + chunk * 32), r);
        tc::tmem_ld_wait();
        if (!rows_ok || col0 >= p.N) continue;
        float v[32];
        const float bcol = (p.bias && col0 + lane < p.N) ? __ldg(p.bias + col0 + lane) : 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(p.alpha, __uint_as_float(r[j]), __shfl_sync(0xffffffffu, bcol, j));
        if (xload) {
          // prefetch the next chunk's operand into the other buffer (its readers are done: syncwarp)
          __syncwarp();
          if (el && chunk + 1 < c_end && col0 + 32 < p.N) {
            tc::mbar_expect_tx(xbar(ew, (xi + 1) & 1), 2048);
            tc::tma_load_4d(x_buf((xi + 1) & 1), &tmX, xbar(ew, (xi + 1) & 1), col0 + 32, row0, t.z2, t.z1);
          }
          if (xi & 1) { tc::mbar_wait(xbar(ew, 1), xphase1); xphase1 ^= 1u; }
          else { tc::mbar_wait(xbar(ew, 0), xphase0); xphase0 ^= 1u; }
          const uint32_t xb = x_buf(xi & 1) + row_off;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                         : "r"(xb + (((uint32_t)c ^ swz) << 4)));
            const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
              if (p.x_mode == 3) {
                v[c * 8 + 2 * i] *= gelu_grad_fast(f.x);
                v[c * 8 + 2 * i + 1] *= gelu_grad_fast(f.y);
              } else {
                v[c * 8 + 2 * i] += f.x;
                v[c * 8 + 2 * i + 1] += f.y;
              }
            }
          }
          ++xi;
        }
        // staging buffers of the chunk before last must have been read by their stores
        if (el) {
          if (p.c_f32) tc::bulk_wait_read<0>();
          else tc::bulk_wait_read<1>();
        }
        __syncwarp();
        if (p.x_mode == 1) {   // pre-activation copy (bf16), then the activation
          const uint32_t pb = x_buf(si & 1) + row_off;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[c * 8 + 2 * i], v[c * 8 + 2 * i + 1]);
              w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pb + (((uint32_t)c ^ swz) << 4)),
                         "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
          }
        }
        if (p.act == S4_ACT_GELU) {
          if (p.c_f32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_bf16(v[j]);
          }
        }
        if (STATS && p.colsum) {
          // rows past the problem / the conv tile contribute nothing (TMA clips their stores)
          const bool lane_ok = quad * 32 + lane < p.rows_valid && row0 + lane < p.M;
          float sv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) sv[j] = lane_ok ? v[j] : 0.f;
          const float cs = warp_colsum32(sv, lane);
          if (col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, cs);
          if (p.colsq) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sv[j] *= sv[j];
            const float cq = warp_colsum32(sv, lane);
            if (col0 + lane < p.N) atomicAdd(p.colsq + col0 + lane, cq);
          }
        }
        if (p.c_f32) {
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            const uint32_t ob = out_buf(hb) + row_off;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ob + (((uint32_t)c ^ swz) << 4)),
                           "r"(__float_as_uint(v[hb * 16 + c * 4])), "r"(__float_as_uint(v[hb * 16 + c * 4 + 1])),
                           "r"(__float_as_uint(v[hb * 16 + c * 4 + 2])), "r"(__float_as_uint(v[hb * 16 + c * 4 + 3]))
                           : "memory");
          }
        } else {
          const uint32_t ob = out_buf(si & 1) + row_off;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[c * 8 + 2 * i], v[c * 8 + 2 * i + 1]);
              w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ob + (((uint32_t)c ^ swz) << 4)),
                         "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
          }
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (el) {
          if (p.c_f32) {
            if (p.accumulate) {
              tc::tma_reduce_add_4d(&tmC, out_buf(0), col0, row0, t.z2, t.z1);
              if (col0 + 16 < p.N) tc::tma_reduce_add_4d(&tmC, out_buf(1), col0 + 16, row0, t.z2, t.z1);
            } else {
              tc::tma_store_4d(&tmC, out_buf(0), col0, row0, t.z2, t.z1);
              if (col0 + 16 < p.N) tc::tma_store_4d(&tmC, out_buf(1), col0 + 16, row0, t.z2, t.z1);
            }
          } else {
            tc::tma_store_4d(&tmC, out_buf(si & 1), col0, row0, t.z2, t.z1);
          }
          if (p.x_mode == 1) tc::tma_store_4d(&tmX, x_buf(si & 1), col0, row0, t.z2, t.z1);
          tc::bulk_commit();
        }
        ++si;
      }
      tc::fence_before_sync();
      __syncwarp();
      if (el) {
        if (PAIR) tc::mbar_arrive_leader(tempty_bar(acc));
        else tc::mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (el) tc::bulk_wait<0>();
  } else if (warp >= 4) {
    // ================================ epilogue (direct stores) ==============================
    const int ew = warp - 4;
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may access
    const int half = ew >> 2;                   // column half handled by this warp
    constexpr int CHUNKS = BN / 32;
    constexpr int CH_PER_WARP = (CHUNKS + 1) / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = seq.take(tile_first); tile < total_tiles; tile = seq.take(tile_first)) {
      const TileCoord t = decode_tile(p, tile, BN, CT, rank);
      tc::mbar_wait(tfull_bar(acc), acc_phase);
      tc::fence_after_sync();
      const int row_in_tile = quad * 32 + lane;
      const int row = (t.m0 / BM) * p.row_pitch + row_in_tile;
      const bool row_ok = row < p.M && row_in_tile < p.rows_valid;
      const size_t zoff = (size_t)t.z1 * p.c_b1 + (size_t)t.z2 * p.c_b2;
      const size_t roff = zoff + (size_t)row * p.c_sm;
#pragma unroll 1
      for (int cc = 0; cc < CH_PER_WARP; ++cc) {
        const int chunk = half * CH_PER_WARP + cc;
        if (chunk >= CHUNKS) break;
        const int col0 = t.n0 + chunk * 32;
        uint32_t r[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + chunk * 32), r);
        tc::tmem_ld_wait();
        if (!row_ok || col0 >= p.N) continue;
        if (p.atomic) {
          float* cp = reinterpret_cast<float*>(p.c) + roff;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) atomicAdd(cp + (size_t)(col0 + j) * p.c_sn, p.alpha * __uint_as_float(r[j]));
          continue;
        }
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int col = col0 + g8 * 8;
          if (col >= p.N) break;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = p.alpha * __uint_as_float(r[g8 * 8 + j]);
          const bool full8 = col + 8 <= p.N;
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (full8 || col + j < p.N) v[j] += __ldg(p.bias + col + j);
          }
          if (full8) {
            if (p.aux) {
              float a[8];
              load8_bf16(p.aux + roff + col, a);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] *= gelu_grad_fast(a[j]);
            }
            if (p.pre) store8_bf16(p.pre + roff + col, v);
            if (p.act == S4_ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = gelu_fast(v[j]);
            }
            if (p.res) {
              float a[8];
              load8_bf16(p.res + roff + col, a);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += a[j];
            }
            if (p.c_f32) {
              float* cp = reinterpret_cast<float*>(p.c) + roff + col;
              if (p.accumulate) {
                const float4 o0 = *reinterpret_cast<const float4*>(cp);
                const float4 o1 = *reinterpret_cast<const float4*>(cp + 4);
                v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w;
                v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
              }
              *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(cp + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
              __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.c) + roff + col;
              if (p.accumulate) {
                float a[8];
                load8_bf16(cp, a);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += a[j];
              }
              store8_bf16(cp, v);
            }
          } else {
            // ragged tail of N: scalar path
            for (int j = 0; j < 8 && col + j < p.N; ++j) {
              const size_t o = roff + col + j;
              float x = v[j];
              if (p.aux) x *= gelu_grad_fast(__bfloat162float(p.aux[o]));
              if (p.pre) p.pre[o] = __float2bfloat16_rn(x);
              if (p.act == S4_ACT_GELU) x = gelu_fast(x);
              if (p.res) x += __bfloat162float(p.res[o]);
              if (p.c_f32) {
                float* cp = reinterpret_cast<float*>(p.c) + o;
                if (p.accumulate) x += *cp;
                *cp = x;
              } else {
                __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.c) + o;
                if (p.accumulate) x += __bfloat162float(*cp);
                *cp = __float2bfloat16_rn(x);
              }
            }
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) tc::mbar_arrive_leader(tempty_bar(acc));
        else tc::mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc::fence_before_sync();
  if (PAIR) tc::cluster_sync();     // no CTA leaves while its peer may still read / signal it
  else __syncthreads();
  if (warp == 2) {
    tc::fence_after_sync();
    if (PAIR) tc::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    else tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)f;
  }
  return fn;
}

// Counters of the dynamic tile scheduler: one pair per (device, stream slot), zeroed once (the kernel
// leaves them zeroed).  Kernels of one stream run one after the other, so they can share a pair;
// Static round-robin is the default (it is 1-2 us per launch faster when the GPU is not shared);
// s4_set_tc_sched(1) / S4_TC_SCHED=1 selects the dynamic scheduler - the data-parallel wrapper does
// when NCCL kernels overlap the backward.  Never allocated during a stream capture.
int g_sched_mode = -1;     // 0 static round-robin, 1 dynamic (global tile counter)
int env_sched_mode() {
  if (g_sched_mode < 0) {
    const char* e = getenv("S4_TC_SCHED");
    g_sched_mode = (e && e[0] == '1') ? 1 : 0;
  }
  return g_sched_mode;
}

int* sched_counters(cudaStream_t stream) {
  if (!env_sched_mode()) return nullptr;
  constexpr int SLOTS = 32, MAXDEV = 16;
  static int* pool[MAXDEV] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) return nullptr;
  if (!pool[dev]) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return nullptr;
    }
    int* ptr = nullptr;
    if (cudaMalloc(&ptr, SLOTS * 2 * sizeof(int) * 16) != cudaSuccess ||
        cudaMemset(ptr, 0, SLOTS * 2 * sizeof(int) * 16) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    pool[dev] = ptr;
  }
  // one pair per stream (exact table, 128 bytes apart); more than SLOTS streams: static scheduling
  static std::mutex mu;
  static cudaStream_t owner[MAXDEV][SLOTS];
  static int used[MAXDEV] = {0};
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < used[dev]; ++i)
    if (owner[dev][i] == stream) return pool[dev] + (size_t)i * 32;
  if (used[dev] >= SLOTS) return nullptr;
  owner[dev][used[dev]] = stream;
  return pool[dev] + (size_t)(used[dev]++) * 32;
}

// p.tiles_m already counts CTA-group tiles (pairs for CT = 2)
template <int BN, int CT, bool STATS>
int launch_bn_s(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tcm, const CUtensorMap& txm,
                const TcParams& p_in, cudaStream_t stream) {
  using C = Cfg<BN, CT>;
  TcParams p = p_in;
  p.sched = sched_counters(stream);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, CT, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) {
      s4_set_error("gemm_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return S4_ERR_CUDA;
    }
    attr_set = true;
  }
  const long long total = (long long)p.tiles_m * p.tiles_n * p.nb * p.splits;
  const int groups = (int)std::min<long long>(total, s4_num_sms() / CT);
  if (CT == 1) {
    gemm_tc_kernel<BN, CT, STATS><<<groups, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, tcm, txm, p);
    return s4_check_launch("gemm_tc");
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CT);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CT;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CT, STATS>, ta, tb, tcm, txm, p);
  if (e != cudaSuccess) {
    s4_set_error("gemm_tc: cluster launch failed: %s", cudaGetErrorString(e));
    return S4_ERR_CUDA;
  }
  return s4_check_launch("gemm_tc");
}

template <int BN, int CT>
int launch_bn(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tcm, const CUtensorMap& txm,
              const TcParams& p, cudaStream_t stream) {
  if (p.colsum) return launch_bn_s<BN, CT, true>(ta, tb, tcm, txm, p, stream);
  return launch_bn_s<BN, CT, false>(ta, tb, tcm, txm, p, stream);
}

int launch_any(int BN, int CT, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tcm,
               const CUtensorMap& txm, const TcParams& p, cudaStream_t stream) {
  if (CT == 2) {
    if (BN == 256) return launch_bn<256, 2>(ta, tb, tcm, txm, p, stream);
    if (BN == 192) return launch_bn<192, 2>(ta, tb, tcm, txm, p, stream);
    return launch_bn<128, 2>(ta, tb, tcm, txm, p, stream);
  }
  if (BN == 256) return launch_bn<256, 1>(ta, tb, tcm, txm, p, stream);
  if (BN == 128) return launch_bn<128, 1>(ta, tb, tcm, txm, p, stream);
  return launch_bn<64, 1>(ta, tb, tcm, txm, p, stream);
}

// Tensor maps of the TMA-staged epilogue: C (and the one extra [M,N] operand) as
// {N, M, nb2, nb1} with 32x32 (bf16) / 16x32 (fp32) boxes of 64-byte rows, 64B swizzle.
int make_epilogue_maps(CUtensorMap* tcm, CUtensorMap* txm, void* c, const void* x, bool c_f32,
                       long long M, long long N, int nb1, int nb2, long long c_sm, long long c_b1,
                       long long c_b2) {
  const uint64_t dims[4] = {(uint64_t)N, (uint64_t)M, (uint64_t)nb2, (uint64_t)nb1};
  const uint64_t dummy = (uint64_t)c_sm * 8;
  const uint64_t str[3] = {(uint64_t)c_sm, nb2 > 1 ? (uint64_t)c_b2 : dummy, nb1 > 1 ? (uint64_t)c_b1 : dummy};
  const uint32_t box_c[4] = {c_f32 ? 16u : 32u, 32, 1, 1};
  int rc = s4_make_tmap(tcm, c, c_f32 ? S4_F32 : S4_BF16, dims, str, box_c, 64);
  if (rc) return rc;
  if (x) {
    const uint32_t box_x[4] = {32, 32, 1, 1};
    rc = s4_make_tmap(txm, x, S4_BF16, dims, str, box_x, 64);
  } else {
    *txm = *tcm;
  }
  return rc;
}

int g_pair_mode = -1;
int env_pair_mode() {   // S4_TC_PAIR: 0 = never, 1 = cost model (default), 2 = whenever legal
  if (g_pair_mode < 0) {
    const char* e = getenv("S4_TC_PAIR");
    g_pair_mode = e ? atoi(e) : 1;
    if (g_pair_mode < 0 || g_pair_mode > 2) g_pair_mode = 1;
  }
  return g_pair_mode;
}

// Tile configuration by a small cost model.  A k-block of a CTA costs
//   max( MMA time = 2*BN clk , operand bytes / ~43 B/clk )
// (the second term is the measured L2 -> SM delivery rate with every SM pulling; it, not the tensor
// pipe, bounds the 128-row tiles), a tile costs kb k-blocks plus a fixed bubble, and the launch costs
// ceil(tiles / CTA groups) tiles.  b_chunked: B is staged as 64-wide MN boxes (wgrad), so a CTA's
// share of the tile must be a multiple of 64 columns.
struct TileCfg { int BN, CT, splits; };

// kblocks: k-blocks of the whole reduction.  splits > 0: fixed by the caller; splits == 0: chosen
// here too (split-K weight gradients) -- a split count whose work items fill whole waves of CTA
// groups beats "as many as possible" (e.g. 9 taps x 33 splits on 74 pairs is 4.01 waves).
TileCfg pick_cfg(long long m_tiles, int N, int kblocks, int nb, int splits, bool b_chunked, bool allow_pair) {
  const int sms = s4_num_sms();
  const double feed = 43.0;
  TileCfg best{N <= 64 ? 64 : (N <= 128 ? 128 : 256), 1, splits > 0 ? splits : 1};
  double best_cost = 1e30;
  const int mode = env_pair_mode();
  struct Cand { int BN, CT; };
  const Cand cands[] = {{256, 1}, {128, 1}, {64, 1}, {256, 2}, {192, 2}, {128, 2}};
  for (const Cand& c : cands) {
    if (c.CT == 2 && (!allow_pair || mode == 0)) continue;
    if (c.CT == 1 && mode == 2 && allow_pair && N > 64) continue;
    if (c.BN == 64 && N > 64) continue;
    if (c.BN == 128 && c.CT == 1 && N <= 64) continue;
    if (c.BN >= 192 && N <= 128) continue;
    if (c.CT == 2 && b_chunked && (c.BN / 2) % 64) continue;
    const long long base = ((m_tiles + c.CT - 1) / c.CT) * ((N + c.BN - 1) / c.BN) * nb;
    const long long groups = sms / c.CT;
    const double bytes = 16384.0 + (double)(c.BN / c.CT) * 128.0;
    const double per_kb = std::max(2.0 * c.BN, bytes / feed);
    const int s_lo = splits > 0 ? splits : 1;
    const int s_hi = splits > 0 ? splits : std::min(kblocks, 96);
    for (int sp = s_lo; sp <= s_hi; ++sp) {
      const int kb_per = (kblocks + sp - 1) / sp;
      const int sp_eff = (kblocks + kb_per - 1) / kb_per;
      const long long tiles = base * sp_eff;
      const long long waves = (tiles + groups - 1) / groups;
      // every extra split adds one fp32 reduce-add pass of the tile through L2
      const double cost = (double)waves * (kb_per * per_kb + 500.0 + (sp_eff > 1 ? 700.0 : 0.0)) *
                          (c.CT == 2 ? 1.02 : 1.0);
      if (cost < best_cost) { best_cost = cost; best = TileCfg{c.BN, c.CT, sp_eff}; }
    }
  }
  return best;
}

bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace

extern "C" int s4_set_tc_sched(int mode) {
  const int prev = env_sched_mode();
  if (mode == 0 || mode == 1) g_sched_mode = mode;
  return prev;
}

extern "C" int s4_set_tc_pair_mode(int mode) {
  const int prev = env_pair_mode();
  if (mode >= 0 && mode <= 2) g_pair_mode = mode;
  return prev;
}

int s4_make_tmap_bf16(CUtensorMap* out, const void* base, const uint64_t dims[4],
                      const uint64_t strides_elems[3], const uint32_t box[4]) {
  return s4_make_tmap(out, base, S4_BF16, dims, strides_elems, box, 128);
}

int s4_make_tmap(CUtensorMap* out, const void* base, int dtype, const uint64_t dims[4],
                 const uint64_t strides_elems[3], const uint32_t box[4], int swizzle_bytes) {
  EncodeFn fn = get_encode_fn();
  if (!fn) {
    s4_set_error("cuTensorMapEncodeTiled not available from the driver");
    return S4_ERR_CUDA;
  }
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  const int esz = dtype == S4_F32 ? 4 : 2;
  for (int i = 0; i < 3; ++i) gstr[i] = strides_elems[i] * esz;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dtype == S4_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                  const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    s4_set_error("cuTensorMapEncodeTiled failed (%d): dims=[%llu,%llu,%llu,%llu] strides=[%llu,%llu,%llu] box=[%u,%u,%u,%u]",
                 (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
                 (unsigned long long)dims[2], (unsigned long long)dims[3],
                 (unsigned long long)strides_elems[0], (unsigned long long)strides_elems[1],
                 (unsigned long long)strides_elems[2], box[0], box[1], box[2], box[3]);
    return S4_ERR_CUDA;
  }
  return S4_OK;
}

static int env_no_tma_epilogue() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("S4_TC_DIRECT_EPILOGUE");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

static int env_tc_disable_mn() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("S4_TC_NO_MNMAJOR");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

bool s4_gemm_tc_supported(const S4GemmParams& p) {
  if (p.dtype != S4_BF16) return false;
  if (p.M < 1 || p.N < 1 || p.K < 1) return false;
  if (!aligned16(p.a) || !aligned16(p.b) || !aligned16(p.c)) return false;
  const bool a_k = p.a_sk == 1, a_mn = p.a_sm == 1 && !a_k;
  const bool b_k = p.b_sk == 1, b_mn = p.b_sn == 1 && !b_k;
  if (!(a_k || a_mn) || !(b_k || b_mn)) return false;
  if ((a_mn || b_mn) && env_tc_disable_mn()) return false;
  const long long a_ld = a_k ? p.a_sm : p.a_sk, b_ld = b_k ? p.b_sn : p.b_sk;
  if (a_ld % 8 || b_ld % 8) return false;
  if (p.nb1 > 1 && (p.a_b1 % 8 || p.b_b1 % 8)) return false;
  if (p.nb2 > 1 && (p.a_b2 % 8 || p.b_b2 % 8)) return false;
  const int cvec = p.c_dtype == S4_F32 ? 4 : 8;
  if (p.c_sm % cvec) return false;
  if (p.nb1 > 1 && p.c_b1 % cvec) return false;
  if (p.nb2 > 1 && p.c_b2 % cvec) return false;
  if ((p.aux && !aligned16(p.aux)) || (p.res && !aligned16(p.res)) || (p.pre && !aligned16(p.pre)))
    return false;
  if ((p.aux || p.res || p.pre) && p.c_sm % 8) return false;
  if ((p.split_k > 1 || p.split_k < 0) && !(p.accumulate && p.c_dtype == S4_F32 && !p.bias && !p.aux &&
                                            !p.res && !p.pre && p.act == S4_ACT_NONE))
    return false;
  if ((long long)p.nb1 * p.nb2 > 65535) return false;
  return true;
}

int s4_gemm_tc_launch(const S4GemmParams& g, cudaStream_t stream) {
  const bool a_mn = g.a_sk != 1, b_mn = g.b_sk != 1;
  const int nb = g.nb1 * g.nb2;
  const int kblocks = (g.K + BK - 1) / BK;
  // split_k: > 1 fixed, 0 / 1 none, < 0 chosen by the cost model (split-K capable problems only)
  int splits = g.split_k > 1 ? g.split_k : 1;
  if (splits > kblocks) splits = kblocks;
  const bool auto_split = g.split_k < 0;
  const TileCfg tcfg = pick_cfg((g.M + BM - 1) / BM, g.N, kblocks, nb, auto_split ? 0 : splits, b_mn, true);
  const int BN = tcfg.BN, CT = tcfg.CT;
  splits = tcfg.splits;
  int kb_per = (kblocks + splits - 1) / splits;
  splits = (kblocks + kb_per - 1) / kb_per;

  CUtensorMap ta, tb;
  int rc;
  {
    // strides of size-1 batch dims only need to be legal (16-B multiples); the coordinate is 0
    const uint64_t ld = a_mn ? g.a_sk : g.a_sm;
    const uint64_t inner = a_mn ? g.M : g.K, outer = a_mn ? g.K : g.M;
    const uint64_t dims[4] = {inner, outer, (uint64_t)g.nb2, (uint64_t)g.nb1};
    const uint64_t s2[3] = {ld, g.nb2 > 1 ? (uint64_t)g.a_b2 : ld * 8, g.nb1 > 1 ? (uint64_t)g.a_b1 : ld * 8};
    const uint32_t box[4] = {64, a_mn ? (uint32_t)BK : (uint32_t)BM, 1, 1};
    if ((rc = s4_make_tmap_bf16(&ta, g.a, dims, s2, box))) return rc;
  }
  {
    const uint64_t ld = b_mn ? g.b_sk : g.b_sn;
    const uint64_t inner = b_mn ? g.N : g.K, outer = b_mn ? g.K : g.N;
    const uint64_t dims[4] = {inner, outer, (uint64_t)g.nb2, (uint64_t)g.nb1};
    uint64_t s2[3] = {ld, g.nb2 > 1 ? (uint64_t)g.b_b2 : ld * 8, g.nb1 > 1 ? (uint64_t)g.b_b1 : ld * 8};
    const uint32_t box[4] = {64, b_mn ? (uint32_t)BK : (uint32_t)(BN / CT), 1, 1};
    if ((rc = s4_make_tmap_bf16(&tb, g.b, dims, s2, box))) return rc;
  }
  TcParams p{};
  p.tiles_m = ((g.M + BM - 1) / BM + CT - 1) / CT;
  p.tiles_n = (g.N + BN - 1) / BN;
  p.nb = nb; p.nb2 = g.nb2; p.splits = splits;
  p.M = g.M; p.N = g.N;
  p.kblocks = kblocks; p.kb_per_split = kb_per;
  p.a_mode = a_mn ? OP_MNMAJOR : OP_KMAJOR;
  p.b_mode = b_mn ? OP_MNMAJOR : OP_KMAJOR;
  p.rows_valid = BM; p.row_pitch = BM;
  p.c = g.c; p.bias = g.bias;
  p.aux = (const __nv_bfloat16*)g.aux; p.res = (const __nv_bfloat16*)g.res; p.pre = (__nv_bfloat16*)g.pre;
  p.c_sm = g.c_sm; p.c_sn = 1; p.c_b1 = g.c_b1; p.c_b2 = g.c_b2;
  p.alpha = g.alpha; p.act = g.act; p.accumulate = g.accumulate;
  p.c_f32 = g.c_dtype == S4_F32; p.atomic = splits > 1;
  CUtensorMap tcm = ta, txm = ta;
  const int n_x = (g.pre ? 1 : 0) + (g.res ? 1 : 0) + (g.aux ? 1 : 0);
  const bool need_add = g.accumulate || splits > 1;
  if (n_x <= 1 && !(need_add && !p.c_f32) && !env_no_tma_epilogue()) {
    const void* x = g.pre ? g.pre : (g.res ? g.res : g.aux);
    if ((rc = make_epilogue_maps(&tcm, &txm, g.c, x, p.c_f32, g.M, g.N, g.nb1, g.nb2, g.c_sm, g.c_b1, g.c_b2)))
      return rc;
    p.tma_epi = 1;
    p.x_mode = g.pre ? 1 : (g.res ? 2 : (g.aux ? 3 : 0));
    p.accumulate = need_add ? 1 : 0;
  }
  if (g.colsum) {
    if (!p.tma_epi || need_add || nb != 1) {
      s4_set_error("gemm_tc: the fused column sum needs the TMA-staged, non-accumulating, unbatched epilogue");
      return S4_ERR_ARG;
    }
    p.colsum = g.colsum;
  }
  S4ProfScope prof("gemm_tc", 2.0 * g.M * g.N * (double)g.K * nb, 0, stream);
  return launch_any(BN, CT, ta, tb, tcm, txm, p, stream);
}

// ------------------------------------------------------------------------------------------
// 3x3 convolution (NHWC bf16) as implicit GEMM
// ------------------------------------------------------------------------------------------
static bool conv_tile_shape(int H, int W, int& TW, int& TH) {
  if (W >= 128) {
    // widest divisor of W that is <= 128 and a multiple of 8
    TW = 0;
    for (int t = 128; t >= 8; t -= 8)
      if (W % t == 0) { TW = t; break; }
    TH = 1;
    return TW >= 64;
  }
  TW = W;
  TH = 128 / W;
  while (TH > 1 && H % TH) --TH;
  return TW % 8 == 0 && TH >= 1 && TW * TH >= 64;
}

bool s4_conv3x3_tc_supported(int B, int H, int W, int Cin, int Cout, int dtype) {
  if (dtype != S4_BF16) return false;
  if (Cin % 64 || Cout % 8) return false;
  int TW, TH;
  return conv_tile_shape(H, W, TW, TH);
}

int s4_conv3x3_tc_stats(const void* x, const void* w_packed, void* y, float* sum, float* sumsq, int B,
                        int H, int W, int Cin, int Cout, cudaStream_t stream);

int s4_conv3x3_tc(const void* x, const void* w_packed, void* y, int B, int H, int W, int Cin,
                  int Cout, cudaStream_t stream) {
  return s4_conv3x3_tc_stats(x, w_packed, y, nullptr, nullptr, B, H, W, Cin, Cout, stream);
}

// true when the conv forward can emit the BatchNorm statistics from its epilogue
bool s4_conv3x3_tc_stats_supported(int B, int H, int W, int Cin, int Cout, int dtype) {
  int TW, TH;
  return s4_conv3x3_tc_supported(B, H, W, Cin, Cout, dtype) && conv_tile_shape(H, W, TW, TH) &&
         (TW * TH) % 32 == 0 && !env_no_tma_epilogue();
}

int s4_conv3x3_tc_stats(const void* x, const void* w_packed, void* y, float* sum, float* sumsq, int B,
                        int H, int W, int Cin, int Cout, cudaStream_t stream) {
  int TW, TH;
  if (!conv_tile_shape(H, W, TW, TH)) {
    s4_set_error("conv3x3_tc: unsupported spatial shape %dx%d", H, W);
    return S4_ERR_UNSUPPORTED;
  }
  const int K = 9 * Cin;
  const int tiles_m = B * (H / TH) * (W / TW);
  // every CTA streams the same weight k-blocks in step (the L2 serves them once per wave), so the
  // single-CTA tile is not feed-bound here: pairs are not used
  const TileCfg tcfg = pick_cfg(tiles_m, Cout, 9 * (Cin / 64), 1, 1, false, false);
  const int BN = tcfg.BN;
  CUtensorMap ta, tb;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Cin, (uint64_t)W * Cin, (uint64_t)H * W * Cin};
    const uint32_t box[4] = {64, (uint32_t)TW, (uint32_t)TH, 1};
    if ((rc = s4_make_tmap_bf16(&ta, x, dims, str, box))) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)K, (uint64_t)Cout, 1, 1};
    const uint64_t str[3] = {(uint64_t)K, (uint64_t)K * 8, (uint64_t)K * 8};
    const uint32_t box[4] = {64, (uint32_t)BN, 1, 1};
    if ((rc = s4_make_tmap_bf16(&tb, w_packed, dims, str, box))) return rc;
  }
  TcParams p{};
  p.tiles_m = tiles_m;
  p.tiles_n = (Cout + BN - 1) / BN;
  p.nb = 1; p.nb2 = 1; p.splits = 1;
  // tile mt covers pixels [mt*rows_valid, (mt+1)*rows_valid): contiguous in NHWC because
  // either TW == W (whole rows) or TH == 1 (a row segment)
  p.rows_valid = TW * TH; p.row_pitch = TW * TH;
  p.M = B * H * W;
  p.N = Cout;
  p.kblocks = 9 * (Cin / 64);
  p.kb_per_split = p.kblocks;
  p.a_mode = OP_CONV_K; p.b_mode = OP_KMAJOR;
  p.cH = H; p.cW = W; p.cTW = TW; p.cTH = TH; p.cblocks = Cin / 64;
  p.c = y;
  p.c_sm = Cout; p.c_sn = 1; p.c_b1 = 0; p.c_b2 = 0;
  p.alpha = 1.f; p.c_f32 = 0;
  CUtensorMap tcm = ta, txm = ta;
  if (p.rows_valid % 32 == 0 && !env_no_tma_epilogue()) {
    if ((rc = make_epilogue_maps(&tcm, &txm, y, nullptr, false, p.M, Cout, 1, 1, Cout, 0, 0))) return rc;
    p.tma_epi = 1;
  }
  if (sum) {
    if (!p.tma_epi) {
      s4_set_error("conv3x3_tc: fused BN statistics need the TMA-staged epilogue");
      return S4_ERR_ARG;
    }
    p.colsum = sum;
    p.colsq = sumsq;
  }
  S4ProfScope prof("conv3x3_tc", 2.0 * B * H * W * (double)Cout * 9.0 * Cin, 0, stream);
  return launch_any(BN, 1, ta, tb, tcm, txm, p, stream);
}

// dw[co][ci][tap] += sum_pix dy[pix][co] * x[pix+tap][ci]   (split over pixels, fp32 atomics)
// pixel window of one wgrad k-block: TW x TH pixels, TW * TH <= 64.  64-pixel windows when W allows
// (W % 64 == 0, or whole rows with 64 % W == 0); otherwise the widest row segment <= 64 that
// divides W and is a multiple of 8 (W = 48, 96: 48 pixels; the k-block is zero-padded to 64).
static bool conv_wgrad_kblock(int H, int W, int* TW, int* TH) {
  int tw = 0, th = 1;
  if (W >= 64 && W % 64 == 0) tw = 64;
  else if (W < 64 && 64 % W == 0 && H % (64 / W) == 0 && W % 8 == 0) { tw = W; th = 64 / W; }
  else {
    for (int t = 64; t >= 16; t -= 8)
      if (W % t == 0) { tw = t; break; }
  }
  if (tw == 0) return false;
  if (TW) *TW = tw;
  if (TH) *TH = th;
  return true;
}

bool s4_conv3x3_wgrad_tc_supported(int B, int H, int W, int Cin, int Cout, int dtype) {
  if (dtype != S4_BF16 || env_tc_disable_mn()) return false;
  if (Cin % 8 || Cout % 8) return false;
  // a k-block is up to 64 consecutive pixels inside one image: a row segment, or whole rows
  return conv_wgrad_kblock(H, W, nullptr, nullptr);
}

int s4_conv3x3_wgrad_tc(const void* x, const void* dy, float* dw, int B, int H, int W, int Cin,
                        int Cout, cudaStream_t stream) {
  const long long P = (long long)B * H * W;
  int TWk = 64, THk = 1;
  if (!conv_wgrad_kblock(H, W, &TWk, &THk)) {
    s4_set_error("conv3x3_wgrad_tc: unsupported spatial shape %dx%d", H, W);
    return S4_ERR_UNSUPPORTED;
  }
  const int krows = TWk * THk;                     // pixels per k-block (64, or 48 / 40 / ... padded)
  const int kblocks = (int)(P / krows);
  const int BN = Cin >= 256 ? 256 : (Cin >= 128 ? 128 : 64);
  // CTA pairs: both 128-row halves of dy^T share the x window, each CTA stages half of it
  const int CT = (env_pair_mode() != 0 && BN >= 128 && Cout % 256 == 0) ? 2 : 1;
  CUtensorMap ta, tb;
  int rc;
  {
    // A = dy^T, MN-major: inner = Cout, outer = pixels
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)P, 1, 1};
    const uint64_t str[3] = {(uint64_t)Cout, (uint64_t)Cout * 8, (uint64_t)Cout * 8};
    const uint32_t box[4] = {64, (uint32_t)krows, 1, 1};
    if ((rc = s4_make_tmap_bf16(&ta, dy, dims, str, box))) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Cin, (uint64_t)W * Cin, (uint64_t)H * W * Cin};
    const uint32_t box[4] = {64, (uint32_t)TWk, (uint32_t)THk, 1};
    if ((rc = s4_make_tmap_bf16(&tb, x, dims, str, box))) return rc;
  }
  TcParams p{};
  p.tiles_m = ((Cout + BM - 1) / BM + CT - 1) / CT;
  p.tiles_n = (Cin + BN - 1) / BN;
  p.nb = 9; p.nb2 = 9;
  // work items = taps x tiles x splits: fill whole waves of CTA groups (never a ragged last wave)
  const int base_tiles = p.tiles_m * p.tiles_n * 9;
  const int groups = s4_num_sms() / CT;
  int splits = groups / base_tiles;
  if (splits < 1) splits = 1;
  if (kblocks / splits > 512) splits = (2 * groups) / base_tiles;    // long reductions: two waves
  if (splits > kblocks) splits = kblocks;
  if (splits < 1) splits = 1;
  int kb_per = (kblocks + splits - 1) / splits;
  splits = (kblocks + kb_per - 1) / kb_per;
  p.splits = splits;
  p.M = Cout; p.N = Cin;
  p.kblocks = kblocks; p.kb_per_split = kb_per;
  p.a_mode = OP_MNMAJOR; p.b_mode = OP_CONV_MN;
  p.cH = H; p.cW = W; p.cTW = TWk; p.cTH = THk; p.cblocks = 1;
  p.rows_valid = BM; p.row_pitch = BM; p.a_nobatch = 1;
  p.k_rows = krows;
  p.c = dw;
  p.c_sm = (long long)Cin * 9; p.c_sn = 9; p.c_b1 = 0; p.c_b2 = 1;   // z2 = tap
  p.alpha = 1.f; p.c_f32 = 1; p.atomic = 1; p.accumulate = 1;
  S4ProfScope prof("conv3x3_wgrad_tc", 2.0 * B * H * W * (double)Cout * 9.0 * Cin, 0, stream);
  return launch_any(BN, CT, ta, tb, ta, ta, p, stream);
}
