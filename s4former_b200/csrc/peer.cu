// SyncBatchNorm's cross-rank statistics sum over NVLink peer memory (the
// reference gets it from torch.nn.SyncBatchNorm's all_gather of (mean, invstd, count),
// mmseg configs norm_cfg = dict(type='SyncBN')).
//
// The 40 reductions of a step carry 2 KB each ([2, C] fp32 sums): through NCCL each costs ~16 us of
// launch + protocol latency on the compute stream.  Here ONE small kernel per reduction does a
// one-shot all-reduce over symmetric (peer-mapped) buffers:
//   1. every rank writes its vector into ITS OWN buffer (double-buffered by call parity),
//   2. stores the call's sequence number into flag[rank] of EVERY peer (st.release.sys over NVLink),
//   3. waits until all its own flags have reached the sequence number (ld.acquire.sys, local),
//   4. loads the vectors of all ranks (peer loads) and adds them in rank order, so every rank
//      gets bit-identical sums.
// A rank cannot start call k+2 (which reuses call k's buffer) before every rank has signalled
// k+1, i.e. has finished reading call k: two buffers are enough.  The sequence number lives in
// device memory, so the kernel replays unchanged inside a CUDA graph.
#include "common.cuh"

namespace {

constexpr int PEER_NMAX = 2048;        // floats per call
constexpr int PEER_MAXW = 64;          // ranks
constexpr size_t PEER_DATA_BYTES = 2 * (size_t)PEER_NMAX * sizeof(float);

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// steps 1-3 for the calling block: publish `data`, signal, wait.  Returns the parity of the call.
__device__ __forceinline__ unsigned peer_exchange(const float* __restrict__ data, int n,
                                                  const unsigned long long* sp, int rank, int world,
                                                  unsigned seq) {
  const int tid = threadIdx.x;
  const unsigned par = seq & 1u;
  float* mine = reinterpret_cast<float*>(sp[rank]) + par * PEER_NMAX;
  for (int i = tid; i < n; i += blockDim.x) mine[i] = data[i];
  __threadfence_system();
  __syncthreads();
  if (tid < world) {
    unsigned* theirs = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(sp[tid]) + PEER_DATA_BYTES) + rank;
    st_release_sys(theirs, seq);
    const unsigned* flag = reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(sp[rank]) + PEER_DATA_BYTES) + tid;
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(flag) - seq) < 0) {
      if ((++spins & 0x3FFu) == 0) {       // a rank that never arrives: trap instead of hanging the GPU
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20000000000ull) __trap();
      }
    }
  }
  __syncthreads();
  return par;
}

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(float* __restrict__ data, int n, const unsigned long long* __restrict__ peers, int rank,
                      int world, unsigned* __restrict__ seq_state) {
  __shared__ unsigned long long sp[PEER_MAXW];
  const int tid = threadIdx.x;
  if (tid < world) sp[tid] = peers[tid];
  const unsigned seq = *seq_state + 1u;
  __syncthreads();
  const unsigned par = peer_exchange(data, n, sp, rank, world, seq);
  for (int i = tid; i < n; i += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < world; ++r)
      acc += ld_relaxed_sys(reinterpret_cast<const float*>(sp[r]) + par * PEER_NMAX + i);
    data[i] = acc;
  }
  if (tid == 0) *seq_state = seq;
}

// SyncBN forward in ONE kernel: exchange the local (sum, sumsq) [2, C] over peer memory, add them in
// rank order, and finalize (mean, invstd, scale = gamma invstd, shift, running statistics) exactly as
// bn_finalize_kernel (head.cu) does from NCCL-reduced sums.
__global__ void __launch_bounds__(256)
peer_bn_finalize_kernel(const float* __restrict__ stats, double count, float eps, float momentum,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ scale,
                        float* __restrict__ shift, float* __restrict__ rmean, float* __restrict__ rvar,
                        long long* __restrict__ nbt, int C, const unsigned long long* __restrict__ peers,
                        int rank, int world, unsigned* __restrict__ seq_state) {
  __shared__ unsigned long long sp[PEER_MAXW];
  const int tid = threadIdx.x;
  if (tid < world) sp[tid] = peers[tid];
  const unsigned seq = *seq_state + 1u;
  __syncthreads();
  const unsigned par = peer_exchange(stats, 2 * C, sp, rank, world, seq);
  for (int c = tid; c < C; c += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int r = 0; r < world; ++r) {
      const float* pr = reinterpret_cast<const float*>(sp[r]) + par * PEER_NMAX;
      s1 += ld_relaxed_sys(pr + c);
      s2 += ld_relaxed_sys(pr + C + c);
    }
    const double mu = (double)s1 / count;
    double var = (double)s2 / count - mu * mu;
    if (var < 0) var = 0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    mean[c] = (float)mu;
    invstd[c] = is;
    const float sc = gamma[c] * is;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mu * sc;
    if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mu;
    if (rvar) {
      const double unb = count > 1 ? var * count / (count - 1.0) : var;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
    }
  }
  if (tid == 0) {
    *seq_state = seq;
    if (nbt) *nbt += 1;
  }
}

}  // namespace

extern "C" long long s4_peer_allreduce_buffer_bytes(void) {
  return (long long)(PEER_DATA_BYTES + PEER_MAXW * sizeof(unsigned));
}

extern "C" int s4_peer_allreduce_max_elems(void) { return PEER_NMAX; }

extern "C" int s4_peer_allreduce_f32(float* data, int n, const void* peer_bufs_dev, int rank, int world,
                                     unsigned* seq_state, cudaStream_t stream) {
  S4ProfScope prof_("peer_allreduce", 0.0, 1, stream);
  S4_REQUIRE(n >= 0 && n <= PEER_NMAX, "peer_allreduce: n=%d not in [0,%d]", n, PEER_NMAX);
  S4_REQUIRE(world >= 1 && world <= PEER_MAXW && rank >= 0 && rank < world, "peer_allreduce: rank %d / world %d",
             rank, world);
  S4_REQUIRE(data && peer_bufs_dev && seq_state, "peer_allreduce: null pointer");
  peer_allreduce_kernel<<<1, 256, 0, stream>>>(data, n, (const unsigned long long*)peer_bufs_dev, rank, world,
                                               seq_state);
  return s4_check_launch("peer_allreduce");
}

extern "C" int s4_bn_finalize_peer(const float* stats, double count, float eps, float momentum,
                                   const float* gamma, const float* beta, float* mean, float* invstd,
                                   float* scale, float* shift, float* running_mean, float* running_var,
                                   long long* num_batches_tracked, int C, const void* peer_bufs_dev, int rank,
                                   int world, unsigned* seq_state, cudaStream_t stream) {
  S4ProfScope prof_("bn_finalize_peer", 0.0, 1, stream);
  S4_REQUIRE(C >= 1 && 2 * C <= PEER_NMAX, "bn_finalize_peer: C=%d not in [1,%d]", C, PEER_NMAX / 2);
  S4_REQUIRE(world >= 1 && world <= PEER_MAXW && rank >= 0 && rank < world, "bn_finalize_peer: rank %d / world %d",
             rank, world);
  S4_REQUIRE(stats && peer_bufs_dev && seq_state, "bn_finalize_peer: null pointer");
  peer_bn_finalize_kernel<<<1, 256, 0, stream>>>(stats, count, eps, momentum, gamma, beta, mean, invstd, scale,
                                                 shift, running_mean, running_var, num_batches_tracked, C,
                                                 (const unsigned long long*)peer_bufs_dev, rank, world, seq_state);
  return s4_check_launch("bn_finalize_peer");
}
