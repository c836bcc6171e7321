// dist_kernels.cu — the gradient all-reduce of the env-sharded trainer as ONE hand-written kernel over NVLink peer memory.
//
// Every rank keeps its flat gradient (and the 4-scalar KL / loss tail behind it) in a SYMMETRIC buffer: the same allocation on every GPU of the
// node, each mapped into every process (torch.distributed._symmetric_memory does the allocation and the handle exchange — plumbing).  The weight-
// gradient GEMMs write straight into that buffer; there is no staging copy.  The kernel below is a one-shot all-reduce:
//   1. publish "my gradient is complete" (a flag word in every peer's buffer; stream order already guarantees the producing kernels have finished),
//   2. wait until every peer has published,
//   3. out[i] = sum over ranks r = 0 .. W-1 of peer_r[i]   — 128-bit loads straight from the peers' HBM through NVSwitch, the SAME summation order
//      on every rank, so all ranks hold bit-identical sums (the KL-adaptive learning rate and the parameters stay identical without a broadcast),
//   4. publish "I am done reading" and wait for every peer's: after that the kernel exits and the next mini-batch may overwrite the gradient.
// 2 MB of gradient x 8 ranks is ~16 MB of peer reads per rank per optimiser step — latency, not bandwidth; the whole exchange is one launch that
// lives INSIDE the update's CUDA graph (no NCCL kernel in the graph, no host-launched collective between graph segments).
// With four or more ranks the same launch runs TWO-SHOT when the result buffer is symmetric as well: rank r sums only slice r of the vector (1/W of
// the peer reads) and stores the sums into EVERY rank's result buffer (peer stores are fire-and-forget); the "done" barrier of step 4 then also
// means "every slice of my result has landed".  Each element is still summed once, in rank order, so the ranks stay bit-identical.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/go2_b200.h"
#include "common.cuh"

namespace go2 {

constexpr int P2P_MAX_RANKS = 8;
struct P2PPeers {
  const float* data[P2P_MAX_RANKS];   // every rank's symmetric buffer (this process's mapping), index = rank
  uint32_t* flags[P2P_MAX_RANKS];     // every rank's flag block: ready[8] at +0, done[8] at +8 (uint32 words)
  float* out[P2P_MAX_RANKS];          // two-shot only: every rank's (symmetric) result buffer
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float* p, float4 v) {
  asm volatile("st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_peer(const float* p) {   // L2 only (peer lines are never valid in this SM's L1 across launches anyway)
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// ctr[0] = number of completed exchanges (the epoch both flag words compare against), ctr[1] = ticket counter of the running launch
template <bool TWO_SHOT>
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(const P2PPeers peers, float* __restrict__ out, long off, long n4_all, int rank, int W,
                                                            uint32_t* __restrict__ ctr) {
  __shared__ uint32_t s_epoch, s_last;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(ctr) + 1u;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  uint32_t* mine = peers.flags[rank];
  if (blockIdx.x == 0 && (int)threadIdx.x < W) {
    __threadfence_system();
    st_release_sys(peers.flags[threadIdx.x] + rank, epoch);                       // ready[rank] in peer threadIdx.x's block
  }
  if ((int)threadIdx.x < W) {
    while ((int32_t)(ld_acquire_sys(mine + threadIdx.x) - epoch) < 0) {}           // every peer's gradient is complete
  }
  __syncthreads();
  // two-shot: this rank owns vectors [lo, lo + n4) of the exchange; one-shot: all of them
  long lo = 0, n4 = n4_all;
  if (TWO_SHOT) {
    const long per = (n4_all + W - 1) / W;
    lo = per * rank;
    n4 = lo < n4_all ? (lo + per <= n4_all ? per : n4_all - lo) : 0;
    off += 4 * lo;
  }
  // four vectors per thread and trip: 4 x W independent peer loads in flight before the first add (a peer load is a ~2 us NVLink round trip)
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * stride) {
    float4 acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long i = i0 + u * stride;
      acc[u] = i < n4 ? ld_peer(peers.data[0] + off + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int r = 1; r < W; ++r) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long i = i0 + u * stride;
        v[u] = i < n4 ? ld_peer(peers.data[r] + off + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long i = i0 + u * stride;
      if (i >= n4) continue;
      if (TWO_SHOT) {
        for (int r = 0; r < W; ++r) st_peer(peers.out[r] + off + 4 * i, acc[u]);     // this slice of EVERY rank's result
      } else {
        *reinterpret_cast<float4*>(out + off + 4 * i) = acc[u];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (TWO_SHOT) __threadfence_system(); else
    __threadfence();
    s_last = (atomicAdd(ctr + 1, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  // the last block of this rank: all of this rank's peer reads have been issued and consumed (two-shot: and its slice has been stored everywhere —
  // every block fenced at system scope before taking its ticket)
  if ((int)threadIdx.x < W) {
    if (TWO_SHOT) __threadfence_system();
    st_release_sys(peers.flags[threadIdx.x] + P2P_MAX_RANKS + rank, epoch);       // done[rank] in peer threadIdx.x's block
    while ((int32_t)(ld_acquire_sys(mine + P2P_MAX_RANKS + threadIdx.x) - epoch) < 0) {}   // every peer is done reading MY gradient
  }
  __syncthreads();
  if (threadIdx.x == 0) { ctr[1] = 0u; __threadfence(); *reinterpret_cast<volatile uint32_t*>(ctr) = epoch; }
}

}  // namespace go2

using namespace go2;

extern "C" {

// out[off .. off + n) = sum over the W ranks of peer_r[off .. off + n).  peer_data / peer_flags: HOST arrays of W device pointers (this process's
// mappings of every rank's symmetric buffer and flag block, index = rank; peer_flags[r] holds 16 zero-initialised uint32 words).  off and n are
// multiples of 4 floats, every pointer 16-byte aligned.  ctr: 2 zero-initialised uint32 words in THIS rank's memory.  All W ranks must call with
// the same (off, n) in the same order.
int go2_allreduce_p2p(const float* const* peer_data, uint32_t* const* peer_flags, float* out, long off, long n, int rank, int world, uint32_t* ctr,
                      void* stream) {
  return go2_allreduce_p2p2(peer_data, peer_flags, nullptr, out, off, n, rank, world, ctr, stream);
}

// Same exchange; peer_out != NULL (HOST array of `world` device pointers to every rank's SYMMETRIC result buffer, peer_out[rank] == out) selects
// the two-shot schedule for world >= 4: each rank sums one slice and stores it into every rank's result.
int go2_allreduce_p2p2(const float* const* peer_data, uint32_t* const* peer_flags, float* const* peer_out, float* out, long off, long n, int rank,
                       int world, uint32_t* ctr, void* stream) {
  if (!peer_data || !peer_flags || !out || !ctr || world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world)
    return set_error(1, "go2_allreduce_p2p: 1..8 ranks, no null pointers");
  if ((off & 3) || (n & 3) || n <= 0) return set_error(1, "go2_allreduce_p2p: off and n must be positive multiples of 4 floats");
  P2PPeers P{};
  const bool two_shot = peer_out != nullptr && world >= 4;
  for (int r = 0; r < world; ++r) {
    if (!peer_data[r] || !peer_flags[r] || ((uintptr_t)peer_data[r] & 15)) return set_error(1, "go2_allreduce_p2p: null / misaligned peer pointer");
    P.data[r] = peer_data[r]; P.flags[r] = peer_flags[r];
    if (two_shot) {
      if (!peer_out[r] || ((uintptr_t)peer_out[r] & 15)) return set_error(1, "go2_allreduce_p2p2: null / misaligned result pointer");
      P.out[r] = peer_out[r];
    }
  }
  if (two_shot && peer_out[rank] != out) return set_error(1, "go2_allreduce_p2p2: peer_out[rank] must be this rank's result buffer");
  const long n4 = n / 4;
  const long mine = two_shot ? (n4 + world - 1) / world : n4;
  // at most one co-resident wave: every block spins on the peers' flags
  const int blocks = (int)((mine + 255) / 256 < 128 ? (mine + 255) / 256 : 128);
  if (two_shot) p2p_allreduce_kernel<true><<<blocks < 1 ? 1 : blocks, 256, 0, (cudaStream_t)stream>>>(P, out, off, n4, rank, world, ctr);
  else p2p_allreduce_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(P, out, off, n4, rank, world, ctr);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
