// gemm_tc.cu — the trainer's dense contraction on Blackwell tensor cores: C[M,N] = A[M,K] · B[N,K]^T, fp32 in HBM,
// tf32 multiply / fp32 accumulate in TMEM (tcgen05.mma.kind::tf32), operands streamed by TMA (cp.async.bulk.tensor, 128B
// swizzle) through an mbarrier ring, fused epilogues for the three uses of the MLP path:
//   forward   Y  = act(X W^T + b)                     (A = X  [rows, in],  B = W   [out, in])        actor_critic.py:58-79
//   dgrad     dX = (dZ W) * ELU'(Y_prev)              (A = dZ [rows, out], B = W^T [in, out])        autograd of the above
//   wgrad     dW = dZ^T X  (split over the rows)      (A = dZ^T [out, rows], B = X^T [in, rows])
// Both operands are always K-major (contraction index contiguous): the producers of activations / gradients emit the
// transposed copy next to the row-major one (epilogue flag), which costs one extra coalesced store instead of an MN-major
// tf32 descriptor (tf32 only allows the special 128B_BASE32B layout there).
//
// CTA = 6 warps: warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM -> registers -> global; warp w owns TMEM lanes 32*(w%4)..+31).  One 128 x BN tile per CTA.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "../../include/go2_b200.h"
#include "common.cuh"

namespace go2 {

constexpr int TC_BM = 128;        // UMMA M (cta_group::1)
constexpr int TC_BK = 32;         // floats per stage along K = one 128-byte swizzle atom
constexpr int TC_UMMA_K = 8;      // tf32: 32 bytes of K per instruction
constexpr int TC_THREADS = 192;

enum { TC_EPI_PLAIN = 0, TC_EPI_BIAS = 1, TC_EPI_BIAS_ELU = 2, TC_EPI_MUL_ELU_GRAD = 3 };

struct TcParams {
  int M, N, K;             // logical sizes (rows of A, rows of B, contraction)
  int kb_per_split;        // K blocks (of TC_BK) each grid.z slice handles
  float* C; long ldc;      // row-major output (may be null)
  float* Ct; long ldct;    // transposed output Ct[n][m] (may be null)
  const float* bias;       // [N] or null
  const float* aux; long ldaux;  // ELU output of the previous layer (dgrad), row-major, or null
  const float* aux_t; long ldaux_t;  // the same activation transposed [N, M] (preferred: coalesced in the epilogue)
  int epi;
  long split_stride;       // elements between consecutive split slices of C
  long long* dbg;          // optional [gridDim.x*gridDim.y][8] clock64 stamps (profiling aid)
  int mn_major;            // persistent kernel only: A [K, M] and B [K, N] row-major (contraction index = rows)
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
               "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, "
      "%29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows at 128-byte pitch, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

// MN-major tf32 operand tile (the only layout tcgen05 accepts for it: 128B swizzle with 32-byte atoms, written by TMA mode
// SWIZZLE_128B_ATOM_32B): boxes of [32 contraction rows][32 MN elements = 128 B]; 4-row groups 512 B apart (stride offset),
// 32-element MN groups one 4 KB box apart (leading offset)
__device__ __forceinline__ uint64_t make_desc_mnmajor_sw128_32b(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(4096 >> 4) << 16;                  // leading byte offset: next group of 32 MN elements
  d |= (uint64_t)(512 >> 4) << 32;                   // stride byte offset: next group of 4 contraction rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                            // SWIZZLE_128B_BASE32B
  return d;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, keeps the shared address space
  constexpr int A_BYTES = TC_BM * TC_BK * 4, B_BYTES = BN * TC_BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
  long long* dbg = p.dbg ? p.dbg + (long)(blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  const int total_kb = (p.K + TC_BK - 1) / TC_BK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int nkb = max(0, min(p.kb_per_split, total_kb - kb0));

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation: BN fp32 accumulator columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(BN < 32 ? 32 : BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == 0) {
    // ===== TMA producer
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty_bar + s, ph ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        mbar_expect_tx(full_bar + s, STAGE_BYTES);
        tma_load_2d(&tmA, full_bar + s, sa, (kb0 + kb) * TC_BK, m0);
        tma_load_2d(&tmB, full_bar + s, sa + A_BYTES, (kb0 + kb) * TC_BK, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(full_bar + s, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t da = make_desc_kmajor_sw128(sa), db = make_desc_kmajor_sw128(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
          // advance 32 bytes of K inside the swizzle atom: +2 in the (address >> 4) field
          umma_tf32(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (kb | k) ? 1u : 0u);
        }
        umma_commit(empty_bar + s);                       // frees the smem slot when these MMAs retire
        if (kb == nkb - 1) umma_commit(tmem_full);        // accumulator complete
      }
      __syncwarp();
    }
    if (nkb == 0 && elect_one()) {                        // empty K range: nothing to wait for
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(tmem_full)) : "memory");
    }
  } else {
    // ===== epilogue warps: quadrant q of the 128 TMEM lanes.  Thread = one output row; the transposed copy is stored straight
    // from registers (lanes = consecutive rows -> coalesced), the row-major copy goes through a padded 32x32 shared tile so
    // that every store instruction writes one 128-byte row segment.  ELU' (dgrad) reads the TRANSPOSED activation, also coalesced.
    const int q = warp & 3;
    float* tile = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256) + (warp - 2) * (32 * 36);
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (dbg && threadIdx.x == 64) dbg[2] = clock64();
    const int mrow0 = m0 + 32 * q;
    const int m = mrow0 + lane;
    float* Cz = p.C ? p.C + (long)blockIdx.z * p.split_stride : nullptr;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int nb = n0 + c * 32;
      if (nb >= p.N) break;
      uint32_t r[32];
      if (nkb > 0) tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(c * 32), r);
      else { for (int j = 0; j < 32; ++j) r[j] = 0; }
      float v[32];
      float bias_lane = 0.0f;   // one coalesced load per chunk, broadcast by shuffle
      if ((p.epi == TC_EPI_BIAS || p.epi == TC_EPI_BIAS_ELU) && nb + lane < p.N) bias_lane = __ldg(p.bias + nb + lane);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(r[j]);
        if (p.epi == TC_EPI_BIAS || p.epi == TC_EPI_BIAS_ELU) x += __shfl_sync(0xffffffffu, bias_lane, j);
        if (p.epi == TC_EPI_BIAS_ELU) x = x > 0.0f ? x : (__expf(x) - 1.0f);   // |abs err| < 2e-7: below tf32 resolution of the product
        v[j] = x;
      }
      if (p.epi == TC_EPI_MUL_ELU_GRAD && (m < p.M || (p.aux_t && mrow0 + 31 < p.M))) {
        if (p.aux_t && (p.ldaux_t & 3) == 0 && mrow0 + 31 < p.M) {
          // 128-bit loads of the transposed activation (4 columns x 128 B per instruction), redistributed through the shared tile
          const int mc = (lane & 7) * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int jj = it * 4 + (lane >> 3), n = nb + jj;
            float4 y = make_float4(1.f, 1.f, 1.f, 1.f);
            if (n < p.N) y = __ldg(reinterpret_cast<const float4*>(p.aux_t + (long)n * p.ldaux_t + mrow0 + mc));
            *reinterpret_cast<float4*>(tile + jj * 36 + mc) = y;
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; ++j) { const float y = tile[j * 36 + lane]; v[j] *= (y > 0.0f ? 1.0f : y + 1.0f); }
          __syncwarp();
        } else if (p.aux_t) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { const int n = nb + j; if (n < p.N) { const float y = __ldg(p.aux_t + (long)n * p.ldaux_t + m); v[j] *= (y > 0.0f ? 1.0f : y + 1.0f); } }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) { const int n = nb + j; if (n < p.N) { const float y = p.aux[(long)m * p.ldaux + n]; v[j] *= (y > 0.0f ? 1.0f : y + 1.0f); } }
        }
      }
      // ---- stores: 128-bit vectors, 4 rows x 128 B per warp instruction, staged through a padded shared tile (pitch 36 floats)
      constexpr int TP = 36;
      const bool vec_ok = ((p.N & 3) == 0);
      if (Cz) {
        if (vec_ok && (p.ldc & 3) == 0) {
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) *reinterpret_cast<float4*>(tile + lane * TP + 4 * q4) = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
          __syncwarp();
          const int cc = (lane & 7) * 4, n = nb + cc;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3), mm = mrow0 + rr;
            if (mm < p.M && n < p.N) *reinterpret_cast<float4*>(Cz + (long)mm * p.ldc + n) = *reinterpret_cast<const float4*>(tile + rr * TP + cc);
          }
          __syncwarp();
        } else if (m < p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { const int n = nb + j; if (n < p.N) Cz[(long)m * p.ldc + n] = v[j]; }
        }
      }
      if (p.Ct) {
        if ((p.ldct & 3) == 0 && mrow0 + 31 < p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) tile[j * TP + lane] = v[j];
          __syncwarp();
          const int mc = (lane & 7) * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int jj = it * 4 + (lane >> 3), n = nb + jj;
            if (n < p.N) *reinterpret_cast<float4*>(p.Ct + (long)n * p.ldct + mrow0 + mc) = *reinterpret_cast<const float4*>(tile + jj * TP + mc);
          }
          __syncwarp();
        } else if (m < p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { const int n = nb + j; if (n < p.N) p.Ct[(long)n * p.ldct + m] = v[j]; }
        }
      }
    }
    if (dbg && threadIdx.x == 64) dbg[3] = clock64();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[4] = clock64();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(BN < 32 ? 32 : BN)) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ persistent variant
// One CTA per SM walks a static list of (split, m-tile, n-tile) units.  The accumulator is double-buffered in TMEM (columns 0 /
// 256), so the MMA warp starts the next unit while the epilogue drains the previous one; the operand ring keeps running across
// units.  The epilogue (two groups of four warps, alternating 32-column chunks) never issues a global store itself: each
// 128 x 32 chunk is staged in shared memory (128B-swizzled for the row-major copy, dense [32][128] for the transposed copy) and
// leaves through TMA bulk-tensor stores, which clip the M / N tails; the ELU' operand of dgrad arrives the same way (TMA load
// into the row-major staging buffer, consumed in place).
struct TcParamsP {
  int M, N, K;
  int m_tiles, n_tiles, splits, kb_per_split, total_kb;
  int rows_pad;            // rows between consecutive split slices in the C map (multiple of 128)
  int stages;              // operand ring depth
  int mn_major;            // wgrad from row-major operands: A = dZ[batch, out], B = X[batch, in] are MN-major (contraction index = rows)
  long long* dbg;          // optional [gridDim.x][16] cycle counters (go2_gemm_set_debug): where each role of a CTA waited
  int split_rewrite;       // 3xTF32 split: 0 = lo only (hardware truncation is the hi part), 1 = raw stage rewritten with rn_tf32(a) (A/B check)
  const float* bias;
  int epi, has_c, has_ct, has_aux;
};

// ---- 3xTF32 (error-compensated split): every fp32 operand a is used as hi = rn_tf32(a) and lo = a - hi (exact in fp32), and the product is
// accumulated as lo_a hi_b + hi_a lo_b + hi_a hi_b (the lo lo term, 2^-22 relative, is dropped; hardware truncation of lo adds 2^-21): fp32-class
// products out of the tf32 tensor-core path, 3 MMAs per K step from the SAME TMA-loaded tiles.  Four "splitter" warps sit between the TMA ring and
// the MMA warp: they rewrite each landed stage in place with hi and write lo into a 2-slot side ring of the same (swizzled) layout, so the
// descriptors of the lo operands are the raw ones at another base address.  The GEMMs of this path are HBM / L2 bound (33 flop/B), so the two extra
// MMAs and the shared-memory pass hide behind the operand stream.
// Default split (p.split_rewrite == 0): the raw stage is left alone — tcgen05 kind::tf32 reads fp32 words and drops the 13 low mantissa bits, so
// the hardware's hi is trunc(a) and the splitters only write lo = rn_tf32(a - trunc(a)); the hi hi MMAs of a stage are issued as soon as its TMA
// bytes land, the two cross terms when the lo slot is ready.  EIGHT splitter warps, every thread loads its 8-9 vectors of the stage before the
// first store (round 2's first version serialised LDS -> cvt -> STS per vector in four warps: 3 us per 32 KB stage, the whole kernel's pace).
// lo word of one fp32 operand.  REWRITE = 0: the tensor core reads the raw word and ignores its 13 low mantissa bits, i.e. it multiplies
// hi = trunc_tf32(a); lo = rn_tf32(a - hi) (a - hi is exact).  REWRITE = 1: hi = rn_tf32(a) replaces the raw word, lo = a - hi.
template <int REWRITE>
__device__ __forceinline__ uint32_t split_word(uint32_t& u) {
  const float a = __uint_as_float(u);
  uint32_t h, l;
  if (REWRITE) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
    const float hf = __uint_as_float(h);
    l = __float_as_uint((fabsf(hf) < INFINITY) ? a - hf : 0.0f);
    u = h;
  } else {
    h = u & 0xFFFFE000u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(a - __uint_as_float(h)));
  }
  return l;
}
template <int REWRITE>
__device__ __forceinline__ uint4 split_tf32(uint4& v) {
  uint4 lo;
  lo.x = split_word<REWRITE>(v.x); lo.y = split_word<REWRITE>(v.y); lo.z = split_word<REWRITE>(v.z); lo.w = split_word<REWRITE>(v.w);
  return lo;
}
constexpr int TCP_LO_SLOTS = 2;
// mbarrier wait that adds the cycles it blocked to *acc when profiling (acc == nullptr otherwise)
__device__ __forceinline__ long long globaltimer_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, long long* acc) {
  if (acc) { const long long t0 = clock64(); mbar_wait(bar, parity); *acc += clock64() - t0; }
  else mbar_wait(bar, parity);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map), "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}

constexpr int TCP_THREADS = 320;                 // producer warp, MMA warp, 2 x 4 epilogue warps
constexpr int TCP_SPLIT_WARPS = 8;
constexpr int TCP_THREADS_X3 = 320 + 32 * TCP_SPLIT_WARPS;   // + splitter warps (3xTF32)
constexpr int TCP_CHUNK_BYTES = TC_BM * 32 * 4;  // one staged 128 x 32 chunk
constexpr int TCP_MAX_STAGES = 4;
constexpr int TCP_SMEM_MAX = 232448;             // 227 KB opt-in limit per CTA
static int tcp_smem_bytes(int BN, int stages, bool has_ct, bool x3) {
  return (stages + (x3 ? TCP_LO_SLOTS : 0)) * (TC_BM * TC_BK * 4 + BN * TC_BK * 4) + (has_ct ? 6 : 4) * TCP_CHUNK_BYTES + 256 + 1024;
}
// ring depth: 4 stages for the single-pass kernel; with the lo side ring what the 227 KB leave (3 at BN = 128, 2 at BN = 160 or with Ct)
static int tcp_stages(int BN, bool has_ct, bool x3) {
  int s = TCP_MAX_STAGES;
  while (s > 2 && tcp_smem_bytes(BN, s, has_ct, x3) > TCP_SMEM_MAX) --s;
  return s;
}

template <int BN, bool X3>
__global__ void __launch_bounds__(X3 ? TCP_THREADS_X3 : TCP_THREADS, 1)
gemm_tf32_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                         const __grid_constant__ CUtensorMap tmCt, const __grid_constant__ CUtensorMap tmAux, const TcParamsP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = TC_BM * TC_BK * 4, B_BYTES = BN * TC_BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int NCH = (BN + 31) / 32;
  const int S = p.stages;
  uint8_t* lo_ring = smem + S * STAGE_BYTES;                   // 3xTF32: TCP_LO_SLOTS stages holding the lo parts (same layout as the raw stage)
  uint8_t* cst = lo_ring + (X3 ? TCP_LO_SLOTS * STAGE_BYTES : 0);   // [group][2] row-major staging chunks [128 rows][128 B], SW128
  uint8_t* tst = cst + 4 * TCP_CHUNK_BYTES;                    // [group] transposed staging chunk [32 n][128 m] floats (only with Ct)
  uint64_t* full_bar = (uint64_t*)(tst + (p.has_ct ? 2 * TCP_CHUNK_BYTES : 0));
  uint64_t* empty_bar = full_bar + TCP_MAX_STAGES;
  uint64_t* tfull = empty_bar + TCP_MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* auxb = tempty + 2;                                 // [group][2]
  uint64_t* lofull = auxb + 4;                                 // [TCP_LO_SLOTS] splitter -> MMA (4 warps arrive)
  uint64_t* loempty = lofull + TCP_LO_SLOTS;                   // [TCP_LO_SLOTS] MMA commit -> splitter
  uint32_t* tmem_slot = (uint32_t*)(loempty + TCP_LO_SLOTS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = p.m_tiles * p.n_tiles * p.splits;
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 24 + 16] = globaltimer_ns();
  // programmatic dependent launch: the next GEMM of the stream may take over this SM as soon as this CTA exits and run its prologue (barriers, TMEM,
  // tensor-map prefetch) under the tail of this grid; it blocks in griddepcontrol.wait below until this grid has completed and flushed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    if (p.has_c) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmC) : "memory");
    if (p.has_ct) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmCt) : "memory");
    if (p.has_aux) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmAux) : "memory");
    for (int s = 0; s < S; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, 8); }
    for (int b = 0; b < 4; ++b) mbar_init(auxb + b, 1);
    for (int b = 0; b < TCP_LO_SLOTS; ++b) { mbar_init(lofull + b, TCP_SPLIT_WARPS); mbar_init(loempty + b, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // the whole TMEM: two accumulators of up to 256 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 24 + 17] = globaltimer_ns();
  asm volatile("griddepcontrol.wait;" ::: "memory");          // everything before this line touched no global memory (no-op without the launch attribute)
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 24 + 18] = globaltimer_ns();

  if (warp == 0) {
    // ===== TMA producer: the stage ring runs straight through the unit list
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      long long w_empty = 0, *pw = p.dbg ? &w_empty : nullptr;
      const long long t_start = p.dbg ? clock64() : 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int nt = u % p.n_tiles, t = u / p.n_tiles, mt = t % p.m_tiles, z = t / p.m_tiles;
        const int kb0 = z * p.kb_per_split, nkb = min(p.kb_per_split, p.total_kb - kb0);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_t(empty_bar + s, ph ^ 1, pw);
          uint8_t* sa = smem + s * STAGE_BYTES;
          mbar_expect_tx(full_bar + s, STAGE_BYTES);
          if (!p.mn_major) {
            tma_load_2d(&tmA, full_bar + s, sa, (kb0 + kb) * TC_BK, mt * TC_BM);
            tma_load_2d(&tmB, full_bar + s, sa + A_BYTES, (kb0 + kb) * TC_BK, nt * BN);
          } else {   // one [32 rows][32 features] box per 32-wide feature group
#pragma unroll
            for (int g = 0; g < TC_BM / 32; ++g) tma_load_2d(&tmA, full_bar + s, sa + g * 4096, mt * TC_BM + 32 * g, (kb0 + kb) * TC_BK);
#pragma unroll
            for (int g = 0; g < BN / 32; ++g) tma_load_2d(&tmB, full_bar + s, sa + A_BYTES + g * 4096, nt * BN + 32 * g, (kb0 + kb) * TC_BK);
          }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
      if (p.dbg) { long long* d = p.dbg + blockIdx.x * 24; d[0] = clock64() - t_start; d[1] = w_empty; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread)
    if (elect_one()) {
      constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int s = 0, ui = 0, l = 0;
      uint32_t ph = 0, lph = 0;
      long long w_tempty = 0, w_full = 0, w_lofull = 0, nst = 0;
      long long *pw_tempty = p.dbg ? &w_tempty : nullptr, *pw_full = p.dbg ? &w_full : nullptr, *pw_lofull = p.dbg ? &w_lofull : nullptr;
      const long long t_start = p.dbg ? clock64() : 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
        const int z = u / (p.n_tiles * p.m_tiles);
        const int kb0 = z * p.kb_per_split, nkb = min(p.kb_per_split, p.total_kb - kb0);
        const int buf = ui & 1;
        nst += nkb;
        mbar_wait_t(tempty + buf, ((ui >> 1) & 1) ^ 1, pw_tempty);             // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + (uint32_t)(buf * 256);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_t(full_bar + s, ph, pw_full);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sl = smem_u32(lo_ring + l * STAGE_BYTES);
          const bool kmaj = !p.mn_major;
          // K-major: +32 bytes of K inside the swizzle atom per instruction; MN-major: 8 contraction rows = 1024 B further into every box
          const uint64_t da = kmaj ? make_desc_kmajor_sw128(sa) : make_desc_mnmajor_sw128_32b(sa);
          const uint64_t db = kmaj ? make_desc_kmajor_sw128(sa + A_BYTES) : make_desc_mnmajor_sw128_32b(sa + A_BYTES);
          const uint64_t kstep = kmaj ? 2 : 64;
          const uint32_t idesc = kmaj ? IDESC : (IDESC | (1u << 15) | (1u << 16));
          const bool hi_first = !X3 || !p.split_rewrite;      // the raw stage is (or stays) the hi operand: no need to wait for the splitters
          if (hi_first) {
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) umma_tf32(tacc, da + kstep * k, db + kstep * k, idesc, (kb | k) ? 1u : 0u);
          }
          if (X3) {
            mbar_wait_t(lofull + l, lph, pw_lofull);              // the splitter warps have filled the lo slot (and, rewrite mode, rewritten the stage as hi)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t la = kmaj ? make_desc_kmajor_sw128(sl) : make_desc_mnmajor_sw128_32b(sl);
            const uint64_t lb = kmaj ? make_desc_kmajor_sw128(sl + A_BYTES) : make_desc_mnmajor_sw128_32b(sl + A_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
              umma_tf32(tacc, la + kstep * k, db + kstep * k, idesc, (hi_first || (kb | k)) ? 1u : 0u);
              umma_tf32(tacc, da + kstep * k, lb + kstep * k, idesc, 1u);
              if (!hi_first) umma_tf32(tacc, da + kstep * k, db + kstep * k, idesc, 1u);
            }
          }
          umma_commit(empty_bar + s);
          if (X3) { umma_commit(loempty + l); if (++l == TCP_LO_SLOTS) { l = 0; lph ^= 1; } }
          if (++s == S) { s = 0; ph ^= 1; }
        }
        umma_commit(tfull + buf);
      }
      if (p.dbg) { long long* d = p.dbg + blockIdx.x * 24; d[2] = clock64() - t_start; d[3] = w_tempty; d[4] = w_full; d[5] = w_lofull; d[6] = nst; }
    }
  } else if (X3 && warp >= 10) {
    // ===== splitter warps (3xTF32): stage s landed -> hi in place, lo into slot l; then hand both to the MMA warp
    const int t = threadIdx.x - 320;
    constexpr int NV = STAGE_BYTES / 16, NT = 32 * TCP_SPLIT_WARPS, PER = (NV + NT - 1) / NT;
    int s = 0, l = 0;
    uint32_t ph = 0, lph = 0;
    long long w_full = 0, w_loempty = 0, t_work = 0;
    long long *pw_full = p.dbg ? &w_full : nullptr, *pw_loempty = p.dbg ? &w_loempty : nullptr;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int z = u / (p.n_tiles * p.m_tiles);
      const int kb0 = z * p.kb_per_split, nkb = min(p.kb_per_split, p.total_kb - kb0);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_t(full_bar + s, ph, pw_full);                   // TMA bytes have landed (async proxy -> visible after the wait)
        const long long t_w0 = p.dbg ? clock64() : 0;
        uint4* raw = reinterpret_cast<uint4*>(smem + s * STAGE_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(lo_ring + l * STAGE_BYTES);
        uint4 v[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) if (NV % NT == 0 || t + NT * j < NV) v[j] = raw[t + NT * j];
        mbar_wait_t(loempty + l, lph ^ 1, pw_loempty);            // the MMAs that read this lo slot have retired
        if (p.split_rewrite) {
#pragma unroll
          for (int j = 0; j < PER; ++j) if (NV % NT == 0 || t + NT * j < NV) { const uint4 w = split_tf32<1>(v[j]); raw[t + NT * j] = v[j]; lo[t + NT * j] = w; }
        } else {
#pragma unroll
          for (int j = 0; j < PER; ++j) if (NV % NT == 0 || t + NT * j < NV) lo[t + NT * j] = split_tf32<0>(v[j]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(lofull + l)) : "memory");
        if (p.dbg) t_work += clock64() - t_w0;
        if (++l == TCP_LO_SLOTS) { l = 0; lph ^= 1; }
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
    if (p.dbg && t == 0) { long long* d = p.dbg + blockIdx.x * 24; d[7] = w_full; d[8] = w_loempty; d[9] = t_work; }   // t_work includes w_loempty
  } else {
    // ===== epilogue: group grp takes chunks grp, grp+2, ... of every unit; warp -> TMEM lane quadrant q; thread = one tile row
    const int grp = (warp - 2) >> 2, q = warp & 3;
    const int row = 32 * q + lane;
    const bool leader = (((warp - 2) & 3) == 0) && lane == 0;
    const int sw = row & 7;
    uint8_t* cst_g = cst + grp * 2 * TCP_CHUNK_BYTES;
    float* ts = reinterpret_cast<float*>(tst + grp * TCP_CHUNK_BYTES);
    uint64_t* auxb_g = auxb + grp * 2;
    auto n_chunks = [&](int u) { return min(NCH, (p.N - (u % p.n_tiles) * BN + 31) / 32); };
    // the chunk this group handles after (u, c) — the leader uses it to prefetch the ELU' operand one chunk ahead
    auto advance = [&](int& u, int& c) {
      c += 2;
      while (u < units && c >= n_chunks(u)) { u += gridDim.x; c = grp; }
      return u < units;
    };
    auto aux_load = [&](int u, int c, int slot) {
      const int nt = u % p.n_tiles, mt = (u / p.n_tiles) % p.m_tiles;
      mbar_expect_tx(auxb_g + slot, TCP_CHUNK_BYTES);
      tma_load_2d(&tmAux, auxb_g + slot, cst_g + slot * TCP_CHUNK_BYTES, nt * BN + c * 32, mt * TC_BM);
    };
    long long w_tfull = 0, *pw_tfull = (p.dbg && leader) ? &w_tfull : nullptr;
    long long w_aux = 0, *pw_aux = (p.dbg && leader) ? &w_aux : nullptr;
    const long long t_start = p.dbg ? clock64() : 0;
    uint32_t g = 0;        // this group's running chunk counter: staging slot = g & 1, aux barrier phase = (g >> 1) & 1
    int ui = 0;
    if (p.has_aux && leader) {
      int u0 = blockIdx.x, c0 = grp - 2;
      if (advance(u0, c0)) aux_load(u0, c0, 0);
    }
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
      const int nt = u % p.n_tiles, t = u / p.n_tiles, mt = t % p.m_tiles, z = t / p.m_tiles;
      const int m0 = mt * TC_BM, n0 = nt * BN;
      const int nch = n_chunks(u);
      const int buf = ui & 1;
      mbar_wait_t(tfull + buf, (ui >> 1) & 1, pw_tfull);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c = grp; c < nch; c += 2, ++g) {
        const int nb = n0 + c * 32;
        const int slot = g & 1;
        float* cs = reinterpret_cast<float*>(cst_g + slot * TCP_CHUNK_BYTES);
        if (leader) {
          if (p.has_aux) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the other slot's store has left it: refill it with chunk g+1's operand
            int nu = u, nc = c;
            if (advance(nu, nc)) aux_load(nu, nc, slot ^ 1);
          } else {
            // groups are committed transposed-first: at most one pending = the row-major store of chunk g-1 (other slot);
            // the transposed staging chunk and this row-major slot (chunk g-2) have been read out
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * 256 + c * 32), r);
        float v[32];
        float bias_lane = 0.0f;
        if ((p.epi == TC_EPI_BIAS || p.epi == TC_EPI_BIAS_ELU) && nb + lane < p.N) bias_lane = __ldg(p.bias + nb + lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(r[j]);
          if (p.epi == TC_EPI_BIAS || p.epi == TC_EPI_BIAS_ELU) x += __shfl_sync(0xffffffffu, bias_lane, j);
          if (p.epi == TC_EPI_BIAS_ELU) x = x > 0.0f ? x : (__expf(x) - 1.0f);
          v[j] = x;
        }
        if (p.has_aux) {
          mbar_wait_t(auxb_g + slot, (g >> 1) & 1, pw_aux);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 y = *reinterpret_cast<const float4*>(cs + row * 32 + ((j4 ^ sw) << 2));
            v[4 * j4 + 0] *= (y.x > 0.0f ? 1.0f : y.x + 1.0f);
            v[4 * j4 + 1] *= (y.y > 0.0f ? 1.0f : y.y + 1.0f);
            v[4 * j4 + 2] *= (y.z > 0.0f ? 1.0f : y.z + 1.0f);
            v[4 * j4 + 3] *= (y.w > 0.0f ? 1.0f : y.w + 1.0f);
          }
        }
        if (p.has_c) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            *reinterpret_cast<float4*>(cs + row * 32 + ((j4 ^ sw) << 2)) = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        }
        if (p.has_ct) {
#pragma unroll
          for (int j = 0; j < 32; ++j) ts[j * TC_BM + row] = v[j];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        if (leader) {
          if (p.has_ct) {
            tma_store_2d(&tmCt, ts, m0, nb);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (p.has_c) {
            tma_store_2d(&tmC, cs, nb, z * p.rows_pad + m0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      // this warp has read everything it needs from the accumulator: hand it back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(tempty + buf)) : "memory");
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (p.dbg && leader) { long long* d = p.dbg + blockIdx.x * 24 + 10 + 2 * grp; d[0] = clock64() - t_start; d[1] = w_tfull; p.dbg[blockIdx.x * 24 + 14 + grp] = w_aux; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 24 + 19] = globaltimer_ns();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ CTA-pair variant (cta_group::2)
// The GEMMs of this path are bound by SHARED-MEMORY bandwidth, not by HBM or the tensor pipe (measured with go2_gemm_set_debug, round 2): a
// 128 x 128 x 8 tf32 MMA reads 8 KB of operands per 64 cycles = the SM's whole 128 B/clk, and the TMA writes and the 3xTF32 split come on top
// (3xTF32: 48 KB of shared-memory traffic per 192 tensor cycles -> 375 cycles per K step).  Two CTAs on the two SMs of a TPC share one 256 x BN
// tile: each stages its own 128 rows of A and HALF of B, the leader's one thread issues tcgen05.mma.cta_group::2 over both SMs' operands, and each
// tensor core reads only its own SM's 4 + BN/64 KB per MMA — at BN = 256 the same 48 KB now cover 384 tensor cycles.
//   roles per CTA (18 warps): 0 TMA producer (own A rows, own half of B) · 1 MMA issuer in the leader / landing relay in the peer (tells the
//   leader that the peer's stage has landed) · 2..9 epilogue (own 128 accumulator rows, own TMEM) · 10..17 splitters (own stage -> own lo slot)
//   barriers: full / loempty / empty / tfull are CTA-local (commits are multicast to both CTAs); the leader's pfull, lofull and tempty also take
//   the peer's arrivals through the cluster window (mapa + mbarrier.arrive.shared::cluster).
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `target` of this cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t target) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(target));
  // default semantics (.release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id): the .release.cluster form compiles to
  // MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of every arrival (measured: +900 cycles per 32 KB stage in the splitter warps).  What the arrival
  // publishes is shared memory of THIS SM for the tensor core of THIS SM (made visible to the async proxy by the fence.proxy.async before it).
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// commit of the leader's MMAs, delivered to the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

constexpr int TCQ_THREADS = 320, TCQ_THREADS_X3 = 320 + 32 * TCP_SPLIT_WARPS;
constexpr int TCQ_BM = 256;                      // rows of a pair tile
// shared-memory slot of one CTA's half of B: BN/2 rows of 128 B (K-major) or ceil(BN/64) boxes of [32 rows][32 features] (MN-major)
__host__ __device__ constexpr int tcq_b_bytes(int BN) { return (BN / 2 + 31) / 32 * 4096; }
static int tcq_smem_bytes(int BN, int stages, bool x3) {
  return (stages + (x3 ? TCP_LO_SLOTS : 0)) * (TC_BM * TC_BK * 4 + tcq_b_bytes(BN)) + 4 * TCP_CHUNK_BYTES + 512 + 1024;
}
static int tcq_stages(int BN, bool x3) {
  int s = TCP_MAX_STAGES;
  while (s > 2 && tcq_smem_bytes(BN, s, x3) > TCP_SMEM_MAX) --s;
  return s;
}

template <int BN, bool X3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(X3 ? TCQ_THREADS_X3 : TCQ_THREADS, 1)
gemm_tf32_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                      const __grid_constant__ CUtensorMap tmAux, const TcParamsP p) {
  static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "pair tile width: multiple of 32 (epilogue chunks), at most 256 (one tcgen05.mma)");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int BH = BN / 2;                                   // rows (features) of B this CTA stages
  constexpr int A_BYTES = TC_BM * TC_BK * 4, B_BYTES = tcq_b_bytes(BN), STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int NBOX_B = (BH + 31) / 32;
  constexpr int NCH = BN / 32;
  const int S = p.stages;
  uint8_t* lo_ring = smem + S * STAGE_BYTES;
  uint8_t* cst = lo_ring + (X3 ? TCP_LO_SLOTS * STAGE_BYTES : 0);   // [group][2] row-major staging chunks [128 rows][128 B], SW128
  uint64_t* full_bar = (uint64_t*)(cst + 4 * TCP_CHUNK_BYTES);
  uint64_t* pfull_bar = full_bar + TCP_MAX_STAGES;             // leader: the peer's stage has landed
  uint64_t* empty_bar = pfull_bar + TCP_MAX_STAGES;
  uint64_t* tfull = empty_bar + TCP_MAX_STAGES;
  uint64_t* tempty = tfull + 2;                                // leader: 16 epilogue warps (8 local, 8 of the peer)
  uint64_t* auxb = tempty + 2;                                 // [group][2]
  uint64_t* lofull = auxb + 4;                                 // leader: 16 splitter warps
  uint64_t* loempty = lofull + TCP_LO_SLOTS;
  uint32_t* tmem_slot = (uint32_t*)(loempty + TCP_LO_SLOTS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;      // cluster index / number of clusters
  const int units = p.m_tiles * p.n_tiles * p.splits;          // m_tiles = 256-row pair tiles
  if (p.dbg && threadIdx.x == 0 && rank == 0) p.dbg[cid * 24 + 16] = globaltimer_ns();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // see gemm_tf32_persist_kernel

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    if (p.has_c) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmC) : "memory");
    if (p.has_aux) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmAux) : "memory");
    for (int s = 0; s < S; ++s) { mbar_init(full_bar + s, 1); mbar_init(pfull_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, 16); }
    for (int b = 0; b < 4; ++b) mbar_init(auxb + b, 1);
    for (int b = 0; b < TCP_LO_SLOTS; ++b) { mbar_init(lofull + b, 2 * TCP_SPLIT_WARPS); mbar_init(loempty + b, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();                                          // both CTAs' barriers exist before anything can arrive remotely
  if (warp == 1) {  // the pair's TMEM: two accumulators of up to 256 columns in each SM (same warp id in both CTAs)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (p.dbg && threadIdx.x == 0 && rank == 0) p.dbg[cid * 24 + 17] = globaltimer_ns();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.dbg && threadIdx.x == 0 && rank == 0) p.dbg[cid * 24 + 18] = globaltimer_ns();

  if (warp == 0) {
    // ===== TMA producer: own 128 rows of A, own half of B
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      long long w_empty = 0, *pw = p.dbg ? &w_empty : nullptr;
      const long long t_start = p.dbg ? clock64() : 0;
      for (int u = cid; u < units; u += ncl) {
        const int nt = u % p.n_tiles, t = u / p.n_tiles, mt = t % p.m_tiles, z = t / p.m_tiles;
        const int kb0 = z * p.kb_per_split, nkb = min(p.kb_per_split, p.total_kb - kb0);
        const int arow = mt * TCQ_BM + (int)rank * TC_BM, brow = nt * BN + (int)rank * BH;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_t(empty_bar + s, ph ^ 1, pw);
          uint8_t* sa = smem + s * STAGE_BYTES;
          if (!p.mn_major) {
            mbar_expect_tx(full_bar + s, A_BYTES + BH * TC_BK * 4);
            tma_load_2d(&tmA, full_bar + s, sa, (kb0 + kb) * TC_BK, arow);
            tma_load_2d(&tmB, full_bar + s, sa + A_BYTES, (kb0 + kb) * TC_BK, brow);
          } else {   // one [32 rows][32 features] box per 32-wide feature group
            mbar_expect_tx(full_bar + s, A_BYTES + NBOX_B * 4096);
#pragma unroll
            for (int g = 0; g < TC_BM / 32; ++g) tma_load_2d(&tmA, full_bar + s, sa + g * 4096, arow + 32 * g, (kb0 + kb) * TC_BK);
#pragma unroll
            for (int g = 0; g < NBOX_B; ++g) tma_load_2d(&tmB, full_bar + s, sa + A_BYTES + g * 4096, brow + 32 * g, (kb0 + kb) * TC_BK);
          }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
      if (p.dbg && rank == 0) { long long* d = p.dbg + cid * 24; d[0] = clock64() - t_start; d[1] = w_empty; }
    }
  } else if (warp == 1 && rank != 0) {
    // ===== landing relay (peer): stage s of this CTA has landed -> the leader may multiply it
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int u = cid; u < units; u += ncl) {
        const int z = u / (p.n_tiles * p.m_tiles);
        const int kb0 = z * p.kb_per_split, nkb = min(p.kb_per_split, p.total_kb - kb0);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar + s, ph);
          mbar_arrive_cluster(pfull_bar + s, 0);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread of the leader CTA, for both SMs)
    if (elect_one()) {
      constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TCQ_BM >> 4) << 24);
      int s = 0, ui = 0, l = 0;
      uint32_t ph = 0, lph = 0;
      long long w_tempty = 0, w_full = 0, w_lofull = 0, nst = 0;
      long long *pw_tempty = p.dbg ? &w_tempty : nullptr, *pw_full = p.dbg ? &w_full : nullptr, *pw_lofull = p.dbg ? &w_lofull : nullptr;
      const long long t_start = p.dbg ? clock64() : 0;
      for (int u = cid; u < units; u += ncl, ++ui) {
        const int z = u / (p.n_tiles * p.m_tiles);
        const int kb0 = z * p.kb_per_split, nkb = min(p.kb_per_split, p.total_kb - kb0);
        const int buf = ui & 1;
        nst += nkb;
        mbar_wait_t(tempty + buf, ((ui >> 1) & 1) ^ 1, pw_tempty);            // both CTAs' epilogues have drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + (uint32_t)(buf * 256);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_t(full_bar + s, ph, pw_full);
          mbar_wait_t(pfull_bar + s, ph, pw_full);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sl = smem_u32(lo_ring + l * STAGE_BYTES);
          const bool kmaj = !p.mn_major;
          const uint64_t da = kmaj ? make_desc_kmajor_sw128(sa) : make_desc_mnmajor_sw128_32b(sa);
          const uint64_t db = kmaj ? make_desc_kmajor_sw128(sa + A_BYTES) : make_desc_mnmajor_sw128_32b(sa + A_BYTES);
          const uint64_t kstep = kmaj ? 2 : 64;
          const uint32_t idesc = kmaj ? IDESC : (IDESC | (1u << 15) | (1u << 16));
          const bool hi_first = !X3 || !p.split_rewrite;
          if (hi_first) {
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) umma_tf32_pair(tacc, da + kstep * k, db + kstep * k, idesc, (kb | k) ? 1u : 0u);
          }
          if (X3) {
            mbar_wait_t(lofull + l, lph, pw_lofull);              // both CTAs' splitters have filled their lo slots
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t la = kmaj ? make_desc_kmajor_sw128(sl) : make_desc_mnmajor_sw128_32b(sl);
            const uint64_t lb = kmaj ? make_desc_kmajor_sw128(sl + A_BYTES) : make_desc_mnmajor_sw128_32b(sl + A_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
              umma_tf32_pair(tacc, la + kstep * k, db + kstep * k, idesc, (hi_first || (kb | k)) ? 1u : 0u);
              umma_tf32_pair(tacc, da + kstep * k, lb + kstep * k, idesc, 1u);
              if (!hi_first) umma_tf32_pair(tacc, da + kstep * k, db + kstep * k, idesc, 1u);
            }
          }
          umma_commit_pair(empty_bar + s);
          if (X3) { umma_commit_pair(loempty + l); if (++l == TCP_LO_SLOTS) { l = 0; lph ^= 1; } }
          if (++s == S) { s = 0; ph ^= 1; }
        }
        umma_commit_pair(tfull + buf);
      }
      if (p.dbg) { long long* d = p.dbg + cid * 24; d[2] = clock64() - t_start; d[3] = w_tempty; d[4] = w_full; d[5] = w_lofull; d[6] = nst; }
    }
  } else if (X3 && warp >= 10) {
    // ===== splitter warps (3xTF32): own stage -> own lo slot; every warp of BOTH CTAs arrives on the leader's lofull
    const int t = threadIdx.x - 320;
    constexpr int NV = STAGE_BYTES / 16, NT = 32 * TCP_SPLIT_WARPS, PER = (NV + NT - 1) / NT;
    int s = 0, l = 0;
    uint32_t ph = 0, lph = 0;
    long long w_full = 0, w_loempty = 0, t_work = 0;
    long long *pw_full = p.dbg ? &w_full : nullptr, *pw_loempty = p.dbg ? &w_loempty : nullptr;
    for (int u = cid; u < units; u += ncl) {
      const int z = u / (p.n_tiles * p.m_tiles);
      const int kb0 = z * p.kb_per_split, nkb = min(p.kb_per_split, p.total_kb - kb0);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_t(full_bar + s, ph, pw_full);
        const long long t_w0 = p.dbg ? clock64() : 0;
        uint4* raw = reinterpret_cast<uint4*>(smem + s * STAGE_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(lo_ring + l * STAGE_BYTES);
        uint4 v[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) if (NV % NT == 0 || t + NT * j < NV) v[j] = raw[t + NT * j];
        mbar_wait_t(loempty + l, lph ^ 1, pw_loempty);            // the MMAs that read this lo slot (in both SMs) have retired
        if (p.split_rewrite) {
#pragma unroll
          for (int j = 0; j < PER; ++j) if (NV % NT == 0 || t + NT * j < NV) { const uint4 w = split_tf32<1>(v[j]); raw[t + NT * j] = v[j]; lo[t + NT * j] = w; }
        } else {
#pragma unroll
          for (int j = 0; j < PER; ++j) if (NV % NT == 0 || t + NT * j < NV) lo[t + NT * j] = split_tf32<0>(v[j]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lofull + l, 0);
        if (p.dbg) t_work += clock64() - t_w0;
        if (++l == TCP_LO_SLOTS) { l = 0; lph ^= 1; }
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
    if (p.dbg && t == 0 && rank == 0) { long long* d = p.dbg + cid * 24; d[7] = w_full; d[8] = w_loempty; d[9] = t_work; }
  } else if (warp >= 2 && warp < 10) {
    // ===== epilogue: own 128 accumulator rows; group grp takes chunks grp, grp+2, ...; warp -> TMEM lane quadrant q; thread = one tile row
    const int grp = (warp - 2) >> 2, q = warp & 3;
    const int row = 32 * q + lane;
    const bool leader = (((warp - 2) & 3) == 0) && lane == 0;
    const int sw = row & 7;
    uint8_t* cst_g = cst + grp * 2 * TCP_CHUNK_BYTES;
    uint64_t* auxb_g = auxb + grp * 2;
    auto n_chunks = [&](int u) { return min(NCH, (p.N - (u % p.n_tiles) * BN + 31) / 32); };
    auto advance = [&](int& u, int& c) {
      c += 2;
      while (u < units && c >= n_chunks(u)) { u += ncl; c = grp; }
      return u < units;
    };
    auto aux_load = [&](int u, int c, int slot) {
      const int nt = u % p.n_tiles, mt = (u / p.n_tiles) % p.m_tiles;
      mbar_expect_tx(auxb_g + slot, TCP_CHUNK_BYTES);
      tma_load_2d(&tmAux, auxb_g + slot, cst_g + slot * TCP_CHUNK_BYTES, nt * BN + c * 32, mt * TCQ_BM + (int)rank * TC_BM);
    };
    long long w_tfull = 0, *pw_tfull = (p.dbg && leader && rank == 0) ? &w_tfull : nullptr;
    long long w_aux = 0, *pw_aux = (p.dbg && leader && rank == 0) ? &w_aux : nullptr;
    const long long t_start = p.dbg ? clock64() : 0;
    uint32_t g = 0;
    int ui = 0;
    if (p.has_aux && leader) {
      int u0 = cid, c0 = grp - 2;
      if (advance(u0, c0)) aux_load(u0, c0, 0);
    }
    for (int u = cid; u < units; u += ncl, ++ui) {
      const int nt = u % p.n_tiles, t = u / p.n_tiles, mt = t % p.m_tiles, z = t / p.m_tiles;
      const int m0 = mt * TCQ_BM + (int)rank * TC_BM, n0 = nt * BN;
      const int nch = n_chunks(u);
      const int buf = ui & 1;
      mbar_wait_t(tfull + buf, (ui >> 1) & 1, pw_tfull);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c = grp; c < nch; c += 2, ++g) {
        const int nb = n0 + c * 32;
        const int slot = g & 1;
        float* cs = reinterpret_cast<float*>(cst_g + slot * TCP_CHUNK_BYTES);
        if (leader) {
          if (p.has_aux) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            int nu = u, nc = c;
            if (advance(nu, nc)) aux_load(nu, nc, slot ^ 1);
          } else {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * 256 + c * 32), r);
        float v[32];
        float bias_lane = 0.0f;
        if ((p.epi == TC_EPI_BIAS || p.epi == TC_EPI_BIAS_ELU) && nb + lane < p.N) bias_lane = __ldg(p.bias + nb + lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(r[j]);
          if (p.epi == TC_EPI_BIAS || p.epi == TC_EPI_BIAS_ELU) x += __shfl_sync(0xffffffffu, bias_lane, j);
          if (p.epi == TC_EPI_BIAS_ELU) x = x > 0.0f ? x : (__expf(x) - 1.0f);
          v[j] = x;
        }
        if (p.has_aux) {
          mbar_wait_t(auxb_g + slot, (g >> 1) & 1, pw_aux);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 y = *reinterpret_cast<const float4*>(cs + row * 32 + ((j4 ^ sw) << 2));
            v[4 * j4 + 0] *= (y.x > 0.0f ? 1.0f : y.x + 1.0f);
            v[4 * j4 + 1] *= (y.y > 0.0f ? 1.0f : y.y + 1.0f);
            v[4 * j4 + 2] *= (y.z > 0.0f ? 1.0f : y.z + 1.0f);
            v[4 * j4 + 3] *= (y.w > 0.0f ? 1.0f : y.w + 1.0f);
          }
        }
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          *reinterpret_cast<float4*>(cs + row * 32 + ((j4 ^ sw) << 2)) = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        if (leader) {
          tma_store_2d(&tmC, cs, nb, z * p.rows_pad + m0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      // this warp has read everything it needs from the accumulator: hand it back to the leader's MMA thread
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty + buf, 0);
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (p.dbg && leader && rank == 0) { long long* d = p.dbg + cid * 24 + 10 + 2 * grp; d[0] = clock64() - t_start; d[1] = w_tfull; p.dbg[cid * 24 + 14 + grp] = w_aux; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  cluster_sync_all();                                          // nobody leaves (or frees TMEM) while the partner may still touch this SM
  if (p.dbg && threadIdx.x == 0 && rank == 0) p.dbg[cid * 24 + 19] = globaltimer_ns();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// out[i] = sum_z part[z][i] (deterministic split-K reduction); rows x cols with output leading dimension ld_out
// cols = K (+1 when the bias gradient rides along as an extra "ones" column of X^T: that column goes to db)
__global__ void tc_splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int rows, int Z, long ld_part, long split_stride, long ld_out,
                                        int cols, int k_real, float* __restrict__ db) {
  // block = 32 column-quads x 8 slice lanes: lane y sums slices y, y+8, ... of four consecutive columns of one row (ld_part is a
  // multiple of 4, slices 16-byte aligned); the eight partial sums are combined in a fixed order -> deterministic
  __shared__ float4 sm[8][32];
  const int groups = (int)(ld_part >> 2);
  const long i = (long)blockIdx.x * 32 + threadIdx.x;
  const bool live = i < (long)rows * groups;
  const int r = live ? (int)(i / groups) : 0, c0 = live ? (int)(i % groups) * 4 : 0;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live && c0 < cols) {
    const float4* src = reinterpret_cast<const float4*>(part + (long)r * ld_part + c0);
    const long zs = split_stride >> 2;
    for (int z = threadIdx.y; z < Z; z += 8) { const float4 a = __ldg(src + (long)z * zs); s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w; }
  }
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y != 0 || !live || c0 >= cols) return;
  for (int y = 1; y < 8; ++y) { const float4 a = sm[y][threadIdx.x]; s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w; }
  const float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + j;
    if (c < k_real) out[(long)r * ld_out + c] = v[j];
    else if (c < cols && db) db[r] = v[j];
  }
}

// out[r][c] = in[c][r]  (weights W -> W^T for dgrad), tile transpose through shared memory
__global__ void transpose_kernel(const float* __restrict__ in, long ldin, float* __restrict__ out, long ldout, int rows, int cols) {
  __shared__ float t[32][33];
  int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 32 + threadIdx.y;
  for (int i = 0; i < 32; i += 8) if (r + i < rows && c < cols) t[threadIdx.y + i][threadIdx.x] = in[(long)(r + i) * ldin + c];
  __syncthreads();
  int oc = blockIdx.y * 32 + threadIdx.x, orow = blockIdx.x * 32 + threadIdx.y;
  for (int i = 0; i < 32; i += 8) if (orow + i < cols && oc < rows) out[(long)(orow + i) * ldout + oc] = t[threadIdx.x][threadIdx.y + i];
}

// Up to 8 pitch-copies / transposes of small matrices in ONE launch: the operand copies an MLP engine refreshes after every optimiser
// step (zero-padded first-layer weight, W^T of the hidden layers for dgrad).  32 x 32 tiles; CTA -> (job, tile) by a prefix scan.
struct RefreshJobs {
  const float* src[8]; float* dst[8];
  int ldin[8], ldout[8], rows[8], cols[8], transpose[8], tile0[9];
  int n;
};
__global__ void refresh_weights_kernel(const __grid_constant__ RefreshJobs J) {
  __shared__ float t[32][33];
  int j = 0;
  while (j + 1 < J.n && (int)blockIdx.x >= J.tile0[j + 1]) ++j;
  const int tile = blockIdx.x - J.tile0[j], tx = (J.cols[j] + 31) / 32;
  const int c0 = (tile % tx) * 32, r0 = (tile / tx) * 32;
  const float* in = J.src[j];
  float* out = J.dst[j];
  const int rows = J.rows[j], cols = J.cols[j];
  const long ldin = J.ldin[j], ldout = J.ldout[j];
  if (!J.transpose[j]) {
    for (int i = 0; i < 32; i += 8) {
      const int r = r0 + threadIdx.y + i, c = c0 + threadIdx.x;
      if (r < rows && c < cols) out[(long)r * ldout + c] = in[(long)r * ldin + c];
    }
    return;
  }
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + threadIdx.y + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) t[threadIdx.y + i][threadIdx.x] = in[(long)r * ldin + c];
  }
  __syncthreads();
  for (int i = 0; i < 32; i += 8) {
    const int orow = c0 + threadIdx.y + i, oc = r0 + threadIdx.x;
    if (orow < cols && oc < rows) out[(long)orow * ldout + oc] = t[threadIdx.x][threadIdx.y + i];
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)sym;
  });
  return fn;
}
// row-major [rows, cols] fp32 with leading dimension ld (elements); box = box_cols floats x box_rows; 128B swizzle when the box is
// one swizzle atom wide (32 floats), dense otherwise; zero OOB fill on loads, clipping on stores
static int make_map(CUtensorMap* m, const float* ptr, long rows, long cols, long ld, int box_rows, int box_cols = TC_BK, bool atom32 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(4, "cuTensorMapEncodeTiled unavailable");
  if (((uintptr_t)ptr & 15) || ((ld * 4) & 15)) return set_error(5, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : box_cols == TC_BK ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(6, "cuTensorMapEncodeTiled failed");
  return 0;
}

template <int BN, int STAGES>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p, int splits, cudaStream_t st) {
  constexpr int smem = STAGES * (TC_BM * TC_BK * 4 + BN * TC_BK * 4) + 1024 + 256 + 4 * 32 * 36 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    GO2_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((p.M + TC_BM - 1) / TC_BM, (p.N + BN - 1) / BN, splits);
  gemm_tf32_kernel<BN, STAGES><<<grid, TC_THREADS, smem, st>>>(ta, tb, p);
  count_launch();
  return 0;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
static bool legacy_only() {
  static int v = -1;
  if (v < 0) v = getenv("GO2_GEMM_LEGACY") ? 1 : 0;
  return v == 1;
}
static int persist_min_n() {   // narrowest output the persistent kernel takes (GO2_GEMM_PERSIST_MIN_N, tuning aid)
  static int v = -1;
  if (v < 0) { const char* e = getenv("GO2_GEMM_PERSIST_MIN_N"); v = e ? atoi(e) : 1; }
  return v;
}
static bool aligned16(const void* ptr, long ld) { return (((uintptr_t)ptr & 15) == 0) && ((ld & 3) == 0); }
// tile width of the persistent kernel: 160 wins when N is just past a multiple of 128 (the "ones" column of the bias gradient)
static int persist_bn(int N) {
  const long c128 = (long)((N + 127) / 128) * (128 + 128), c160 = (long)((N + 159) / 160) * (128 + 160);
  return c160 < c128 ? 160 : 128;
}
static bool persist_ok(const TcParams& p) {
  if (legacy_only() || p.dbg || p.N < persist_min_n()) return false;
  if (p.split_stride && p.split_stride != (long)((p.M + TC_BM - 1) / TC_BM * TC_BM) * p.ldc) return false;
  if (p.C && !aligned16(p.C, p.ldc)) return false;
  if (p.Ct && !aligned16(p.Ct, p.ldct)) return false;
  if (p.epi == TC_EPI_MUL_ELU_GRAD && !(p.aux && aligned16(p.aux, p.ldaux))) return false;
  return p.C || p.Ct;
}

// GO2_GEMM_PDL=0 / go2_gemm_set_pdl(0): plain stream-ordered launches of the persistent GEMMs (A/B aid)
static int g_tc_pdl = -1;
static bool tc_pdl() {
  if (g_tc_pdl < 0) { const char* e = getenv("GO2_GEMM_PDL"); g_tc_pdl = (e && !strcmp(e, "0")) ? 0 : 1; }
  return g_tc_pdl == 1;
}
// launch with programmatic stream serialization: the grid may start while the previous kernel of the stream drains (its griddepcontrol.wait orders
// the memory traffic); captured into CUDA graphs as a programmatic dependency edge
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, int smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = tc_pdl() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int BN, bool X3>
static int launch_persist(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tct, const CUtensorMap& taux,
                          const TcParamsP& pp, cudaStream_t st) {
  const int smem = tcp_smem_bytes(BN, pp.stages, pp.has_ct != 0, X3);
  GO2_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_persist_kernel<BN, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_MAX));   // per device; cheap
  const int units = pp.m_tiles * pp.n_tiles * pp.splits;
  GO2_CUDA_OK(launch_pdl(gemm_tf32_persist_kernel<BN, X3>, min(units, sm_count()), X3 ? TCP_THREADS_X3 : TCP_THREADS, smem, st, ta, tb, tc, tct, taux, pp));
  count_launch();
  return 0;
}

// multiply precision of the tensor-core GEMMs: 3 = 3xTF32 split (default, fp32-class products), 1 = single tf32 pass (round 1's kernel;
// GO2_GEMM=tf32 or go2_gemm_set_passes(1))
static int g_tc_passes = 0;
static int tc_passes() {
  if (!g_tc_passes) { const char* e = getenv("GO2_GEMM"); g_tc_passes = (e && !strcmp(e, "tf32")) ? 1 : 3; }
  return g_tc_passes;
}

// how the 3xTF32 kernel gets its hi operand: 0 = hardware truncation of the raw word (default), 1 = raw stage rewritten with rn_tf32 (GO2_GEMM_SPLIT=rewrite)
static int g_tc_split = -1;
static int tc_split_rewrite() {
  if (g_tc_split < 0) { const char* e = getenv("GO2_GEMM_SPLIT"); g_tc_split = (e && !strcmp(e, "rewrite")) ? 1 : 0; }
  return g_tc_split;
}

static long long* g_tc_dbg = nullptr;   // go2_gemm_set_debug

// CTA-pair kernel on / off (GO2_GEMM_PAIR=0 or go2_gemm_set_pair(0): the one-CTA persistent kernel everywhere; A/B aid)
static int g_tc_pair = -1;
static bool tc_pair() {
  if (g_tc_pair < 0) { const char* e = getenv("GO2_GEMM_PAIR"); g_tc_pair = (e && !strcmp(e, "0")) ? 0 : 1; }
  return g_tc_pair == 1;
}
// tile width of the pair kernel: the padded width ceil(N / BN) * BN decides, wider tiles win ties (fewer shared-memory bytes per flop)
static int pair_bn(int N) {
  if (N <= 64) return 64;
  const int cand[4] = {256, 192, 160, 128};
  int best = 256; long cost = -1;
  for (int i = 0; i < 4; ++i) {
    const long c = (long)((N + cand[i] - 1) / cand[i]) * cand[i];
    if (cost < 0 || c < cost) { cost = c; best = cand[i]; }
  }
  return best;
}
// Pairs pay off when there are enough 256-row tiles to go round: with a handful of units (rollout inference, 4096 rows = 16 pair tiles) the one-CTA
// kernel spreads the same work over four times as many SMs (measured: 17 us vs 22 us for 4096 x 512 x 264).  `units` = pair tiles x splits.
static bool pair_ok(const TcParams& p, int splits = 1) {
  if (!(tc_pair() && tc_passes() == 3 && !p.Ct && p.C && p.M > TC_BM && (sm_count() & 1) == 0)) return false;
  const long units = (long)((p.M + TCQ_BM - 1) / TCQ_BM) * ((p.N + pair_bn(p.N) - 1) / pair_bn(p.N)) * splits;
  return units >= sm_count() / 4;
}

template <int BN>
static int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& taux, const TcParamsP& pp, cudaStream_t st) {
  const int smem = tcq_smem_bytes(BN, pp.stages, true);
  GO2_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_pair_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_MAX));
  const int units = pp.m_tiles * pp.n_tiles * pp.splits;
  GO2_CUDA_OK(launch_pdl(gemm_tf32_pair_kernel<BN, true>, 2 * min(units, sm_count() / 2), TCQ_THREADS_X3, smem, st, ta, tb, tc, taux, pp));
  count_launch();
  return 0;
}

// C = A B^T on CTA pairs (3xTF32): 256-row tiles, split-K slices rows_pad = roundup(M, 256) rows apart
static int gemm_tc_pair(const float* A, long lda, const float* B, long ldb, const TcParams& p, int splits, cudaStream_t st) {
  const int BN = pair_bn(p.N);
  TcParamsP pp{};
  pp.stages = tcq_stages(BN, true);
  pp.M = p.M; pp.N = p.N; pp.K = p.K;
  pp.m_tiles = (p.M + TCQ_BM - 1) / TCQ_BM; pp.n_tiles = (p.N + BN - 1) / BN;
  pp.total_kb = (p.K + TC_BK - 1) / TC_BK;
  pp.kb_per_split = (pp.total_kb + splits - 1) / splits;
  pp.splits = (pp.total_kb + pp.kb_per_split - 1) / pp.kb_per_split;
  pp.rows_pad = pp.m_tiles * TCQ_BM;
  pp.bias = p.bias; pp.epi = p.epi; pp.has_c = 1; pp.has_ct = 0; pp.has_aux = p.epi == TC_EPI_MUL_ELU_GRAD;
  pp.mn_major = p.mn_major; pp.split_rewrite = tc_split_rewrite(); pp.dbg = g_tc_dbg;
  if (splits > 1 && p.split_stride != (long)pp.rows_pad * p.ldc) return set_error(5, "gemm_tc_pair: split stride must be roundup(M,256) * ldc");
  CUtensorMap ta, tb, tc, taux;
  int rc = p.mn_major ? make_map(&ta, A, p.K, p.M, lda, 32, 32, true) : make_map(&ta, A, p.M, p.K, lda, TC_BM);
  if (rc) return rc;
  rc = p.mn_major ? make_map(&tb, B, p.K, p.N, ldb, 32, 32, true) : make_map(&tb, B, p.N, p.K, ldb, BN / 2);
  if (rc) return rc;
  rc = make_map(&tc, p.C, splits > 1 ? (long)splits * pp.rows_pad : p.M, p.N, p.ldc, TC_BM);
  if (rc) return rc;
  taux = ta;
  if (pp.has_aux) { rc = make_map(&taux, p.aux, p.M, p.N, p.ldaux, TC_BM); if (rc) return rc; }
  switch (BN) {
    case 64: rc = launch_pair<64>(ta, tb, tc, taux, pp, st); break;
    case 128: rc = launch_pair<128>(ta, tb, tc, taux, pp, st); break;
    case 160: rc = launch_pair<160>(ta, tb, tc, taux, pp, st); break;
    case 192: rc = launch_pair<192>(ta, tb, tc, taux, pp, st); break;
    default: rc = launch_pair<256>(ta, tb, tc, taux, pp, st); break;
  }
  if (rc) return rc;
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

// split-K slices of C sit rows_pad = roundup(M, 128) rows apart so that one 2-D map covers all of them
static int gemm_tc_persist(const float* A, long lda, const float* B, long ldb, const TcParams& p, int splits, cudaStream_t st) {
  const bool x3 = tc_passes() == 3;
  // the 160-wide tile has no room for the transposed staging buffers; with the lo side ring of 3xTF32 it only keeps 2 stages: used for N <= 160 only
  const int BN = (p.Ct || (x3 && p.N > 160)) ? 128 : persist_bn(p.N);
  TcParamsP pp{};
  pp.stages = tcp_stages(BN, p.Ct != nullptr, x3);
  pp.M = p.M; pp.N = p.N; pp.K = p.K;
  pp.m_tiles = (p.M + TC_BM - 1) / TC_BM; pp.n_tiles = (p.N + BN - 1) / BN;
  pp.total_kb = (p.K + TC_BK - 1) / TC_BK;
  pp.kb_per_split = (pp.total_kb + splits - 1) / splits;
  pp.splits = (pp.total_kb + pp.kb_per_split - 1) / pp.kb_per_split;      // every slice gets at least one K block
  pp.rows_pad = pp.m_tiles * TC_BM;
  pp.bias = p.bias; pp.epi = p.epi; pp.has_c = p.C != nullptr; pp.has_ct = p.Ct != nullptr; pp.has_aux = p.epi == TC_EPI_MUL_ELU_GRAD;
  if (splits > 1 && p.split_stride != (long)pp.rows_pad * p.ldc) return set_error(5, "gemm_tc_persist: split stride must be roundup(M,128) * ldc");
  CUtensorMap ta, tb, tc, tct, taux;
  pp.mn_major = p.mn_major; pp.split_rewrite = tc_split_rewrite(); pp.dbg = g_tc_dbg;
  int rc = p.mn_major ? make_map(&ta, A, p.K, p.M, lda, 32, 32, true) : make_map(&ta, A, p.M, p.K, lda, TC_BM);
  if (rc) return rc;
  rc = p.mn_major ? make_map(&tb, B, p.K, p.N, ldb, 32, 32, true) : make_map(&tb, B, p.N, p.K, ldb, BN);
  if (rc) return rc;
  tc = ta; tct = ta; taux = ta;
  if (p.C) { rc = make_map(&tc, p.C, splits > 1 ? (long)splits * pp.rows_pad : p.M, p.N, p.ldc, TC_BM); if (rc) return rc; }
  if (p.Ct) { rc = make_map(&tct, p.Ct, p.N, p.M, p.ldct, 32, TC_BM); if (rc) return rc; }
  if (pp.has_aux) { rc = make_map(&taux, p.aux, p.M, p.N, p.ldaux, TC_BM); if (rc) return rc; }
  if (x3) rc = BN == 160 ? launch_persist<160, true>(ta, tb, tc, tct, taux, pp, st) : launch_persist<128, true>(ta, tb, tc, tct, taux, pp, st);
  else rc = BN == 160 ? launch_persist<160, false>(ta, tb, tc, tct, taux, pp, st) : launch_persist<128, false>(ta, tb, tc, tct, taux, pp, st);
  if (rc) return rc;
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

// C[M,N] = A[M,K] B[N,K]^T with both operands K-major; see TcParams for the epilogue
static int gemm_tc(const float* A, long lda, const float* B, long ldb, TcParams p, int splits, cudaStream_t st) {
  if (persist_ok(p) && pair_ok(p) && splits == 1) return gemm_tc_pair(A, lda, B, ldb, p, 1, st);
  if (persist_ok(p)) return gemm_tc_persist(A, lda, B, ldb, p, splits, st);
  // the one-tile-per-CTA kernel below is the single-pass tf32 kernel of round 1 (debug stamps, GO2_GEMM_LEGACY, operands the TMA store path
  // cannot take): never a silent precision downgrade of the 3xTF32 default
  if (tc_passes() == 3 && !legacy_only() && !p.dbg)
    return set_error(5, "gemm_tc: operands do not meet the persistent kernel's alignment rules (16-byte aligned C / aux, ld % 4 == 0); "
                        "the single-pass tf32 kernel needs go2_gemm_set_passes(1)");
  const int BN = p.N > 64 ? 128 : 64;
  CUtensorMap ta, tb;
  int rc = make_map(&ta, A, p.M, p.K, lda, TC_BM);
  if (rc) return rc;
  rc = make_map(&tb, B, p.N, p.K, ldb, BN);
  if (rc) return rc;
  const int total_kb = (p.K + TC_BK - 1) / TC_BK;
  p.kb_per_split = (total_kb + splits - 1) / splits;
  // 2 x 32 KB stages + 17 KB epilogue staging = 82 KB -> two CTAs per SM: the epilogue of one overlaps the mainloop of the other
  if (BN == 128) rc = launch_tc<128, 2>(ta, tb, p, splits, st);
  else rc = launch_tc<64, 3>(ta, tb, p, splits, st);
  if (rc) return rc;
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace go2

using namespace go2;

extern "C" {

int go2_gemm_set_passes(int passes) {
  if (passes != 1 && passes != 3) return set_error(1, "go2_gemm_set_passes: 1 (single tf32 pass) or 3 (3xTF32 split)");
  g_tc_passes = passes;
  return 0;
}
int go2_gemm_get_passes(void) { return tc_passes(); }
int go2_gemm_set_pdl(int on) {
  if (on != 0 && on != 1) return set_error(1, "go2_gemm_set_pdl: 0 or 1");
  g_tc_pdl = on;
  return 0;
}
int go2_gemm_set_pair(int on) {
  if (on != 0 && on != 1) return set_error(1, "go2_gemm_set_pair: 0 or 1");
  g_tc_pair = on;
  return 0;
}
int go2_gemm_set_debug(long long* counters) { g_tc_dbg = counters; return 0; }
int go2_gemm_set_split(int rewrite) {
  if (rewrite != 0 && rewrite != 1) return set_error(1, "go2_gemm_set_split: 0 (lo only, hardware truncation is hi) or 1 (stage rewritten with rn_tf32)");
  g_tc_split = rewrite;
  return 0;
}

// Y[M,N] (and optionally Yt[N,M]) = act(X[M,K] W[N,K]^T + b)
int go2_linear_forward_tc(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy, float* Yt, int ldyt, int M, int N,
                          int K, int act, void* stream) {
  TcParams p{};
  p.M = M; p.N = N; p.K = K; p.C = Y; p.ldc = ldy; p.Ct = Yt; p.ldct = ldyt; p.bias = b; p.epi = act ? TC_EPI_BIAS_ELU : TC_EPI_BIAS;
  return gemm_tc(X, ldx, W, ldw, p, 1, (cudaStream_t)stream);
}

// dX[M,K] (and dXt[K,M]) = (dZ[M,N] Wt[K,N]^T) * ELU'(act_in);  Wt = W^T stored [K, N]
int go2_linear_dgrad_tc(const float* dZ, int lddz, const float* Wt, int ldwt, const float* act_in, int ldact, const float* act_in_t, int ldact_t,
                        float* dX, int lddx, float* dXt, int lddxt, int M, int N, int K, void* stream) {
  TcParams p{};
  p.M = M; p.N = K; p.K = N; p.C = dX; p.ldc = lddx; p.Ct = dXt; p.ldct = lddxt; p.aux = act_in; p.ldaux = ldact; p.aux_t = act_in_t; p.ldaux_t = ldact_t;
  p.epi = (act_in || act_in_t) ? TC_EPI_MUL_ELU_GRAD : TC_EPI_PLAIN;
  return gemm_tc(dZ, lddz, Wt, ldwt, p, 1, (cudaStream_t)stream);
}

// dW[N,K] = dZt[N,M] Xt[K,M]^T, split over the M rows through `workspace` (deterministic)
int go2_linear_wgrad_tc(const float* dZt, int lddzt, const float* Xt, int ldxt, float* dW, int lddw, float* db, int M, int N, int K, float* workspace,
                        long workspace_floats, void* stream) {
  // db != NULL: Xt carries one extra row of ones after its K feature rows, so column K of the product is the bias gradient
  cudaStream_t st = (cudaStream_t)stream;
  const int k_real = K;
  if (db) K = K + 1;
  const long ldp = (K + 3) / 4 * 4;
  const int total_kb = (M + TC_BK - 1) / TC_BK;
  const int rows_pad = (N + TC_BM - 1) / TC_BM * TC_BM;                   // slice pitch in rows (what the persistent kernel's C map needs)
  const bool persist = !legacy_only() && K >= persist_min_n() && workspace && (long)rows_pad * ldp <= workspace_floats;
  const int rows_slice = persist ? rows_pad : N;                          // the legacy kernel takes any slice pitch
  const int BN = persist ? persist_bn(K) : (K > 64 ? 128 : 64);
  const int tiles = ((N + TC_BM - 1) / TC_BM) * ((K + BN - 1) / BN);
  int splits;
  if (persist) splits = max(1, min(min(total_kb / 8, 48), (sm_count() + tiles / 2) / tiles));   // ~ one unit per SM
  else splits = max(1, min(min(total_kb / 4, 64), (296 + tiles - 1) / tiles));
  while (splits > 1 && (long)splits * rows_slice * ldp > workspace_floats) --splits;
  if (!workspace) splits = 1;
  splits = (total_kb + (total_kb + splits - 1) / splits - 1) / ((total_kb + splits - 1) / splits);   // drop empty trailing slices
  TcParams p{};
  p.M = N; p.N = K; p.K = M; p.epi = TC_EPI_PLAIN;
  if (splits == 1 && lddw % 4 == 0 && !db) {
    p.C = dW; p.ldc = lddw;
    return gemm_tc(dZt, lddzt, Xt, ldxt, p, 1, st);
  }
  if (!workspace || (long)rows_slice * ldp > workspace_floats) return set_error(5, "go2_linear_wgrad_tc: workspace too small");
  p.C = workspace; p.ldc = ldp; p.split_stride = (long)rows_slice * ldp;
  int rc = gemm_tc(dZt, lddzt, Xt, ldxt, p, splits, st);
  if (rc) return rc;
  const long n = (long)N * (ldp / 4);
  tc_splitk_reduce_kernel<<<(unsigned)((n + 31) / 32), dim3(32, 8), 0, st>>>(workspace, dW, N, splits, ldp, p.split_stride, lddw, K, k_real, db);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

// dW[N,K] = dZ[M,N]^T X[M,K] straight from the ROW-MAJOR activations / gradients (MN-major tf32 operands: no transposed copies).
// db != NULL: X carries a column of ones at column K (ldx > K), so column K of the product is the bias gradient.
int go2_linear_wgrad_tc_rm(const float* dZ, int lddz, const float* X, int ldx, float* dW, int lddw, float* db, int M, int N, int K, float* workspace,
                           long workspace_floats, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int k_real = K;
  if (db) K = K + 1;
  if (!workspace) return set_error(5, "go2_linear_wgrad_tc_rm: workspace required");
  const long ldp = (K + 3) / 4 * 4;
  const int total_kb = (M + TC_BK - 1) / TC_BK;
  TcParams p{};
  p.M = N; p.N = K; p.K = M; p.epi = TC_EPI_PLAIN; p.mn_major = 1;
  p.C = workspace; p.ldc = ldp;
  const bool pair = pair_ok(p, 64) && (long)((N + TCQ_BM - 1) / TCQ_BM * TCQ_BM) * ldp <= workspace_floats;
  const int tile_m = pair ? TCQ_BM : TC_BM;
  const int rows_pad = (N + tile_m - 1) / tile_m * tile_m;
  const int BN = pair ? pair_bn(K) : persist_bn(K);
  const int tiles = ((N + tile_m - 1) / tile_m) * ((K + BN - 1) / BN);
  // ~ one unit per SM (one-CTA kernel) / at most one unit per CTA pair
  int splits = pair ? max(1, min(min(total_kb / 8, 48), (sm_count() / 2) / tiles)) : max(1, min(min(total_kb / 8, 48), (sm_count() + tiles / 2) / tiles));
  while (splits > 1 && (long)splits * rows_pad * ldp > workspace_floats) --splits;
  if ((long)rows_pad * ldp > workspace_floats) return set_error(5, "go2_linear_wgrad_tc_rm: workspace too small");
  splits = (total_kb + (total_kb + splits - 1) / splits - 1) / ((total_kb + splits - 1) / splits);
  p.split_stride = (long)rows_pad * ldp;
  int rc = pair ? gemm_tc_pair(dZ, lddz, X, ldx, p, splits, st) : gemm_tc_persist(dZ, lddz, X, ldx, p, splits, st);
  if (rc) return rc;
  const long n = (long)N * (ldp / 4);
  tc_splitk_reduce_kernel<<<(unsigned)((n + 31) / 32), dim3(32, 8), 0, st>>>(workspace, dW, N, splits, ldp, p.split_stride, lddw, K, k_real, db);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

// profiling aid: forward GEMM with per-CTA clock64 stamps {start, setup done, accumulator ready, epilogue done, end}
int go2_linear_forward_tc_dbg(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy, float* Yt, int ldyt, int M, int N,
                              int K, int act, long long* dbg, void* stream) {
  TcParams p{};
  p.M = M; p.N = N; p.K = K; p.C = Y; p.ldc = ldy; p.Ct = Yt; p.ldct = ldyt; p.bias = b; p.epi = act ? TC_EPI_BIAS_ELU : TC_EPI_BIAS; p.dbg = dbg;
  return gemm_tc(X, ldx, W, ldw, p, 1, (cudaStream_t)stream);
}

int go2_refresh_weights(int n, const float* const* src, const int* ldin, float* const* dst, const int* ldout, const int* rows, const int* cols,
                        const int* transpose, void* stream) {
  if (n < 1 || n > 8 || !src || !dst || !ldin || !ldout || !rows || !cols || !transpose) return set_error(1, "go2_refresh_weights: 1..8 jobs, no null arrays");
  RefreshJobs J{};
  J.n = n;
  int tiles = 0;
  for (int j = 0; j < n; ++j) {
    if (!src[j] || !dst[j] || rows[j] <= 0 || cols[j] <= 0) return set_error(1, "go2_refresh_weights: bad job");
    J.src[j] = src[j]; J.dst[j] = dst[j]; J.ldin[j] = ldin[j]; J.ldout[j] = ldout[j]; J.rows[j] = rows[j]; J.cols[j] = cols[j]; J.transpose[j] = transpose[j];
    J.tile0[j] = tiles;
    tiles += ((rows[j] + 31) / 32) * ((cols[j] + 31) / 32);
  }
  J.tile0[n] = tiles;
  refresh_weights_kernel<<<tiles, dim3(32, 8), 0, (cudaStream_t)stream>>>(J);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_transpose(const float* in, int ldin, float* out, int ldout, int rows, int cols, void* stream) {
  dim3 block(32, 8), grid((cols + 31) / 32, (rows + 31) / 32);
  transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, ldin, out, ldout, rows, cols);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
