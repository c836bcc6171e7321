// rl_kernels.cu — sm_100a kernels of the PPO trainer path (rsl_rl restated) and their C ABI.
//
// Reference semantics (all fp32):
//   ActorCritic MLPs                 rsl_rl/modules/actor_critic.py:38-136
//   PPO.act / process_env_step       rsl_rl/algorithms/ppo.py:90-114
//   RolloutStorage.compute_returns   rsl_rl/storage/rollout_storage.py:123-137
//   mini_batch_generator (gather)    rsl_rl/storage/rollout_storage.py:147-183
//   PPO.update losses / KL / LR      rsl_rl/algorithms/ppo.py:120-187
//   clip_grad_norm_ + Adam           torch.nn.utils.clip_grad_norm_, torch.optim.Adam defaults (ppo.py:67,176-177)
//
// This file holds the CUDA-core (SIMT) kernels: the elementwise / reduction work of the trainer and a tiled fp32 GEMM used
// for small or oddly shaped layers and as the numerical cross-check of the tcgen05 GEMM (gemm_tc.cu).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/go2_b200.h"
#include "common.cuh"
#include "env_step_core.cuh"  // philox

namespace go2 {

// ================================================================================================ SIMT GEMM
// C[i][j] (+)= sum_k A(i,k) B(k,j) ;  A(i,k) = A[i*sai + k*sak], B(k,j) = B[k*sbk + j*sbj]
enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_ELU = 2, EPI_MUL_ELU_GRAD = 3 };

template <bool A_K_CONTIG, bool B_J_CONTIG>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, long sai, long sak, const float* __restrict__ B, long sbk,
                                                        long sbj, float* __restrict__ C, long ldc, int M, int N, int K, int k_chunk,
                                                        const float* __restrict__ bias, int epi, const float* __restrict__ aux, long ldaux,
                                                        long split_stride, float* __restrict__ Ct, long ldct) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4], Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_chunk, kend = min(K, kbeg + k_chunk);
  float acc[4][4] = {};
  for (int kt = kbeg; kt < kend; kt += BK) {
    if (A_K_CONTIG) {
      const int i = tid / 4, k4 = (tid % 4) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gi = i0 + i, gk = kt + k4 + q;
        As[k4 + q][i] = (gi < M && gk < kend) ? A[gi * sai + gk * sak] : 0.0f;
      }
    } else {
      const int k = tid / 16, i4 = (tid % 16) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gi = i0 + i4 + q, gk = kt + k;
        As[k][i4 + q] = (gi < M && gk < kend) ? A[gi * sai + gk * sak] : 0.0f;
      }
    }
    if (B_J_CONTIG) {
      const int k = tid / 16, j4 = (tid % 16) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gj = j0 + j4 + q, gk = kt + k;
        Bs[k][j4 + q] = (gj < N && gk < kend) ? B[gk * sbk + gj * sbj] : 0.0f;
      }
    } else {
      const int j = tid / 4, k4 = (tid % 4) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gj = j0 + j, gk = kt + k4 + q;
        Bs[k4 + q][j] = (gj < N && gk < kend) ? B[gk * sbk + gj * sbj] : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = As[k][ty * 4 + q]; b[q] = Bs[k][tx * 4 + q]; }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
    }
    __syncthreads();
  }
  float* Cz = C + (long)blockIdx.z * split_stride;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int gi = i0 + ty * 4 + p;
    if (gi >= M) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int gj = j0 + tx * 4 + q;
      if (gj >= N) continue;
      float v = acc[p][q];
      if (epi == EPI_BIAS || epi == EPI_BIAS_ELU) v += bias[gj];
      if (epi == EPI_BIAS_ELU) v = v > 0.0f ? v : expm1f(v);
      if (epi == EPI_MUL_ELU_GRAD) { float y = aux[gi * ldaux + gj]; v *= (y > 0.0f ? 1.0f : y + 1.0f); }
      if (C) Cz[gi * ldc + gj] = v;
      if (Ct) Ct[gj * ldct + gi] = v;
    }
  }
}

// out[i] = sum_z part[z][i]  (deterministic split-K reduction); n = elements, Z = splits
__global__ void splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, long n, int Z, long ld_out, int cols) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0;
  for (int z = 0; z < Z; ++z) s += part[(long)z * n + i];
  out[(i / cols) * ld_out + (i % cols)] = s;
}

// db[j] = sum_m dY[m][j], deterministic two-stage: grid (col blocks of 32, row chunks) -> part[chunk][j], then a fixed-order sum
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ dY, long ld, float* __restrict__ part, int M, int N, int rows_per_chunk) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32, j = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float s = 0;
  if (j < N) for (int m = r0 + ty; m < r1; m += 8) s += dY[(long)m * ld + j];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < N) { float t = 0; for (int r = 0; r < 8; ++r) t += sm[r][tx]; part[(long)blockIdx.y * N + j] = t; }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, float* __restrict__ db, int N, int chunks) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  float t = 0;
  for (int c = 0; c < chunks; ++c) t += part[(long)c * N + j];
  db[j] = t;
}

// ================================================================================================ rollout kernels
// PPO.act tail (ppo.py:94-101; actor_critic.py:119-125): a = mu + std * z, log_prob, and the transition rows.
// mu [N,A] -> actions [N,A], logp [N], mu_out, sigma_out [N,A].  z from Philox(seed; env, step, stream 16, block).
__global__ void sample_actions_kernel(const float* __restrict__ mu, const float* __restrict__ std_param, float* __restrict__ actions,
                                      float* __restrict__ logp, float* __restrict__ mu_out, float* __restrict__ sigma_out, int N, int A,
                                      uint32_t seed_lo, uint32_t seed_hi, uint32_t step, const uint32_t* __restrict__ d_step, int env_offset) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  if (d_step) step = *d_step;      // device-resident step counter (CUDA-graph replays of the rollout)
  float lp = 0;
  for (int b = 0; b < (A + 3) / 4; ++b) {
    U4 r = philox((uint32_t)(env_offset + e), step, 16u, (uint32_t)b, seed_lo, seed_hi);
    // Box-Muller on two pairs of uniforms in (0,1]
    float u0 = ((float)(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f), u1 = u01(r.y), u2 = ((float)(r.z >> 8) + 1.0f) * (1.0f / 16777216.0f), u3 = u01(r.w);
    float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
    float z[4] = {r0 * cosf(6.283185307179586f * u1), r0 * sinf(6.283185307179586f * u1), r1 * cosf(6.283185307179586f * u3), r1 * sinf(6.283185307179586f * u3)};
    for (int q = 0; q < 4; ++q) {
      const int k = 4 * b + q;
      if (k >= A) break;
      const float m = mu[(long)e * A + k], s = std_param[k];
      const float a = m + s * z[q];
      actions[(long)e * A + k] = a;
      mu_out[(long)e * A + k] = m;
      sigma_out[(long)e * A + k] = s;
      const float d = (a - m);
      lp += -(d * d) / (2.0f * s * s) - logf(s) - 0.9189385332046727f;
    }
  }
  logp[e] = lp;
}

// PPO.process_env_step (ppo.py:104-111): rewards += gamma * values * time_outs; store rewards and dones
// perm (optional): storage row e holds env perm[e] (CTS stores teacher envs first, cts.py:144-152)
__global__ void process_env_step_kernel(const float* __restrict__ rew, const uint8_t* __restrict__ dones, const uint8_t* __restrict__ time_outs,
                                        const float* __restrict__ values, float* __restrict__ rew_out, uint8_t* __restrict__ dones_out, int N,
                                        float gamma, const int64_t* __restrict__ perm) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  const long src = perm ? perm[e] : e;
  float r = rew[src];
  if (time_outs) r += gamma * (values[e] * (float)time_outs[src]);
  rew_out[e] = r;
  dones_out[e] = dones[src];
}

// RolloutStorage.compute_returns (rollout_storage.py:123-134): reverse scan per env; also the sums for the normalisation
__global__ void gae_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
                                const float* __restrict__ last_values, float* __restrict__ returns, float* __restrict__ adv, int T, int N,
                                float gamma, float lam, double* __restrict__ stats) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double s1 = 0, s2 = 0;
  if (e < N) {
    float a = 0, next_v = last_values[e];
    for (int t = T - 1; t >= 0; --t) {
      const long o = (long)t * N + e;
      const float nt = 1.0f - (float)dones[o];
      const float v = values[o];
      const float delta = rewards[o] + nt * gamma * next_v - v;
      a = delta + nt * gamma * lam * a;
      const float ret = a + v;
      returns[o] = ret;
      const float ad = ret - v;
      adv[o] = ad;
      s1 += ad; s2 += (double)ad * ad;
      next_v = v;
    }
  }
  // block reduce then one atomic per block (double: order effects are below fp32 resolution of the result)
  __shared__ double r1[256], r2[256];
  r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) { r1[threadIdx.x] += r1[threadIdx.x + s]; r2[threadIdx.x] += r2[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { atomicAdd(stats, r1[0]); atomicAdd(stats + 1, r2[0]); }
}
// advantages = (adv - mean) / (std + 1e-8), unbiased std (rollout_storage.py:136-137)
__global__ void adv_normalize_kernel(float* __restrict__ adv, long n, const double* __restrict__ stats, double count) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mean = stats[0] / count;
  const double var = (stats[1] - count * mean * mean) / (count - 1.0);
  const float sd = (float)sqrt(var > 0 ? var : 0.0);
  adv[i] = (adv[i] - (float)mean) / (sd + 1e-8f);
}

// mini_batch_generator gather (rollout_storage.py:173-181): dst[i, 0:w] = src[idx[i], 0:w], dst row stride ldd; padding columns are zero
// except column w, which is 1 (the "ones" column that yields the bias gradient in the row-major wgrad; the first layer's padded weight
// has zeros there, so the forward pass does not see it)
__global__ void gather_rows_kernel(const float* __restrict__ src, int w, const int64_t* __restrict__ idx, float* __restrict__ dst, int ldd, long n,
                                   float* __restrict__ dst_t) {
  __shared__ float tile[8][33];
  const long i = (long)blockIdx.x * blockDim.y + threadIdx.y;
  const long s = (i < n) ? (idx ? idx[i] : i) : 0;
  for (int c0 = 0; c0 < ldd; c0 += 32) {
    const int c = c0 + threadIdx.x;
    float v = (i < n && c < w) ? src[s * w + c] : (c == w ? 1.0f : 0.0f);
    if (i < n && c < ldd && dst) dst[i * ldd + c] = v;
    if (dst_t) {  // transposed copy dst_t[c][i] (row pitch n), staged through shared memory so both sides stay coalesced
      tile[threadIdx.y][threadIdx.x] = v;
      __syncthreads();
      const int tr = threadIdx.x % 8, tc = threadIdx.y * 4 + threadIdx.x / 8;  // 8 rows i x 32 cols -> each thread one element
      const long ii = (long)blockIdx.x * blockDim.y + tr;
      const int cc = c0 + tc;
      if (ii < n && cc < w) dst_t[(long)cc * n + ii] = tile[tr][tc];
      __syncthreads();
    }
  }
}

// ================================================================================================ PPO loss, forward + backward
// Per sample (ppo.py:131-171).  Outputs d loss / d mu [M,A], d loss / d value [M]; accumulates into `scal`:
//   [0] sum KL, [1] sum surrogate, [2] sum value loss, [3] sum entropy, [4..4+A) d loss / d std
struct PpoLossArgs {
  const float* mu; const float* std_param; const float* value; const float* actions; const float* old_logp; const float* adv;
  const float* target_values; const float* returns; const float* old_mu; const float* old_sigma;
  float* dmu; float* dmu_t; float* dvalue; float* scal;
  int M, A; float clip, value_coef, entropy_coef; int use_clipped_value_loss; float inv_count;  // 1 / (global mini-batch rows)
  // CTS (cts.py / moe_cts.py:166-168): the surrogate is mean over the teacher rows [0, split) PLUS mean over the student rows [split, M)
  int split; float inv_count_a, inv_count_b;
};
__global__ void __launch_bounds__(256) ppo_loss_kernel(PpoLossArgs p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float kl = 0, surr = 0, surr_b = 0, vl = 0, ent = 0;
  float dstd[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) dstd[k] = 0;
  if (i < p.M) {
    float lp = 0;
    for (int k = 0; k < p.A; ++k) {
      const float m = p.mu[(long)i * p.A + k], s = p.std_param[k], a = p.actions[(long)i * p.A + k];
      const float d = a - m;
      lp += -(d * d) / (2.0f * s * s) - logf(s) - 0.9189385332046727f;
      const float om = p.old_mu[(long)i * p.A + k], os = p.old_sigma[(long)i * p.A + k];
      kl += logf(s / os + 1.e-5f) + (os * os + (om - m) * (om - m)) / (2.0f * s * s) - 0.5f;
      ent += 0.5f + 0.9189385332046727f + logf(s);
    }
    const float A_ = p.adv[i];
    const float ratio = expf(lp - p.old_logp[i]);
    const float s1 = -A_ * ratio, s2 = -A_ * fminf(fmaxf(ratio, 1.0f - p.clip), 1.0f + p.clip);
    if (i < p.split) surr = fmaxf(s1, s2); else surr_b = fmaxf(s1, s2);
    // d surr / d lp : the unclipped branch (or a tie) carries gradient -A ratio; the clipped branch only inside the clip range
    float dlp;
    if (s1 >= s2) dlp = -A_ * ratio;
    else dlp = (ratio > 1.0f - p.clip && ratio < 1.0f + p.clip) ? -A_ * ratio : 0.0f;
    dlp *= (i < p.split) ? p.inv_count_a : p.inv_count_b;
    const float v = p.value[i], tv = p.target_values[i], ret = p.returns[i];
    float dv;
    if (p.use_clipped_value_loss) {
      const float diff = v - tv;
      const float vc = tv + fminf(fmaxf(diff, -p.clip), p.clip);
      const float l1 = (v - ret) * (v - ret), l2 = (vc - ret) * (vc - ret);
      vl = fmaxf(l1, l2);
      if (l1 >= l2) dv = 2.0f * (v - ret);
      else dv = (diff > -p.clip && diff < p.clip) ? 2.0f * (vc - ret) : 0.0f;
    } else { vl = (ret - v) * (ret - v); dv = 2.0f * (v - ret); }
    p.dvalue[i] = p.value_coef * dv * p.inv_count;
    for (int k = 0; k < p.A; ++k) {
      const float m = p.mu[(long)i * p.A + k], s = p.std_param[k], a = p.actions[(long)i * p.A + k];
      const float d = a - m;
      const float g = dlp * d / (s * s);                                  // d lp / d mu = (a - mu) / s^2
      p.dmu[(long)i * p.A + k] = g;
      if (p.dmu_t) p.dmu_t[(long)k * p.M + i] = g;
      dstd[k] = dlp * (d * d / (s * s * s) - 1.0f / s)                      // d lp / d s
                - p.entropy_coef * p.inv_count / s;                         // - coef * d mean(entropy) / d s
    }
  }
  // block reduction: warp-shuffle butterflies (fixed tree), one shared-memory hop across the 8 warps, then one atomic per block per scalar
  __shared__ float red[8][22];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto warp_sum = [](float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  float vals[5] = {kl, surr, vl, ent, surr_b};
#pragma unroll
  for (int k = 0; k < 5; ++k) { const float t = warp_sum(vals[k]); if (lane == 0) red[warp][k] = t; }
#pragma unroll
  for (int k = 0; k < 16; ++k) { const float t = warp_sum(dstd[k]); if (lane == 0) red[warp][5 + k] = t; }
  __syncthreads();
  if (threadIdx.x < 21) {
    const int k = threadIdx.x;
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][k];
    if (k < 4) atomicAdd(p.scal + k, t);
    else if (k == 4) atomicAdd(p.scal + 19, t);
    else if (k - 5 < p.A) atomicAdd(p.scal + 4 + (k - 5), t);
  }
}

// KL-adaptive learning rate (ppo.py:139-151), entirely on the device: lr_state = {lr}
__global__ void kl_lr_kernel(const float* __restrict__ scal, float count, float desired_kl, float* __restrict__ lr_state, float* __restrict__ log_out,
                             float count_a, float count_b) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float kl_mean = scal[0] / count;
  float lr = lr_state[0];
  if (desired_kl > 0.0f) {
    if (kl_mean > desired_kl * 2.0f) lr = fmaxf(1e-5f, lr / 1.5f);
    else if (kl_mean < desired_kl / 2.0f && kl_mean > 0.0f) lr = fminf(1e-2f, lr * 1.5f);
  }
  lr_state[0] = lr;
  if (log_out) { log_out[0] += scal[2] / count; log_out[1] += scal[1] / count_a + scal[19] / count_b; log_out[2] = kl_mean; log_out[3] = lr; log_out[4] += scal[3] / count; }
}

// ================================================================================================ clip + Adam
// sum of squares, deterministic two-stage
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long n, float* __restrict__ part) {
  __shared__ float red[256];
  float s = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) s += g[i] * g[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ part, int nb, float* __restrict__ out) {
  __shared__ float red[256];
  float s = 0;
  for (int i = threadIdx.x; i < nb; i += 256) s += part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) out[0] = red[0];
}
// optimiser step counter and bias corrections live on the device (lr_state = {lr, step, 1-beta1^t, sqrt(1-beta2^t)}) so that a
// whole update() can be replayed as a CUDA graph
__global__ void adam_prep_kernel(float* __restrict__ lr_state, float beta1, float beta2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float t = lr_state[1] + 1.0f;
  lr_state[1] = t;
  lr_state[2] = (float)(1.0 - pow((double)beta1, (double)t));
  lr_state[3] = (float)sqrt(1.0 - pow((double)beta2, (double)t));
}
// clip_grad_norm_(max_norm) fused with Adam (torch defaults: beta 0.9/0.999, eps 1e-8, no weight decay)
__global__ void adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
                                 const float* __restrict__ sumsq, float max_norm, const float* __restrict__ lr_state, float beta1, float beta2,
                                 float eps, float grad_scale) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float bc1 = lr_state[2], bc2_sqrt = lr_state[3];
  const float total = sqrtf(sumsq[0]) * grad_scale;
  float coef = max_norm / (total + 1e-6f);
  coef = coef < 1.0f ? coef : 1.0f;
  const float gi = g[i] * grad_scale * coef;
  const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = p[i] - (lr_state[0] / bc1) * (mi / denom);
}

}  // namespace go2

// ================================================================================================ C ABI
using namespace go2;

static int launch_gemm(int a_k_contig, int b_j_contig, const float* A, long sai, long sak, const float* B, long sbk, long sbj, float* C, long ldc,
                       int M, int N, int K, int splits, const float* bias, int epi, const float* aux, long ldaux, long split_stride, cudaStream_t st,
                       float* Ct = nullptr, long ldct = 0) {
  dim3 grid((N + 63) / 64, (M + 63) / 64, splits), block(256);
  const int k_chunk = ((K + splits - 1) / splits + 15) / 16 * 16;
  if (a_k_contig && !b_j_contig) gemm_simt_kernel<true, false><<<grid, block, 0, st>>>(A, sai, sak, B, sbk, sbj, C, ldc, M, N, K, k_chunk, bias, epi, aux, ldaux, split_stride, Ct, ldct);
  else if (a_k_contig && b_j_contig) gemm_simt_kernel<true, true><<<grid, block, 0, st>>>(A, sai, sak, B, sbk, sbj, C, ldc, M, N, K, k_chunk, bias, epi, aux, ldaux, split_stride, Ct, ldct);
  else if (!a_k_contig && b_j_contig) gemm_simt_kernel<false, true><<<grid, block, 0, st>>>(A, sai, sak, B, sbk, sbj, C, ldc, M, N, K, k_chunk, bias, epi, aux, ldaux, split_stride, Ct, ldct);
  else return set_error(3, "gemm layout not instantiated");
  count_launch();
  return 0;
}

extern "C" {

}  // extern "C"
namespace go2 {
// dX[m][k] = dy[m] * w[k] * ELU'(act[m][k]) (+ transposed copy): backward of a 1-wide Linear, 32 x 32 tiles
__global__ void dgrad_rank1_kernel(const float* __restrict__ dY, long lddy, const float* __restrict__ w, const float* __restrict__ act, long ldact,
                                   float* __restrict__ dX, long lddx, float* __restrict__ dXt, long lddxt, int M, int K) {
  __shared__ float t[32][33];
  const int k = blockIdx.x * 32 + threadIdx.x, m0 = blockIdx.y * 32;
  const float wk = k < K ? __ldg(w + k) : 0.0f;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int m = m0 + threadIdx.y + i;
    float v = 0.0f;
    if (m < M && k < K) {
      v = __ldg(dY + (long)m * lddy) * wk;
      if (act) { const float y = act[(long)m * ldact + k]; v *= (y > 0.0f ? 1.0f : y + 1.0f); }
      dX[(long)m * lddx + k] = v;
    }
    t[threadIdx.y + i][threadIdx.x] = v;
  }
  if (!dXt) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int kk = blockIdx.x * 32 + threadIdx.y + i, m = m0 + threadIdx.x;
    if (kk < K && m < M) dXt[(long)kk * lddxt + m] = t[threadIdx.x][threadIdx.y + i];
  }
}
// ---- Linear layers with a narrow output (N <= 16, K <= 128): the actor's 12-wide head (actor_critic.py:66).  A 128-wide tensor-core
// tile would be 90 % padding there; these are streaming fp32 kernels, one warp per row, lane l owning the columns l, l + 32, ...
constexpr int SN_MAXN = 16, SN_KJ = 4;
// Y[m][n] = sum_k X[m][k] W[n][k] + b[n]
__global__ void __launch_bounds__(256) linear_fwd_smalln_kernel(const float* __restrict__ X, long ldx, const float* __restrict__ W, long ldw,
                                                               const float* __restrict__ b, float* __restrict__ Y, long ldy, int M, int N, int K) {
  const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  float w[SN_MAXN][SN_KJ];
#pragma unroll
  for (int n = 0; n < SN_MAXN; ++n)
#pragma unroll
    for (int j = 0; j < SN_KJ; ++j) { const int k = lane + 32 * j; w[n][j] = (n < N && k < K) ? __ldg(W + (long)n * ldw + k) : 0.0f; }
  const float bias = (lane < N && b) ? __ldg(b + lane) : 0.0f;
  float xn[SN_KJ];
#pragma unroll
  for (int j = 0; j < SN_KJ; ++j) { const int k = lane + 32 * j; xn[j] = (gw < M && k < K) ? __ldg(X + (long)gw * ldx + k) : 0.0f; }
  for (int m = gw; m < M; m += nw) {
    float x[SN_KJ];
#pragma unroll
    for (int j = 0; j < SN_KJ; ++j) x[j] = xn[j];
#pragma unroll
    for (int j = 0; j < SN_KJ; ++j) { const int k = lane + 32 * j; xn[j] = (m + nw < M && k < K) ? __ldg(X + (long)(m + nw) * ldx + k) : 0.0f; }   // next row in flight
    float mine = 0.0f;
#pragma unroll
    for (int n = 0; n < SN_MAXN; ++n) {
      if (n < N) {
        float p = 0.0f;
#pragma unroll
        for (int j = 0; j < SN_KJ; ++j) p = fmaf(x[j], w[n][j], p);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        if (lane == n) mine = p;
      }
    }
    if (lane < N) Y[(long)m * ldy + lane] = mine + bias;
  }
}
// dX[m][k] = (sum_n dY[m][n] W[n][k]) * ELU'(act[m][k])
__global__ void __launch_bounds__(256) linear_dgrad_smalln_kernel(const float* __restrict__ dY, long lddy, const float* __restrict__ W, long ldw,
                                                                 const float* __restrict__ act, long ldact, float* __restrict__ dX, long lddx,
                                                                 int M, int N, int K) {
  const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  float w[SN_MAXN][SN_KJ];
#pragma unroll
  for (int n = 0; n < SN_MAXN; ++n)
#pragma unroll
    for (int j = 0; j < SN_KJ; ++j) { const int k = lane + 32 * j; w[n][j] = (n < N && k < K) ? __ldg(W + (long)n * ldw + k) : 0.0f; }
  float dyn = (gw < M && lane < N) ? __ldg(dY + (long)gw * lddy + lane) : 0.0f;
  for (int m = gw; m < M; m += nw) {
    const float dyl = dyn;
    dyn = (m + nw < M && lane < N) ? __ldg(dY + (long)(m + nw) * lddy + lane) : 0.0f;                 // next row in flight
    float yact[SN_KJ];
#pragma unroll
    for (int j = 0; j < SN_KJ; ++j) { const int k = lane + 32 * j; yact[j] = (act && k < K) ? __ldg(act + (long)m * ldact + k) : 1.0f; }
    float acc[SN_KJ] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int n = 0; n < SN_MAXN; ++n) {
      if (n < N) {
        const float dy = __shfl_sync(0xffffffffu, dyl, n);
#pragma unroll
        for (int j = 0; j < SN_KJ; ++j) acc[j] = fmaf(dy, w[n][j], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < SN_KJ; ++j) {
      const int k = lane + 32 * j;
      if (k < K) {
        float v = acc[j];
        if (act) { const float y = yact[j]; v *= (y > 0.0f ? 1.0f : y + 1.0f); }
        dX[(long)m * lddx + k] = v;
      }
    }
  }
}
// partial[c][n][k] = sum over CTA c's rows of dY[m][n] X[m][k];  partial[c][N*K + n] = sum of dY[m][n]   (4 warps per CTA)
__global__ void __launch_bounds__(128) linear_wgrad_smalln_partial_kernel(const float* __restrict__ dY, long lddy, const float* __restrict__ X, long ldx,
                                                                         float* __restrict__ partial, int M, int N, int K) {
  __shared__ float red[4][SN_MAXN * 32 * SN_KJ / 4 + 4];      // one 32-column slab at a time: [warp][n][lane] (+ bias sums)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = (M + gridDim.x - 1) / gridDim.x, m0 = blockIdx.x * rows, m1 = min(M, m0 + rows);
  float acc[SN_MAXN][SN_KJ], sdy = 0.0f;
#pragma unroll
  for (int n = 0; n < SN_MAXN; ++n)
#pragma unroll
    for (int j = 0; j < SN_KJ; ++j) acc[n][j] = 0.0f;
  float xn[SN_KJ], dyn = (m0 + warp < m1 && lane < N) ? __ldg(dY + (long)(m0 + warp) * lddy + lane) : 0.0f;
#pragma unroll
  for (int j = 0; j < SN_KJ; ++j) { const int k = lane + 32 * j; xn[j] = (m0 + warp < m1 && k < K) ? __ldg(X + (long)(m0 + warp) * ldx + k) : 0.0f; }
  for (int m = m0 + warp; m < m1; m += 4) {
    const float dyl = dyn;
    sdy += dyl;
    float x[SN_KJ];
#pragma unroll
    for (int j = 0; j < SN_KJ; ++j) x[j] = xn[j];
    dyn = (m + 4 < m1 && lane < N) ? __ldg(dY + (long)(m + 4) * lddy + lane) : 0.0f;                  // next row in flight
#pragma unroll
    for (int j = 0; j < SN_KJ; ++j) { const int k = lane + 32 * j; xn[j] = (m + 4 < m1 && k < K) ? __ldg(X + (long)(m + 4) * ldx + k) : 0.0f; }
#pragma unroll
    for (int n = 0; n < SN_MAXN; ++n) {
      if (n < N) {
        const float dy = __shfl_sync(0xffffffffu, dyl, n);
#pragma unroll
        for (int j = 0; j < SN_KJ; ++j) acc[n][j] = fmaf(dy, x[j], acc[n][j]);
      }
    }
  }
  float* out = partial + (long)blockIdx.x * (N * K + N);
#pragma unroll
  for (int j = 0; j < SN_KJ; ++j) {              // cross-warp reduction, one 32-column slab per pass, fixed order
#pragma unroll
    for (int n = 0; n < SN_MAXN; ++n) red[warp][n * 32 + lane] = acc[n][j];
    __syncthreads();
    for (int t = threadIdx.x; t < N * 32; t += 128) {
      const int n = t >> 5, k = (t & 31) + 32 * j;
      if (k < K) out[n * K + k] = (red[0][t] + red[1][t]) + (red[2][t] + red[3][t]);
    }
    __syncthreads();
  }
  red[warp][lane] = sdy;
  __syncthreads();
  if (threadIdx.x < N) out[N * K + threadIdx.x] = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}
// out[t] = sum over the n_part partial blocks, t in [0, N K + N): warp w of a CTA sums blocks w, w + 8, ... (independent coalesced loads in
// flight), then the 8 warp sums are added in a fixed tree.  t < N K lands in dW (row pitch lddw), the rest in db.
__global__ void __launch_bounds__(256) wgrad_partials_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, long lddw,
                                                                   float* __restrict__ db, int N, int K, int n_part) {
  __shared__ float red[8][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, per = N * K + N;
  const int t = blockIdx.x * 32 + lane;
  float s = 0.0f;
  if (t < per) {
#pragma unroll 8
    for (int c = warp; c < n_part; c += 8) s += __ldg(partial + (long)c * per + t);
  }
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && t < per) {
    const float r = ((red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane])) + ((red[4][lane] + red[5][lane]) + (red[6][lane] + red[7][lane]));
    if (t < N * K) dW[(long)(t / K) * lddw + (t % K)] = r;
    else if (db) db[t - N * K] = r;
  }
}

// dW[k] = sum_m dy[m] X[m][k], db = sum_m dy[m]: weight gradient of a 1-wide Linear (the critic's head).  One streaming pass over X:
// CTA c takes a contiguous block of rows, warp w its rows w, w+8, ..., lane l the columns l, l+32, ... (coalesced row reads);
// partial[c][0..K-1] = column sums, partial[c][K] = sum of dy.  Fixed summation order everywhere (deterministic).
constexpr int WG1_CTAS = 296, WG1_MAXK = 512;
__global__ void __launch_bounds__(256) wgrad_rank1_partial_kernel(const float* __restrict__ dY, long lddy, const float* __restrict__ X, long ldx,
                                                                 float* __restrict__ partial, int M, int K) {
  __shared__ float red[8][WG1_MAXK + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = (M + gridDim.x - 1) / gridDim.x, m0 = blockIdx.x * rows, m1 = min(M, m0 + rows);
  float acc[WG1_MAXK / 32], sdy = 0.0f;
#pragma unroll
  for (int j = 0; j < WG1_MAXK / 32; ++j) acc[j] = 0.0f;
  for (int m = m0 + warp; m < m1; m += 8) {
    const float dy = __ldg(dY + (long)m * lddy);
    sdy += dy;
    const float* x = X + (long)m * ldx;
#pragma unroll
    for (int j = 0; j < WG1_MAXK / 32; ++j) { const int k = lane + 32 * j; if (k < K) acc[j] = fmaf(dy, __ldg(x + k), acc[j]); }
  }
#pragma unroll
  for (int j = 0; j < WG1_MAXK / 32; ++j) { const int k = lane + 32 * j; if (k < K) red[warp][k] = acc[j]; }
  if (lane == 0) red[warp][K] = sdy;
  __syncthreads();
  for (int k = threadIdx.x; k <= K; k += 256) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][k];
    partial[(long)blockIdx.x * (K + 1) + k] = t;
  }
}
}  // namespace go2
extern "C" {
int go2_linear_forward_smalln(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy, int M, int N, int K, void* stream) {
  if (!X || !W || !Y) return set_error(1, "go2_linear_forward_smalln: null argument");
  if (N > SN_MAXN || K > 32 * SN_KJ) return set_error(1, "go2_linear_forward_smalln: needs N <= 16 and K <= 128");
  const int ctas = max(1, min(148 * 2, (M + 7) / 8));
  linear_fwd_smalln_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(X, ldx, W, ldw, b, Y, ldy, M, N, K);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_linear_dgrad_smalln(const float* dY, int lddy, const float* W, int ldw, const float* act_in, int ldact, float* dX, int lddx, int M, int N, int K,
                            void* stream) {
  if (!dY || !W || !dX) return set_error(1, "go2_linear_dgrad_smalln: null argument");
  if (N > SN_MAXN || K > 32 * SN_KJ) return set_error(1, "go2_linear_dgrad_smalln: needs N <= 16 and K <= 128");
  const int ctas = max(1, min(148 * 2, (M + 7) / 8));
  linear_dgrad_smalln_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(dY, lddy, W, ldw, act_in, ldact, dX, lddx, M, N, K);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_linear_wgrad_smalln(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, float* db, int M, int N, int K, float* workspace,
                            long workspace_floats, void* stream) {
  if (!dY || !X || !dW || !workspace) return set_error(1, "go2_linear_wgrad_smalln: null argument");
  if (N > SN_MAXN || K > 32 * SN_KJ) return set_error(1, "go2_linear_wgrad_smalln: needs N <= 16 and K <= 128");
  const long per = (long)N * K + N;
  const int ctas = (int)min((long)296, min((long)((M + 3) / 4), workspace_floats / per));
  if (ctas < 1) return set_error(1, "go2_linear_wgrad_smalln: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  linear_wgrad_smalln_partial_kernel<<<ctas, 128, 0, st>>>(dY, lddy, X, ldx, workspace, M, N, K);
  count_launch();
  wgrad_partials_reduce_kernel<<<(unsigned)((per + 31) / 32), 256, 0, st>>>(workspace, dW, lddw, db, N, K, ctas);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_linear_wgrad_rank1(const float* dY, int lddy, const float* X, int ldx, float* dW, float* db, int M, int K, float* workspace,
                           long workspace_floats, void* stream) {
  if (!dY || !X || !dW || !workspace) return set_error(1, "go2_linear_wgrad_rank1: null argument");
  if (K > WG1_MAXK) return set_error(1, "go2_linear_wgrad_rank1: K > 512");
  const int ctas = (int)min((long)WG1_CTAS, min((long)((M + 7) / 8), workspace_floats / (K + 1)));
  if (ctas < 1) return set_error(1, "go2_linear_wgrad_rank1: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  wgrad_rank1_partial_kernel<<<ctas, 256, 0, st>>>(dY, lddy, X, ldx, workspace, M, K);
  count_launch();
  wgrad_partials_reduce_kernel<<<(K + 1 + 31) / 32, 256, 0, st>>>(workspace, dW, K, db, 1, K, ctas);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_linear_forward_simt(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy, float* Yt, int ldyt, int M, int N, int K, int act, void* stream) {
  int rc = launch_gemm(1, 0, X, ldx, 1, W, 1, ldw, Y, ldy, M, N, K, 1, b, act ? EPI_BIAS_ELU : EPI_BIAS, nullptr, 0, 0, (cudaStream_t)stream, Yt, ldyt);
  if (rc) return rc;
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_linear_dgrad_simt(const float* dY, int lddy, const float* W, int ldw, const float* act_in, int ldact, float* dX, int lddx, float* dXt, int lddxt, int M, int N, int K, void* stream) {
  // dX[M,K] = dY[M,N] W[N,K], times ELU'(act_in) when act_in != NULL
  if (N == 1) {   // the critic's scalar head: an outer product, one streaming pass
    dim3 block(32, 8), grid((K + 31) / 32, (M + 31) / 32);
    dgrad_rank1_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(dY, lddy, W, act_in, ldact, dX, lddx, dXt, lddxt, M, K);
    count_launch();
    GO2_CUDA_OK(cudaGetLastError());
    return 0;
  }
  int rc = launch_gemm(1, 1, dY, lddy, 1, W, ldw, 1, dX, lddx, M, K, N, 1, nullptr, act_in ? EPI_MUL_ELU_GRAD : EPI_NONE, act_in, ldact, 0, (cudaStream_t)stream, dXt, lddxt);
  if (rc) return rc;
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_linear_wgrad_simt(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, float* db, int M, int N, int K, float* workspace,
                          long workspace_floats, void* stream) {
  // dW[N,K] = dY^T X (contraction over the M rows), db[N] = column sums of dY.  Deterministic split over M.
  cudaStream_t st = (cudaStream_t)stream;
  int tiles = ((N + 63) / 64) * ((K + 63) / 64);
  int splits = max(1, min(64, (2 * 148 + tiles - 1) / tiles));
  while (splits > 1 && (long)splits * N * K > workspace_floats) --splits;
  if (splits > 1 && !workspace) splits = 1;
  if (splits == 1) {
    int rc = launch_gemm(0, 1, dY, 1, lddy, X, ldx, 1, dW, lddw, N, K, M, 1, nullptr, EPI_NONE, nullptr, 0, 0, st);
    if (rc) return rc;
  } else {
    int rc = launch_gemm(0, 1, dY, 1, lddy, X, ldx, 1, workspace, K, N, K, M, splits, nullptr, EPI_NONE, nullptr, 0, (long)N * K, st);
    if (rc) return rc;
    long n = (long)N * K;
    splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(workspace, dW, n, splits, lddw, K);
    count_launch();
  }
  if (db) return set_error(1, "go2_linear_wgrad_simt: bias gradient moved to go2_colsum");
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_colsum(const float* dY, int lddy, float* db, int M, int N, float* scratch /* >= 64*N floats */, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = max(1, min(64, M / 256));
  const int rpc = (M + chunks - 1) / chunks;
  dim3 grid((N + 31) / 32, chunks);
  colsum_partial_kernel<<<grid, 256, 0, st>>>(dY, lddy, scratch, M, N, rpc);
  count_launch();
  colsum_final_kernel<<<(N + 127) / 128, 128, 0, st>>>(scratch, db, N, chunks);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_sample_actions(const float* mu, const float* std_param, float* actions, float* logp, float* mu_out, float* sigma_out, int N, int A,
                       uint64_t seed, uint32_t step, int env_offset, void* stream) {
  sample_actions_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mu, std_param, actions, logp, mu_out, sigma_out, N, A, (uint32_t)seed,
                                                                         (uint32_t)(seed >> 32), step, nullptr, env_offset);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_sample_actions_dev(const float* mu, const float* std_param, float* actions, float* logp, float* mu_out, float* sigma_out, int N, int A,
                           uint64_t seed, const uint32_t* d_step, int env_offset, void* stream) {
  if (!d_step) return set_error(1, "go2_sample_actions_dev: null step pointer");
  sample_actions_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mu, std_param, actions, logp, mu_out, sigma_out, N, A, (uint32_t)seed,
                                                                         (uint32_t)(seed >> 32), 0u, d_step, env_offset);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_process_env_step(const float* rew, const uint8_t* dones, const uint8_t* time_outs, const float* values, float* rew_out, uint8_t* dones_out,
                         int N, float gamma, const int64_t* perm, void* stream) {
  process_env_step_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rew, dones, time_outs, values, rew_out, dones_out, N, gamma, perm);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values, float* returns, float* advantages, int T, int N,
            float gamma, float lam, double* stats, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GO2_CUDA_OK(cudaMemsetAsync(stats, 0, 2 * sizeof(double), st));
  gae_scan_kernel<<<(N + 255) / 256, 256, 0, st>>>(rewards, values, dones, last_values, returns, advantages, T, N, gamma, lam, stats);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_adv_normalize(float* advantages, long n, const double* stats, double global_count, void* stream) {
  adv_normalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(advantages, n, stats, global_count);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_gather_rows(const float* src, int width, const int64_t* idx, float* dst, int ldd, float* dst_t, long n, void* stream) {
  dim3 block(32, 8);
  gather_rows_kernel<<<(unsigned)((n + 7) / 8), block, 0, (cudaStream_t)stream>>>(src, width, idx, dst, ldd, n, dst_t);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_ppo_loss(const float* mu, const float* std_param, const float* value, const float* actions, const float* old_logp, const float* adv,
                 const float* target_values, const float* returns, const float* old_mu, const float* old_sigma, float* dmu, float* dmu_t, float* dvalue,
                 float* scal, int M, int A, float clip, float value_coef, float entropy_coef, int use_clipped_value_loss, float inv_count,
                 int split, float inv_count_a, float inv_count_b, void* stream) {
  if (A > 16) return set_error(1, "go2_ppo_loss: at most 16 actions");
  cudaStream_t st = (cudaStream_t)stream;
  GO2_CUDA_OK(cudaMemsetAsync(scal, 0, sizeof(float) * (4 + 16), st));
  PpoLossArgs p{mu, std_param, value, actions, old_logp, adv, target_values, returns, old_mu, old_sigma, dmu, dmu_t, dvalue, scal,
                M, A, clip, value_coef, entropy_coef, use_clipped_value_loss, inv_count, split, inv_count_a, inv_count_b};
  ppo_loss_kernel<<<(M + 255) / 256, 256, 0, st>>>(p);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_kl_adaptive_lr(const float* scal, float count, float desired_kl, float* lr_state, float* log_out, float count_a, float count_b, void* stream) {
  kl_lr_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scal, count, desired_kl, lr_state, log_out, count_a, count_b);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_adam_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n, float max_grad_norm, float* lr_state,
                       float grad_scale, float* scratch /* >= 1025 floats */, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (int)min((long)1024, (n + 255) / 256);
  sumsq_partial_kernel<<<nb, 256, 0, st>>>(grads, n, scratch + 1);
  count_launch();
  sumsq_final_kernel<<<1, 256, 0, st>>>(scratch + 1, nb, scratch);
  count_launch();
  adam_prep_kernel<<<1, 32, 0, st>>>(lr_state, 0.9f, 0.999f);
  count_launch();
  adam_clip_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, scratch, max_grad_norm, lr_state, 0.9f, 0.999f,
                                                                1e-8f, grad_scale);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
