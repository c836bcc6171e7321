// cts_kernels.cu — the small kernels the Concurrent Teacher-Student (MoE) trainer needs around the GEMMs.
//
// Reference semantics:
//   L2Norm                      rsl_rl/modules/utils.py:24-30   (F.normalize(x, p=2, dim=-1), eps 1e-12)
//   MoE combine + softmax gate  rsl_rl/modules/utils.py:96-126
//   latent / load-balance loss  rsl_rl/algorithms/moe_cts.py:197-216, cts.py (latent MSE)
//   [latent | obs] concat       rsl_rl/modules/actor_critic_moe_cts.py:114-141
//   history roll                rsl_rl/runners/on_policy_runner_cts.py:155-156
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/go2_b200.h"
#include "common.cuh"

namespace go2 {

// out[i, 0:wa] = a[i], out[i, wa:wa+wb] = b[i], zero padded to ld;  out_t (optional) = transposed copy [wa+wb(+ones row kept), n]
__global__ void concat2_kernel(const float* __restrict__ a, int wa, long lda, const float* __restrict__ b, int wb, long ldb, float* __restrict__ out, int ld,
                               float* __restrict__ out_t, long n) {
  __shared__ float tile[8][33];
  const long i = (long)blockIdx.x * blockDim.y + threadIdx.y;
  const int w = wa + wb;
  for (int c0 = 0; c0 < ld; c0 += 32) {
    const int c = c0 + threadIdx.x;
    float v = (c == w) ? 1.0f : 0.0f;                  // padding: column w = 1 (bias-gradient "ones" column), the rest 0
    if (i < n && c < w) v = c < wa ? a[i * lda + c] : b[i * ldb + (c - wa)];
    if (i < n && c < ld) out[i * ld + c] = v;
    if (out_t) {
      tile[threadIdx.y][threadIdx.x] = v;
      __syncthreads();
      const int tr = threadIdx.x % 8, tc = threadIdx.y * 4 + threadIdx.x / 8;
      const long ii = (long)blockIdx.x * blockDim.y + tr;
      const int cc = c0 + tc;
      if (ii < n && cc < w) out_t[(long)cc * n + ii] = tile[tr][tc];
      __syncthreads();
    }
  }
}

// y = x / max(||x||, 1e-12); also keeps the norm for the backward
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, long ldx, float* __restrict__ y, long ldy, float* __restrict__ norm, long n, int d) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0;
  for (int k = 0; k < d; ++k) { float v = x[i * ldx + k]; s += v * v; }
  const float nr = fmaxf(sqrtf(s), 1e-12f);
  if (norm) norm[i] = nr;
  const float inv = 1.0f / nr;
  for (int k = 0; k < d; ++k) y[i * ldy + k] = x[i * ldx + k] * inv;
}
// dx = (dy - y (y . dy)) / ||x||   (clamped-norm rows pass dy / 1e-12 like autograd; never hit in practice)
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, long lddy, const float* __restrict__ y, long ldy, const float* __restrict__ norm,
                                  float* __restrict__ dx, long lddx, float* __restrict__ dx_t, long n, int d) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float dotp = 0;
  for (int k = 0; k < d; ++k) dotp += y[i * ldy + k] * dy[i * lddy + k];
  const float inv = 1.0f / norm[i];
  for (int k = 0; k < d; ++k) {
    const float g = (dy[i * lddy + k] - y[i * ldy + k] * dotp) * inv;
    dx[i * lddx + k] = g;
    if (dx_t) dx_t[(long)k * n + i] = g;
  }
}

// gates = softmax(logits) ; pre[b, o] = sum_e gates[b,e] * expert_out[b, e*D + o]
__global__ void moe_combine_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ eo, float* __restrict__ gates, float* __restrict__ pre,
                                       long n, int E, int D) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g[16];
  float mx = -1e30f;
  for (int e = 0; e < E; ++e) { g[e] = logits[i * E + e]; mx = fmaxf(mx, g[e]); }
  float s = 0;
  for (int e = 0; e < E; ++e) { g[e] = expf(g[e] - mx); s += g[e]; }
  for (int e = 0; e < E; ++e) { g[e] /= s; gates[i * E + e] = g[e]; }
  for (int o = 0; o < D; ++o) {
    float acc = 0;
    for (int e = 0; e < E; ++e) acc += g[e] * eo[i * (long)(E * D) + e * D + o];
    pre[i * D + o] = acc;
  }
}
// column means of the gates (mean usage, moe_cts.py:211) : one block per expert
__global__ void __launch_bounds__(256) gate_usage_kernel(const float* __restrict__ gates, float* __restrict__ usage, long n, int E, float scale = 1.0f) {
  __shared__ float red[256];
  const int e = blockIdx.x;
  float s = 0;
  for (long i = threadIdx.x; i < n; i += 256) s += gates[i * E + e];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) usage[e] = red[0] / (float)n * scale;
}
// backward of the combine + softmax + load-balance term.  dpre [n,D] -> deo [n,E*D] (+ transposed), dlogits [n,E] (+ transposed)
__global__ void moe_combine_bwd_kernel(const float* __restrict__ dpre, const float* __restrict__ gates, const float* __restrict__ eo,
                                       const float* __restrict__ usage, float lb_coef, float* __restrict__ deo, float* __restrict__ deo_t,
                                       float* __restrict__ dlogits, float* __restrict__ dlogits_t, long n, int E, int D) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g[16], dg[16];
  for (int e = 0; e < E; ++e) {
    g[e] = gates[i * E + e];
    // d/dg of lb_coef * mean_e (usage_e - 1/E)^2 with usage_e = mean_b g[b,e]
    dg[e] = lb_coef * (2.0f / (float)E) * (usage[e] - 1.0f / (float)E) / (float)n;
  }
  for (int o = 0; o < D; ++o) {
    const float d = dpre[i * D + o];
    for (int e = 0; e < E; ++e) {
      const long c = (long)e * D + o;
      dg[e] += d * eo[i * (long)(E * D) + c];
      const float v = g[e] * d;
      deo[i * (long)(E * D) + c] = v;
      if (deo_t) deo_t[c * n + i] = v;
    }
  }
  float dot = 0;
  for (int e = 0; e < E; ++e) dot += g[e] * dg[e];
  for (int e = 0; e < E; ++e) {
    const float v = g[e] * (dg[e] - dot);
    dlogits[i * E + e] = v;
    if (dlogits_t) dlogits_t[(long)e * n + i] = v;
  }
}

// latent loss = mean((t - s)^2) over n*d ; ds = 2 (s - t) / (n d) ; acc[0] += sum of squares
__global__ void __launch_bounds__(256) latent_loss_kernel(const float* __restrict__ s, const float* __restrict__ t, float* __restrict__ ds, float* __restrict__ acc,
                                                          long n, int d) {
  __shared__ float red[256];
  const long total = n * d;
  float loc = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const float df = s[i] - t[i];
    loc += df * df;
    ds[i] = 2.0f * df / (float)total;
  }
  red[threadIdx.x] = loc;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) atomicAdd(acc, red[0]);
}
// log[0] += latent loss, log[1] += load-balance loss (device-side logging, no host sync)
__global__ void cts_log_kernel(const float* __restrict__ acc, const float* __restrict__ usage, float* __restrict__ log, long count, int E) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  log[0] += acc[0] / (float)count;
  float lb = 0;
  if (usage) { for (int e = 0; e < E; ++e) { float d = usage[e] - 1.0f / (float)E; lb += d * d; } lb /= (float)E; }
  log[1] += lb;
}

// history[n, 0:H-1] <- history[n, 1:H] (zeroed first when done), history[n, H-1] <- obs[n]   (on_policy_runner_cts.py:155-156)
__global__ void history_update_kernel(float* __restrict__ hist, const float* __restrict__ obs, const uint8_t* __restrict__ dones, long n, int H, int d) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * d) return;
  const long e = idx / d;
  const int j = (int)(idx % d);
  const bool done = dones && dones[e];
  float* h = hist + e * (long)(H * d);
  for (int k = 0; k < H - 1; ++k) h[k * d + j] = done ? 0.0f : h[(k + 1) * d + j];
  h[(H - 1) * d + j] = obs[e * d + j];
}

// out[i] = src[perm[i]] for byte rows / float scalars (teacher-first reordering of the CTS transition)
__global__ void gather_u8_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ perm, uint8_t* __restrict__ out, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[perm[i]];
}

// ================================================================================================ grouped expert layer (fp32, CUDA cores)
// Conv1d(E*H -> E*D, kernel 1, groups = E) (rsl_rl/modules/utils.py:83-88) = E block-diagonal Linear(H -> D) over the feature blocks of the backbone.
// Each expert is a [M, H] x [H, D] product with D ~ 32: eight tensor-core launches of a few tiles each were pure launch latency (12 us apiece in the
// rollout).  ONE launch per direction for all experts, exact fp32 FMAs (sequential over the contraction index inside a thread):
//   forward  Y[m, e D + d]  = b[e D + d] + sum_h X[m, e H + h] W[e D + d, h]
//   dgrad    dX[m, e H + h] = ELU'(act[m, e H + h]) sum_d dY[m, e D + d] W[e D + d, h]          (act = the backbone's last activation, may be null)
//   wgrad    dW[e D + d, h] = sum_m dY[m, e D + d] X[m, e H + h]     partials over row chunks, summed in chunk order by a second kernel (deterministic)
constexpr int GX_BM = 64, GX_BK = 32, GX_BD = 32;

__global__ void __launch_bounds__(256) grouped_fwd_kernel(const float* __restrict__ X, long ldx, const float* __restrict__ W, const float* __restrict__ b,
                                                          float* __restrict__ Y, long ldy, long M, int E, int D, int H) {
  __shared__ float Xs[GX_BM][GX_BK + 1], Ws[GX_BD][GX_BK + 1];
  const int e = blockIdx.y / ((D + GX_BD - 1) / GX_BD), d0 = (blockIdx.y % ((D + GX_BD - 1) / GX_BD)) * GX_BD;
  const long m0 = (long)blockIdx.x * GX_BM;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // thread -> rows ty * 4 .. + 3, columns tx * 2, tx * 2 + 1
  float acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  for (int k0 = 0; k0 < H; k0 += GX_BK) {
    for (int i = threadIdx.x; i < GX_BM * GX_BK; i += 256) {
      const int r = i / GX_BK, k = i % GX_BK;
      Xs[r][k] = (m0 + r < M && k0 + k < H) ? X[(m0 + r) * ldx + (long)e * H + k0 + k] : 0.0f;
    }
    for (int i = threadIdx.x; i < GX_BD * GX_BK; i += 256) {
      const int d = i / GX_BK, k = i % GX_BK;
      Ws[d][k] = (d0 + d < D && k0 + k < H) ? W[((long)e * D + d0 + d) * H + k0 + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < GX_BK; ++k) {
      const float w0 = Ws[tx * 2][k], w1 = Ws[tx * 2 + 1][k];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float x = Xs[ty * 4 + r][k];
        acc[r][0] = fmaf(x, w0, acc[r][0]); acc[r][1] = fmaf(x, w1, acc[r][1]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const long m = m0 + ty * 4 + r; const int d = d0 + tx * 2 + c;
      if (m < M && d < D) Y[m * ldy + (long)e * D + d] = acc[r][c] + b[e * D + d];
    }
}

// dX tile: 64 rows x 64 feature columns of one expert; contraction over D in chunks of 32
__global__ void __launch_bounds__(256) grouped_dgrad_kernel(const float* __restrict__ dY, long lddy, const float* __restrict__ W, const float* __restrict__ act,
                                                            long ldact, float* __restrict__ dX, long lddx, long M, int E, int D, int H) {
  __shared__ float Gs[GX_BM][GX_BD + 1], Ws[GX_BD][64 + 1];
  const int hb = (H + 63) / 64, e = blockIdx.y / hb, h0 = (blockIdx.y % hb) * 64;
  const long m0 = (long)blockIdx.x * GX_BM;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // rows ty * 4 .. + 3, columns tx + 16 c (c = 0..3)
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
  for (int d0 = 0; d0 < D; d0 += GX_BD) {
    for (int i = threadIdx.x; i < GX_BM * GX_BD; i += 256) {
      const int r = i / GX_BD, d = i % GX_BD;
      Gs[r][d] = (m0 + r < M && d0 + d < D) ? dY[(m0 + r) * lddy + (long)e * D + d0 + d] : 0.0f;
    }
    for (int i = threadIdx.x; i < GX_BD * 64; i += 256) {
      const int d = i / 64, h = i % 64;
      Ws[d][h] = (d0 + d < D && h0 + h < H) ? W[((long)e * D + d0 + d) * H + h0 + h] : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int d = 0; d < GX_BD; ++d) {
      float w[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) w[c] = Ws[d][tx + 16 * c];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float g = Gs[ty * 4 + r][d];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(g, w[c], acc[r][c]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const long m = m0 + ty * 4 + r; const int h = h0 + tx + 16 * c;
      if (m < M && h < H) {
        float v = acc[r][c];
        if (act) { const float y = act[m * ldact + (long)e * H + h]; v *= (y > 0.0f ? 1.0f : y + 1.0f); }
        dX[m * lddx + (long)e * H + h] = v;
      }
    }
}

// partial weight gradients: block (chunk, e, h-block of 64): part[chunk][e D + d][h] = sum over the chunk's rows; D <= 32 per pass (grid z = d-block)
constexpr int GX_ROWS = 256;     // rows per chunk
__global__ void __launch_bounds__(256) grouped_wgrad_partial_kernel(const float* __restrict__ dY, long lddy, const float* __restrict__ X, long ldx,
                                                                    float* __restrict__ part, long M, int E, int D, int H) {
  __shared__ float Gs[32][GX_BD + 1], Xs[32][64 + 1];
  const int hb = (H + 63) / 64, e = blockIdx.y / hb, h0 = (blockIdx.y % hb) * 64, d0 = blockIdx.z * GX_BD;
  const long r0 = (long)blockIdx.x * GX_ROWS, r1 = min(M, r0 + GX_ROWS);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // outputs d = ty * 2 + {0, 1}, h = tx + 16 c
  float acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  for (long m0 = r0; m0 < r1; m0 += 32) {
    for (int i = threadIdx.x; i < 32 * GX_BD; i += 256) {
      const int r = i / GX_BD, d = i % GX_BD;
      Gs[r][d] = (m0 + r < r1 && d0 + d < D) ? dY[(m0 + r) * lddy + (long)e * D + d0 + d] : 0.0f;
    }
    for (int i = threadIdx.x; i < 32 * 64; i += 256) {
      const int r = i / 64, h = i % 64;
      Xs[r][h] = (m0 + r < r1 && h0 + h < H) ? X[(m0 + r) * ldx + (long)e * H + h0 + h] : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float g0 = Gs[r][ty * 2], g1 = Gs[r][ty * 2 + 1];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float x = Xs[r][tx + 16 * c];
        acc[0][c] = fmaf(g0, x, acc[0][c]); acc[1][c] = fmaf(g1, x, acc[1][c]);
      }
    }
    __syncthreads();
  }
  float* P = part + (long)blockIdx.x * E * D * H;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int d = d0 + ty * 2 + j, h = h0 + tx + 16 * c;
      if (d < D && h < H) P[((long)e * D + d) * H + h] = acc[j][c];
    }
}
__global__ void grouped_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dW, int chunks, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  for (int c = 0; c < chunks; ++c) s += part[(long)c * n + i];      // chunk order: deterministic
  dW[i] = s;
}

}  // namespace go2

using namespace go2;

extern "C" {

int go2_concat2(const float* a, int wa, int lda, const float* b, int wb, int ldb, float* out, int ld, float* out_t, long n, void* stream) {
  dim3 block(32, 8);
  concat2_kernel<<<(unsigned)((n + 7) / 8), block, 0, (cudaStream_t)stream>>>(a, wa, lda, b, wb, ldb, out, ld, out_t, n);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}
int go2_l2norm_forward(const float* x, int ldx, float* y, int ldy, float* norm, long n, int d, void* stream) {
  l2norm_fwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, norm, n, d);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}
int go2_l2norm_backward(const float* dy, int lddy, const float* y, int ldy, const float* norm, float* dx, int lddx, float* dx_t, long n, int d, void* stream) {
  l2norm_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(dy, lddy, y, ldy, norm, dx, lddx, dx_t, n, d);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}
int go2_moe_combine_forward(const float* logits, const float* expert_out, float* gates, float* pre, long n, int E, int D, void* stream) {
  if (E > 16) return set_error(1, "go2_moe_combine_forward: at most 16 experts");
  moe_combine_fwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(logits, expert_out, gates, pre, n, E, D);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}
int go2_moe_combine_backward(const float* dpre, const float* gates, const float* expert_out, float* usage /* [E] scratch */, float lb_coef, float* dexpert_out,
                             float* dexpert_out_t, float* dlogits, float* dlogits_t, long n, int E, int D, void* stream) {
  if (E > 16) return set_error(1, "go2_moe_combine_backward: at most 16 experts");
  cudaStream_t st = (cudaStream_t)stream;
  gate_usage_kernel<<<E, 256, 0, st>>>(gates, usage, n, E);
  count_launch();
  moe_combine_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(dpre, gates, expert_out, usage, lb_coef, dexpert_out, dexpert_out_t, dlogits, dlogits_t, n, E, D);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}
// The two halves of go2_moe_combine_backward for the env-sharded trainer: the mean gate usage of THIS rank's rows times `scale` (1 / world_size),
// summed over the ranks by the caller (go2_allreduce_p2p2), then the backward with the GLOBAL usage — the load-balance term of moe_cts.py:211-214
// is a function of the whole mini-batch's mean usage, not of one shard's.
int go2_gate_usage(const float* gates, float* usage, long n, int E, float scale, void* stream) {
  if (E > 16) return set_error(1, "go2_gate_usage: at most 16 experts");
  gate_usage_kernel<<<E, 256, 0, (cudaStream_t)stream>>>(gates, usage, n, E, scale);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}
int go2_moe_combine_backward_given_usage(const float* dpre, const float* gates, const float* expert_out, const float* usage, float lb_coef,
                                         float* dexpert_out, float* dexpert_out_t, float* dlogits, float* dlogits_t, long n, int E, int D, void* stream) {
  if (E > 16) return set_error(1, "go2_moe_combine_backward_given_usage: at most 16 experts");
  moe_combine_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(dpre, gates, expert_out, usage, lb_coef, dexpert_out, dexpert_out_t,
                                                                                     dlogits, dlogits_t, n, E, D);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}
int go2_latent_loss(const float* student, const float* teacher, float* dstudent, float* acc /* [1] */, long n, int d, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GO2_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(float), st));
  latent_loss_kernel<<<(unsigned)min((long)296, (n * d + 255) / 256), 256, 0, st>>>(student, teacher, dstudent, acc, n, d);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}
int go2_cts_log(const float* acc, const float* usage, float* log, long count, int E, void* stream) {
  cts_log_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, usage, log, count, E);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}
int go2_history_update(float* history, const float* obs, const uint8_t* dones, long n, int H, int d, void* stream) {
  history_update_kernel<<<(unsigned)((n * d + 255) / 256), 256, 0, (cudaStream_t)stream>>>(history, obs, dones, n, H, d);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}
int go2_gather_u8(const uint8_t* src, const int64_t* perm, uint8_t* out, long n, void* stream) {
  gather_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, perm, out, n);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_grouped_linear_forward(const float* X, long ldx, const float* W, const float* b, float* Y, long ldy, long M, int E, int D, int H, void* stream) {
  if (!X || !W || !b || !Y || M <= 0 || E <= 0 || D <= 0 || H <= 0) return set_error(1, "go2_grouped_linear_forward: bad argument");
  dim3 grid((unsigned)((M + GX_BM - 1) / GX_BM), (unsigned)(E * ((D + GX_BD - 1) / GX_BD)));
  grouped_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, W, b, Y, ldy, M, E, D, H);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}
int go2_grouped_linear_dgrad(const float* dY, long lddy, const float* W, const float* act, long ldact, float* dX, long lddx, long M, int E, int D, int H,
                             void* stream) {
  if (!dY || !W || !dX || M <= 0 || E <= 0 || D <= 0 || H <= 0) return set_error(1, "go2_grouped_linear_dgrad: bad argument");
  dim3 grid((unsigned)((M + GX_BM - 1) / GX_BM), (unsigned)(E * ((H + 63) / 64)));
  grouped_dgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dY, lddy, W, act, ldact, dX, lddx, M, E, D, H);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}
long go2_grouped_linear_wgrad_workspace(long M, int E, int D, int H) { return ((M + GX_ROWS - 1) / GX_ROWS) * (long)E * D * H; }
int go2_grouped_linear_wgrad(const float* dY, long lddy, const float* X, long ldx, float* dW, long M, int E, int D, int H, float* workspace,
                             long workspace_floats, void* stream) {
  if (!dY || !X || !dW || !workspace || M <= 0 || E <= 0 || D <= 0 || H <= 0) return set_error(1, "go2_grouped_linear_wgrad: bad argument");
  const int chunks = (int)((M + GX_ROWS - 1) / GX_ROWS);
  const long n = (long)E * D * H;
  if (workspace_floats < chunks * n) return set_error(1, "go2_grouped_linear_wgrad: workspace too small (go2_grouped_linear_wgrad_workspace)");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)chunks, (unsigned)(E * ((H + 63) / 64)), (unsigned)((D + GX_BD - 1) / GX_BD));
  grouped_wgrad_partial_kernel<<<grid, 256, 0, st>>>(dY, lddy, X, ldx, workspace, M, E, D, H);
  count_launch();
  grouped_wgrad_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(workspace, dW, chunks, n);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
