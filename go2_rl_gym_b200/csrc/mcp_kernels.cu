// mcp_kernels.cu — sm_100a kernels of the MCP-CTS trainer variant (go2_mcp_cts) and their C ABI: the product-of-Gaussians actor head
// (forward / backward) and the sampling / PPO-loss kernels for a state-dependent sigma.  One thread per row; the row arithmetic lives in
// mcp_core.cuh (shared with the host emulation).  HBM-bound streaming kernels: ~(E * 2A + E + 2A) floats per row.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/go2_b200.h"
#include "common.cuh"
#include "mcp_core.cuh"

namespace go2 {

__global__ void __launch_bounds__(128) mcp_compose_fwd_kernel(const float* __restrict__ eo, const float* __restrict__ logits, float* __restrict__ gates,
                                                              float* __restrict__ mu, float* __restrict__ sigma, long n, int E, int A) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  mcp_compose_row(eo + i * (long)(E * 2 * A), logits + i * E, E, A, gates + i * E, mu + i * A, sigma + i * A);
}

__global__ void __launch_bounds__(128) mcp_compose_bwd_kernel(const float* __restrict__ dmu, const float* __restrict__ dsigma, const float* __restrict__ eo,
                                                              const float* __restrict__ gates, float* __restrict__ deo, float* __restrict__ dlogits,
                                                              long n, int E, int A) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  mcp_compose_backward_row(dmu + i * A, dsigma + i * A, eo + i * (long)(E * 2 * A), gates + i * E, E, A, deo + i * (long)(E * 2 * A), dlogits + i * E);
}

__global__ void sample_actions_sigma_kernel(const float* __restrict__ mu, const float* __restrict__ sigma, float* __restrict__ actions,
                                            float* __restrict__ logp, float* __restrict__ mu_out, float* __restrict__ sigma_out, int N, int A,
                                            uint32_t seed_lo, uint32_t seed_hi, uint32_t step, const uint32_t* __restrict__ d_step, int env_offset) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  if (d_step) step = *d_step;      // device-resident step counter (CUDA-graph replays of the rollout)
  sample_sigma_row(mu, sigma, actions, logp, mu_out, sigma_out, e, A, seed_lo, seed_hi, step, env_offset);
}

// block reduction of the five loss sums: warp-shuffle butterflies, one shared-memory hop across the 8 warps, one atomic per block per scalar
// (the reduction of ppo_loss_kernel, rl_kernels.cu)
__global__ void __launch_bounds__(256) ppo_loss_sigma_kernel(PpoSigmaArgs p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  PpoRowSums r{0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  if (i < p.M) r = ppo_sigma_row(p, i);
  __shared__ float red[8][5];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float vals[5] = {r.kl, r.surr_a, r.vl, r.ent, r.surr_b};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    float v = vals[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    const int k = threadIdx.x;
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][k];
    atomicAdd(p.scal + (k < 4 ? k : 19), t);
  }
}

}  // namespace go2

using namespace go2;

extern "C" {

int go2_mcp_compose_forward(const float* expert_out, const float* logits, float* gates, float* mu, float* sigma, long n, int E, int A, void* stream) {
  if (!expert_out || !logits || !gates || !mu || !sigma) return set_error(1, "go2_mcp_compose_forward: null argument");
  if (E > MCP_MAX_E || A > MCP_MAX_A) return set_error(1, "go2_mcp_compose_forward: at most 16 experts and 16 actions");
  mcp_compose_fwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(expert_out, logits, gates, mu, sigma, n, E, A);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_mcp_compose_backward(const float* dmu, const float* dsigma, const float* expert_out, const float* gates, float* dexpert_out, float* dlogits, long n,
                             int E, int A, void* stream) {
  if (!dmu || !dsigma || !expert_out || !gates || !dexpert_out || !dlogits) return set_error(1, "go2_mcp_compose_backward: null argument");
  if (E > MCP_MAX_E || A > MCP_MAX_A) return set_error(1, "go2_mcp_compose_backward: at most 16 experts and 16 actions");
  mcp_compose_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(dmu, dsigma, expert_out, gates, dexpert_out, dlogits, n, E, A);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_sample_actions_sigma(const float* mu, const float* sigma, float* actions, float* logp, float* mu_out, float* sigma_out, int N, int A, uint64_t seed,
                             uint32_t step, const uint32_t* d_step, int env_offset, void* stream) {
  if (!mu || !sigma || !actions || !logp || !mu_out || !sigma_out) return set_error(1, "go2_sample_actions_sigma: null argument");
  sample_actions_sigma_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mu, sigma, actions, logp, mu_out, sigma_out, N, A, (uint32_t)seed,
                                                                               (uint32_t)(seed >> 32), step, d_step, env_offset);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_ppo_loss_sigma(const float* mu, const float* sigma, const float* value, const float* actions, const float* old_logp, const float* adv,
                       const float* target_values, const float* returns, const float* old_mu, const float* old_sigma, float* dmu, float* dsigma,
                       float* dvalue, float* scal, int M, int A, float clip, float value_coef, float entropy_coef, int use_clipped_value_loss,
                       float inv_count, int split, float inv_count_a, float inv_count_b, void* stream) {
  if (!dmu || !dsigma || !dvalue || !scal) return set_error(1, "go2_ppo_loss_sigma: null output");
  if (A > MCP_MAX_A) return set_error(1, "go2_ppo_loss_sigma: at most 16 actions");
  cudaStream_t st = (cudaStream_t)stream;
  GO2_CUDA_OK(cudaMemsetAsync(scal, 0, sizeof(float) * (4 + 16), st));
  PpoSigmaArgs p{mu, sigma, value, actions, old_logp, adv, target_values, returns, old_mu, old_sigma, dmu, dsigma, dvalue, scal,
                 M, A, clip, value_coef, entropy_coef, use_clipped_value_loss, inv_count, split, inv_count_a, inv_count_b};
  ppo_loss_sigma_kernel<<<(M + 255) / 256, 256, 0, st>>>(p);
  count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
