// env_step.cu — sm_100a kernels of the fused Go2 environment step and their C ABI (include/go2_b200.h).
// One warp per env (lane roles in env_step_core.cuh), 4 envs per CTA, per-env rows read/written coalesced,
// per-link inertias and all solver state staged in shared memory.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "env_step_core.cuh"
#include "common.cuh"

namespace go2 {

constexpr int WARPS_PER_CTA = 4;

__global__ void __launch_bounds__(32 * WARPS_PER_CTA, 4)
step_kernel(const Go2EnvConfig* __restrict__ cfg, const Go2Model* __restrict__ mdl, const __grid_constant__ Go2EnvBuffers buf,
            const __grid_constant__ Go2StepParams sp, const float* __restrict__ actions) {
  __shared__ WarpSmem smem[WARPS_PER_CTA];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * WARPS_PER_CTA + warp;
  if (e >= cfg->num_envs) return;
  StepCtx X{cfg, mdl, &buf, &sp, actions};
  Lane L;
  L.nsync = 0; L.ncoarse = 0; L.nmid = 0;
  step_env(lane, L, smem[warp], X, e);
}

// Same step, W warps per CTA with dynamic shared memory; LOCKSTEP: phases end in a CTA-wide named barrier over the warps that own an
// env, so all warps of the CTA stream the same instructions (the step is instruction-fetch bound: ~155 KB of straight-line code)
template <int W, int MINB, int LOCKSTEP>     // LOCKSTEP: 0 none, 1 every phase, 2 substep boundaries only
__global__ void __launch_bounds__(32 * W, MINB)
step_kernel_wide(const Go2EnvConfig* __restrict__ cfg, const Go2Model* __restrict__ mdl, const __grid_constant__ Go2EnvBuffers buf,
                 const __grid_constant__ Go2StepParams sp, const float* __restrict__ actions) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  WarpSmem* smem = reinterpret_cast<WarpSmem*>(smem_dyn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * W + warp;
  if (e >= cfg->num_envs) return;
  StepCtx X{cfg, mdl, &buf, &sp, actions};
  Lane L;
  L.nsync = LOCKSTEP == 1 ? 32 * min(W, cfg->num_envs - blockIdx.x * W) : 0;
  L.ncoarse = LOCKSTEP >= 2 ? 32 * min(W, cfg->num_envs - blockIdx.x * W) : 0;
  L.nmid = LOCKSTEP == 3 ? L.ncoarse : 0;
  step_env(lane, L, smem[warp], X, e);
}

__global__ void __launch_bounds__(32 * WARPS_PER_CTA)
reset_kernel(const Go2EnvConfig* __restrict__ cfg, const Go2Model* __restrict__ mdl, const __grid_constant__ Go2EnvBuffers buf,
             const __grid_constant__ Go2StepParams sp) {
  __shared__ WarpSmem smem[WARPS_PER_CTA];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * WARPS_PER_CTA + warp;
  if (e >= cfg->num_envs) return;
  StepCtx X{cfg, mdl, &buf, &sp, nullptr};
  Lane L;
  L.nsync = 0; L.ncoarse = 0; L.nmid = 0;
  reset_env_initial(lane, L, smem[warp], X, e);
}

__global__ void __launch_bounds__(32 * WARPS_PER_CTA)
substeps_kernel(const Go2EnvConfig* __restrict__ cfg, const Go2Model* __restrict__ mdl, const __grid_constant__ Go2EnvBuffers buf,
                const float* __restrict__ tau, int n) {
  __shared__ WarpSmem smem[WARPS_PER_CTA];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * WARPS_PER_CTA + warp;
  if (e >= cfg->num_envs) return;
  StepCtx X{cfg, mdl, &buf, nullptr, nullptr};
  Lane L;
  L.nsync = 0; L.ncoarse = 0; L.nmid = 0;
  substeps_env(lane, L, smem[warp], X, e, tau, n);
}

// extras["episode"] (legged_robot.py:229-242): refreshed only when at least one env reset this step; then clear the sums
__global__ void finalize_kernel(const Go2EnvConfig* __restrict__ cfg, float* __restrict__ ep_accum, float* __restrict__ ep_stats,
                                const float* __restrict__ id_counts, int slot) {
  const int k = threadIdx.x;
  const float n_reset = ep_accum[GO2_NUM_REW + 10];
  __syncthreads();
  if (n_reset > 0.0f && ep_stats != nullptr) {
    float* st = ep_stats + (size_t)slot * GO2_EP_STATS;
    if (k < GO2_NUM_REW) st[k] = ep_accum[k] / n_reset / cfg->max_episode_length_s;
    else if (k == GO2_NUM_REW) st[k] = cfg->mesh_type == 0 ? 0.0f : ep_accum[k] / (float)cfg->num_envs;
    else if (k < GO2_NUM_REW + 10) st[k] = id_counts[k - GO2_NUM_REW - 1] > 0 ? ep_accum[k] / id_counts[k - GO2_NUM_REW - 1] : 0.0f;
    else if (k == GO2_NUM_REW + 10) st[k] = n_reset;
    else if (k == GO2_NUM_REW + 11) st[k] = 1.0f;
  }
  __syncthreads();
  if (k < GO2_EP_STATS + 2) ep_accum[k] = 0.0f;
}

}  // namespace go2

// ================================================================================================ C ABI
struct Go2Env {
  Go2EnvConfig cfg;
  Go2Model mdl;
  Go2EnvBuffers buf;
  Go2EnvConfig* d_cfg = nullptr;
  Go2Model* d_mdl = nullptr;
  float* d_actions = nullptr;     // staging for the host-buffer entry point
  float* d_id_counts = nullptr;
  int grid = 0;
};

namespace go2 {
template <int W, int MINB, int LOCKSTEP>
static int launch_wide(Go2Env* h, const float* actions, const Go2StepParams* sp, cudaStream_t st) {
  const int smem = W * (int)sizeof(WarpSmem);
  static bool attr = false;
  if (!attr) {
    GO2_CUDA_OK(cudaFuncSetAttribute(step_kernel_wide<W, MINB, LOCKSTEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  step_kernel_wide<W, MINB, LOCKSTEP><<<(h->cfg.num_envs + W - 1) / W, 32 * W, smem, st>>>(h->d_cfg, h->d_mdl, h->buf, *sp, actions);
  return 0;
}
}  // namespace go2

extern "C" {

int go2_env_create(const Go2EnvConfig* cfg, const Go2Model* model, const Go2EnvBuffers* bufs, Go2Env** out) {
  if (!cfg || !model || !bufs || !out) return go2::set_error(1, "go2_env_create: null argument");
  if (cfg->num_envs <= 0) return go2::set_error(1, "go2_env_create: num_envs must be positive");
  // the kernel bakes the Go2 topology: hip = x axis, thigh/calf = y axis, collider lanes grouped per body
  for (int j = 0; j < GO2_NUM_DOF; ++j)
    if (model->joint_axis[j] != ((j % 3 == 0) ? 0 : 1)) return go2::set_error(2, "go2_env_create: joint axes must be x,y,y per leg");
  for (int c = 0; c < GO2_NUM_COL; ++c) {
    int dyn, rep;
    if (c < 8) { dyn = 0; rep = c < 6 ? 0 : c - 5; }
    else { int l = (c - 8) / 6, k = (c - 8) % 6; dyn = 1 + 3 * l + (k == 0 ? 0 : (k == 1 ? 1 : 2)); rep = 3 + 4 * l + (k == 0 ? 0 : (k == 1 ? 1 : (k < 5 ? 2 : 3))); }
    if (model->col_dyn[c] != dyn || model->col_report[c] != rep) return go2::set_error(2, "go2_env_create: collider layout does not match the kernel's lane map");
  }
  Go2Env* h = new Go2Env();
  h->cfg = *cfg; h->mdl = *model; h->buf = *bufs;
  h->grid = (cfg->num_envs + go2::WARPS_PER_CTA - 1) / go2::WARPS_PER_CTA;
  GO2_CUDA_OK(cudaMalloc(&h->d_cfg, sizeof(Go2EnvConfig)));
  GO2_CUDA_OK(cudaMalloc(&h->d_mdl, sizeof(Go2Model)));
  GO2_CUDA_OK(cudaMalloc(&h->d_actions, sizeof(float) * GO2_NUM_DOF * cfg->num_envs));
  GO2_CUDA_OK(cudaMalloc(&h->d_id_counts, sizeof(float) * 9));
  GO2_CUDA_OK(cudaMemcpy(h->d_cfg, cfg, sizeof(Go2EnvConfig), cudaMemcpyHostToDevice));
  GO2_CUDA_OK(cudaMemcpy(h->d_mdl, model, sizeof(Go2Model), cudaMemcpyHostToDevice));
  std::vector<int32_t> ids(cfg->num_envs);
  GO2_CUDA_OK(cudaMemcpy(ids.data(), bufs->terrain_ids, sizeof(int32_t) * cfg->num_envs, cudaMemcpyDeviceToHost));
  float counts[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int e = 0; e < cfg->num_envs; ++e) if (ids[e] >= 0 && ids[e] < 9) counts[ids[e]] += 1.0f;
  GO2_CUDA_OK(cudaMemcpy(h->d_id_counts, counts, sizeof(counts), cudaMemcpyHostToDevice));
  GO2_CUDA_OK(cudaMemset(bufs->ep_accum, 0, sizeof(float) * (GO2_EP_STATS + 2)));
  *out = h;
  return 0;
}

void go2_env_destroy(Go2Env* h) {
  if (!h) return;
  cudaFree(h->d_cfg); cudaFree(h->d_mdl); cudaFree(h->d_actions); cudaFree(h->d_id_counts);
  delete h;
}

int go2_env_step(Go2Env* h, const float* actions, const Go2StepParams* sp, void* stream) {
  if (!h || !actions || !sp) return go2::set_error(1, "go2_env_step: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  // default: 8 warps (envs) per CTA, 2 CTAs/SM, CTA-wide barrier at substep boundaries only ("8p"): the step is instruction-fetch bound
  // (~155 KB of straight-line code), and warps that stay on the same stretch of code share the instruction caches: 239 -> 203 us at
  // 4096 envs, 447 -> 356 us at 8192.  GO2_STEP_MODE (tuning aid): "4" = 4 warps/CTA, no barrier; "<W>" / "<W>s" (every phase) /
  // "<W>p" (substep boundaries) / "<W>q" (+3 points inside a substep)
  static int mode = -1;
  if (mode < 0) {
    const char* m = getenv("GO2_STEP_MODE");
    mode = !m ? 7 : !strcmp(m, "4") ? 0 : !strcmp(m, "16s") ? 3 : !strcmp(m, "16") ? 2 : !strcmp(m, "8s") ? 1 : !strcmp(m, "4s") ? 4 : !strcmp(m, "8") ? 5 : !strcmp(m, "16p") ? 6
           : !strcmp(m, "8p") ? 7 : !strcmp(m, "12") ? 8 : !strcmp(m, "8q") ? 9 : !strcmp(m, "4p") ? 10 : !strcmp(m, "4q") ? 11 : !strcmp(m, "16q") ? 12 : 0;
  }
  if (mode == 0) go2::step_kernel<<<h->grid, 32 * go2::WARPS_PER_CTA, 0, st>>>(h->d_cfg, h->d_mdl, h->buf, *sp, actions);
  else {
    int rc = mode == 3 ? go2::launch_wide<16, 1, 1>(h, actions, sp, st) : mode == 2 ? go2::launch_wide<16, 1, 0>(h, actions, sp, st)
           : mode == 1 ? go2::launch_wide<8, 2, 1>(h, actions, sp, st) : mode == 4 ? go2::launch_wide<4, 4, 1>(h, actions, sp, st)
           : mode == 5 ? go2::launch_wide<8, 2, 0>(h, actions, sp, st) : mode == 6 ? go2::launch_wide<16, 1, 2>(h, actions, sp, st)
           : mode == 7 ? go2::launch_wide<8, 2, 2>(h, actions, sp, st) : mode == 9 ? go2::launch_wide<8, 2, 3>(h, actions, sp, st)
           : mode == 10 ? go2::launch_wide<4, 4, 2>(h, actions, sp, st) : mode == 11 ? go2::launch_wide<4, 4, 3>(h, actions, sp, st)
           : mode == 12 ? go2::launch_wide<16, 1, 3>(h, actions, sp, st) : go2::launch_wide<12, 1, 0>(h, actions, sp, st);
    if (rc) return rc;
  }
  go2::count_launch();
  go2::finalize_kernel<<<1, 32, 0, st>>>(h->d_cfg, h->buf.ep_accum, h->buf.ep_stats, h->d_id_counts, sp->ep_slot);
  go2::count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_env_step_host(Go2Env* h, const float* h_actions, const Go2StepParams* sp, float* h_obs, float* h_priv, float* h_rew,
                      uint8_t* h_reset, void* stream) {
  if (!h || !h_actions || !sp) return go2::set_error(1, "go2_env_step_host: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t N = (size_t)h->cfg.num_envs;
  GO2_CUDA_OK(cudaMemcpyAsync(h->d_actions, h_actions, sizeof(float) * GO2_NUM_DOF * N, cudaMemcpyHostToDevice, st));
  int rc = go2_env_step(h, h->d_actions, sp, stream);
  if (rc) return rc;
  if (h_obs) GO2_CUDA_OK(cudaMemcpyAsync(h_obs, h->buf.obs_buf, sizeof(float) * GO2_NUM_OBS * N, cudaMemcpyDeviceToHost, st));
  if (h_priv) GO2_CUDA_OK(cudaMemcpyAsync(h_priv, h->buf.privileged_obs_buf, sizeof(float) * GO2_NUM_PRIV * N, cudaMemcpyDeviceToHost, st));
  if (h_rew) GO2_CUDA_OK(cudaMemcpyAsync(h_rew, h->buf.rew_buf, sizeof(float) * N, cudaMemcpyDeviceToHost, st));
  if (h_reset) GO2_CUDA_OK(cudaMemcpyAsync(h_reset, h->buf.reset_buf, N, cudaMemcpyDeviceToHost, st));
  GO2_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int go2_env_reset_all(Go2Env* h, const Go2StepParams* sp, void* stream) {
  if (!h || !sp) return go2::set_error(1, "go2_env_reset_all: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  go2::reset_kernel<<<h->grid, 32 * go2::WARPS_PER_CTA, 0, st>>>(h->d_cfg, h->d_mdl, h->buf, *sp);
  go2::count_launch();
  GO2_CUDA_OK(cudaMemsetAsync(h->buf.ep_accum, 0, sizeof(float) * (GO2_EP_STATS + 2), st));
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_env_substeps(Go2Env* h, const float* tau, int n, void* stream) {
  if (!h || !tau) return go2::set_error(1, "go2_env_substeps: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  go2::substeps_kernel<<<h->grid, 32 * go2::WARPS_PER_CTA, 0, st>>>(h->d_cfg, h->d_mdl, h->buf, tau, n);
  go2::count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
