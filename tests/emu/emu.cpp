// Host emulation of the CUDA step kernel: compiles go2_rl_gym_b200/csrc/env_step_core.cuh with g++ and runs the 32 lanes
// of each phase in a loop.  TEST TOOLING (CPU test-suite compares it against the oracle); never a product path.
#include <vector>
#include "../../go2_rl_gym_b200/csrc/env_step_core.cuh"

using namespace go2;

extern "C" {
int go2_emu_step(const Go2EnvConfig* C, const Go2Model* M, const Go2EnvBuffers* B, const float* actions, const Go2StepParams* sp) {
  for (int k = 0; k < GO2_EP_STATS + 2; ++k) B->ep_accum[k] = 0;
  StepCtx X{C, M, B, sp, actions};
  for (int e = 0; e < C->num_envs; ++e) {
    WarpSmem S;
    Lane lanes[32];
    step_env(lanes, S, X, e);
  }
  // finalize extras["episode"] (mirrors the tiny finalize kernel)
  float n_reset = B->ep_accum[GO2_NUM_REW + 10];
  if (n_reset > 0 && B->ep_stats) {
    float* st = B->ep_stats + (size_t)sp->ep_slot * GO2_EP_STATS;
    for (int k = 0; k < GO2_NUM_REW; ++k) st[k] = B->ep_accum[k] / n_reset / C->max_episode_length_s;
    std::vector<int> cnt(9, 0);
    for (int e = 0; e < C->num_envs; ++e) cnt[B->terrain_ids[e]]++;
    st[GO2_NUM_REW] = C->mesh_type == 0 ? 0.0f : B->ep_accum[GO2_NUM_REW] / C->num_envs;
    for (int t = 0; t < 9; ++t) st[GO2_NUM_REW + 1 + t] = cnt[t] ? B->ep_accum[GO2_NUM_REW + 1 + t] / cnt[t] : 0.0f;
    st[GO2_NUM_REW + 10] = n_reset;
    st[GO2_NUM_REW + 11] = 1.0f;
  }
  return 0;
}
int go2_emu_reset_all(const Go2EnvConfig* C, const Go2Model* M, const Go2EnvBuffers* B, const Go2StepParams* sp) {
  StepCtx X{C, M, B, sp, nullptr};
  for (int e = 0; e < C->num_envs; ++e) { WarpSmem S; Lane lanes[32]; reset_env_initial(lanes, S, X, e); }
  return 0;
}
int go2_emu_substeps(const Go2EnvConfig* C, const Go2Model* M, const Go2EnvBuffers* B, const float* tau, int n) {
  StepCtx X{C, M, B, nullptr, nullptr};
  for (int e = 0; e < C->num_envs; ++e) { WarpSmem S; Lane lanes[32]; substeps_env(lanes, S, X, e, tau, n); }
  return 0;
}
int go2_emu_sizeof_warp_smem(void) { return (int)sizeof(WarpSmem); }
}
