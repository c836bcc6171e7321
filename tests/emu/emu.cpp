// Host emulation of the CUDA step kernel: compiles go2_rl_gym_b200/csrc/env_step_core.cuh with g++ and runs the 32 lanes
// of each phase in a loop.  TEST TOOLING (CPU test-suite compares it against the oracle); never a product path.
#include <algorithm>
#include <vector>
#include "../../go2_rl_gym_b200/csrc/env_step_core.cuh"

using namespace go2;
typedef StepT<2, 0, 6> EmuT;   // both roles on every thread, up to two virtual lanes per WIDE thread (Lane.nv)

// One group = the threads that share phase barriers: a single warp (warp-per-env map) or a CTA of 8 warps / 8 envs (packed map).
static int g_packed = 0;
template <class F> static void for_groups(const Go2EnvConfig* C, F&& f) {
  // g_packed: 0 warp per env, 1 / 8 = P2, 4 = Q4, 14 = H14 (half-warp map: 7 WIDE warps + 2 LEGS warps per 14 envs)
  const bool half = g_packed == GO2_HALF_ENVS;
  const int per = g_packed == 1 ? 8 : (g_packed ? g_packed : 1), nwarps = half ? GO2_HALF_ENVS / 2 + 2 : per, NT = 32 * nwarps;
  std::vector<WarpSmem> SM(per);
  std::vector<Lane> lanes(NT);
  for (int e0 = 0; e0 < C->num_envs; e0 += per) {
    const int n_local = std::min(per, C->num_envs - e0);
    for (int t = 0; t < NT; ++t) init_roles(lanes[t], t, half ? 2 : (g_packed ? 1 : 0), e0, n_local, nwarps);
    f(lanes.data(), NT, SM.data());
  }
}

extern "C" {
void go2_emu_set_packed(int packed) { g_packed = packed; }
int go2_emu_step(const Go2EnvConfig* C, const Go2Model* M, const Go2EnvBuffers* B, const float* actions, const Go2StepParams* sp) {
  for (int k = 0; k < GO2_EP_ACCUM_FLOATS; ++k) B->ep_accum[k] = 0;
  StepCtx X{C, M, B, sp, actions, C};
  for_groups(C, [&](Lane* lanes, int NT, WarpSmem* SM) { step_env<EmuT>(lanes, NT, SM, X); });
  // finalize extras["episode"] (mirrors the tiny finalize kernel)
  float n_reset = B->ep_accum[GO2_NUM_REW + 10];
  if (C->num_xrew > 0) {
    long long* xacc = GO2_EXT_PTR(long long*, C, ext_xrew_log);
    float* xst = reinterpret_cast<float*>(xacc + GO2_NUM_XREW);
    for (int k = 0; k < GO2_NUM_XREW; ++k) {
      if (n_reset > 0) xst[(size_t)sp->ep_slot * GO2_NUM_XREW + k] = (float)((double)xacc[k] / (double)GO2_EP_FIXED_ONE) / n_reset / C->max_episode_length_s;
      else xst[(size_t)sp->ep_slot * GO2_NUM_XREW + k] = xst[(size_t)((sp->ep_slot + GO2_EP_SLOTS - 1) % GO2_EP_SLOTS) * GO2_NUM_XREW + k];
      xacc[k] = 0;
    }
  }
  if (n_reset > 0 && B->ep_stats) {
    float* st = B->ep_stats + (size_t)sp->ep_slot * GO2_EP_STATS;
    for (int k = 0; k < GO2_NUM_REW; ++k)
      st[k] = (float)((double)reinterpret_cast<const long long*>(B->ep_accum + GO2_EP_ACC_FIXED_OFF)[k] / (double)GO2_EP_FIXED_ONE) / n_reset / C->max_episode_length_s;
    std::vector<int> cnt(9, 0);
    for (int e = 0; e < C->num_envs; ++e) cnt[B->terrain_ids[e]]++;
    st[GO2_NUM_REW] = C->mesh_type == 0 ? 0.0f : B->ep_accum[GO2_NUM_REW] / C->num_envs;
    for (int t = 0; t < 9; ++t) st[GO2_NUM_REW + 1 + t] = cnt[t] ? B->ep_accum[GO2_NUM_REW + 1 + t] / cnt[t] : 0.0f;
    st[GO2_NUM_REW + 10] = n_reset;
    st[GO2_NUM_REW + 11] = 1.0f;
  } else if (B->ep_stats) {
    const float* prev = B->ep_stats + (size_t)((sp->ep_slot + GO2_EP_SLOTS - 1) % GO2_EP_SLOTS) * GO2_EP_STATS;
    float* st = B->ep_stats + (size_t)sp->ep_slot * GO2_EP_STATS;
    for (int k = 0; k < GO2_EP_STATS; ++k) st[k] = prev[k];
  }
  return 0;
}
int go2_emu_reset_all(const Go2EnvConfig* C, const Go2Model* M, const Go2EnvBuffers* B, const Go2StepParams* sp) {
  StepCtx X{C, M, B, sp, nullptr, C};
  for_groups(C, [&](Lane* lanes, int NT, WarpSmem* SM) { reset_env_initial<EmuT>(lanes, NT, SM, X); });
  return 0;
}
int go2_emu_substeps(const Go2EnvConfig* C, const Go2Model* M, const Go2EnvBuffers* B, const float* tau, int n) {
  StepCtx X{C, M, B, nullptr, nullptr, C};
  for_groups(C, [&](Lane* lanes, int NT, WarpSmem* SM) { substeps_env<EmuT>(lanes, NT, SM, X, tau, n); });
  return 0;
}
int go2_emu_sizeof_warp_smem(void) { return (int)sizeof(WarpSmem); }
}
