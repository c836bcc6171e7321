// Host build of the MCP kernels' row arithmetic (go2_rl_gym_b200/csrc/mcp_core.cuh compiled with g++): the same C entry points as
// csrc/mcp_kernels.cu, each looping over the rows a CUDA thread would own.  TEST TOOLING (tests/emu_rl.py dispatches to it); never a product path.
#include <cstring>
#include "../../go2_rl_gym_b200/csrc/mcp_core.cuh"

using namespace go2;

extern "C" {
int go2_mcp_compose_forward(const float* eo, const float* logits, float* gates, float* mu, float* sigma, long n, int E, int A, void*) {
  if (!eo || !logits || !gates || !mu || !sigma || E > MCP_MAX_E || A > MCP_MAX_A) return 1;
  for (long i = 0; i < n; ++i) mcp_compose_row(eo + i * (long)(E * 2 * A), logits + i * E, E, A, gates + i * E, mu + i * A, sigma + i * A);
  return 0;
}
int go2_mcp_compose_backward(const float* dmu, const float* dsigma, const float* eo, const float* gates, float* deo, float* dlogits, long n, int E, int A, void*) {
  if (!dmu || !dsigma || !eo || !gates || !deo || !dlogits || E > MCP_MAX_E || A > MCP_MAX_A) return 1;
  for (long i = 0; i < n; ++i)
    mcp_compose_backward_row(dmu + i * A, dsigma + i * A, eo + i * (long)(E * 2 * A), gates + i * E, E, A, deo + i * (long)(E * 2 * A), dlogits + i * E);
  return 0;
}
int go2_sample_actions_sigma(const float* mu, const float* sigma, float* actions, float* logp, float* mu_out, float* sigma_out, int N, int A, uint64_t seed,
                             uint32_t step, const uint32_t* d_step, int env_offset, void*) {
  if (!mu || !sigma || !actions || !logp || !mu_out || !sigma_out) return 1;
  if (d_step) step = *d_step;
  for (int e = 0; e < N; ++e) sample_sigma_row(mu, sigma, actions, logp, mu_out, sigma_out, e, A, (uint32_t)seed, (uint32_t)(seed >> 32), step, env_offset);
  return 0;
}
int go2_ppo_loss_sigma(const float* mu, const float* sigma, const float* value, const float* actions, const float* old_logp, const float* adv,
                       const float* target_values, const float* returns, const float* old_mu, const float* old_sigma, float* dmu, float* dsigma,
                       float* dvalue, float* scal, int M, int A, float clip, float value_coef, float entropy_coef, int use_clipped_value_loss,
                       float inv_count, int split, float inv_count_a, float inv_count_b, void*) {
  if (!dmu || !dsigma || !dvalue || !scal || A > MCP_MAX_A) return 1;
  std::memset(scal, 0, sizeof(float) * 20);
  PpoSigmaArgs p{mu, sigma, value, actions, old_logp, adv, target_values, returns, old_mu, old_sigma, dmu, dsigma, dvalue, scal,
                 M, A, clip, value_coef, entropy_coef, use_clipped_value_loss, inv_count, split, inv_count_a, inv_count_b};
  for (int i = 0; i < M; ++i) {
    const PpoRowSums r = ppo_sigma_row(p, i);
    scal[0] += r.kl; scal[1] += r.surr_a; scal[2] += r.vl; scal[3] += r.ent; scal[19] += r.surr_b;
  }
  return 0;
}
}
