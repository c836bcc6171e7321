// mcp_core.cuh — per-row arithmetic of the Multiplicative-Compositional-Policy (MCP) actor head and of the PPO pieces that need a
// state-dependent action sigma.  Host/device source: csrc/mcp_kernels.cu launches these rows one per thread, tests/emu/emu_mcp.cpp loops
// over them with g++ so the CPU suite exercises the kernels' own arithmetic.
//
// Reference semantics:
//   ActorMCP.forward            rsl_rl/modules/actor_critic_mcp_cts.py:220-247  (sigmoid gate, per-expert (mu, log_std) with log_std clamped to
//                               [-5, 2], var = exp(2 log_std) + 1e-9, product of the gated Gaussians)
//   Normal(mean, std) sampling  rsl_rl/modules/actor_critic_mcp_cts.py:146-164, rsl_rl/algorithms/cts.py:112-142
//   surrogate / value / KL      rsl_rl/algorithms/mcp_cts.py:133-181
#pragma once
#include <math.h>
#include <stdint.h>

#include "env_step_core.cuh"  // GO2_HD, philox, u01

namespace go2 {

constexpr int MCP_MAX_E = 16, MCP_MAX_A = 16;

// expert_out row: expert e occupies [e * 2A, (e + 1) * 2A) = [mu_e (A) | log_std_e (A)]   (torch.chunk(expert_out, 2, dim=-1), :236)
GO2_HD void mcp_compose_row(const float* eo, const float* logits, int E, int A, float* gates, float* mu, float* sigma) {
  float w[MCP_MAX_E];
  for (int e = 0; e < E; ++e) { w[e] = 1.0f / (1.0f + expf(-logits[e])); gates[e] = w[e]; }
  for (int k = 0; k < A; ++k) {
    float s = 0.0f, ms = 0.0f;
    for (int e = 0; e < E; ++e) {
      const float ls = fminf(fmaxf(eo[e * 2 * A + A + k], -5.0f), 2.0f);
      const float inv = 1.0f / (expf(2.0f * ls) + 1e-9f);
      s += w[e] * inv;
      ms += w[e] * eo[e * 2 * A + k] * inv;
    }
    const float vt = 1.0f / (s + 1e-9f);
    sigma[k] = sqrtf(vt);
    mu[k] = vt * ms;
  }
}

// d loss / d expert_out row and d loss / d gate logits from d loss / d mu, d loss / d sigma of one row
GO2_HD void mcp_compose_backward_row(const float* dmu, const float* dsigma, const float* eo, const float* gates, int E, int A, float* deo, float* dlogits) {
  float dw[MCP_MAX_E];
  for (int e = 0; e < E; ++e) dw[e] = 0.0f;
  for (int k = 0; k < A; ++k) {
    float inv[MCP_MAX_E], s = 0.0f, ms = 0.0f;
    for (int e = 0; e < E; ++e) {
      const float ls = fminf(fmaxf(eo[e * 2 * A + A + k], -5.0f), 2.0f);
      inv[e] = 1.0f / (expf(2.0f * ls) + 1e-9f);
      s += gates[e] * inv[e];
      ms += gates[e] * eo[e * 2 * A + k] * inv[e];
    }
    const float vt = 1.0f / (s + 1e-9f), sg = sqrtf(vt);
    const float dvt = dmu[k] * ms + dsigma[k] * 0.5f / sg;       // mu = vt ms, sigma = sqrt(vt)
    const float dms = dmu[k] * vt;
    const float ds = -dvt * vt * vt;                             // vt = 1 / (s + eps)
    for (int e = 0; e < E; ++e) {
      const float m = eo[e * 2 * A + k], raw = eo[e * 2 * A + A + k];
      dw[e] += (ds + dms * m) * inv[e];
      deo[e * 2 * A + k] = dms * gates[e] * inv[e];
      const float dvar = -(ds + dms * m) * gates[e] * inv[e] * inv[e];
      const float ls = fminf(fmaxf(raw, -5.0f), 2.0f);
      // clamp passes the gradient on [-5, 2] (bounds included, like torch.clamp)
      deo[e * 2 * A + A + k] = (raw >= -5.0f && raw <= 2.0f) ? dvar * 2.0f * expf(2.0f * ls) : 0.0f;
    }
  }
  for (int e = 0; e < E; ++e) dlogits[e] = dw[e] * gates[e] * (1.0f - gates[e]);
}

// a = mu + sigma z with z ~ N(0,1) from Philox(seed; env, step, stream 16, block) — the same draws as sample_actions_kernel (rl_kernels.cu)
GO2_HD void sample_sigma_row(const float* mu, const float* sigma, float* actions, float* logp, float* mu_out, float* sigma_out, int e, int A,
                             uint32_t seed_lo, uint32_t seed_hi, uint32_t step, int env_offset) {
  float lp = 0.0f;
  for (int b = 0; b < (A + 3) / 4; ++b) {
    U4 r = philox((uint32_t)(env_offset + e), step, 16u, (uint32_t)b, seed_lo, seed_hi);
    const float u0 = ((float)(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f), u1 = u01(r.y);
    const float u2 = ((float)(r.z >> 8) + 1.0f) * (1.0f / 16777216.0f), u3 = u01(r.w);
    const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
    const float z[4] = {r0 * cosf(6.283185307179586f * u1), r0 * sinf(6.283185307179586f * u1), r1 * cosf(6.283185307179586f * u3),
                        r1 * sinf(6.283185307179586f * u3)};
    for (int q = 0; q < 4; ++q) {
      const int k = 4 * b + q;
      if (k >= A) break;
      const float m = mu[(long)e * A + k], s = sigma[(long)e * A + k];
      const float a = m + s * z[q];
      actions[(long)e * A + k] = a;
      mu_out[(long)e * A + k] = m;
      sigma_out[(long)e * A + k] = s;
      const float d = a - m;
      lp += -(d * d) / (2.0f * s * s) - logf(s) - 0.9189385332046727f;
    }
  }
  logp[e] = lp;
}

// PPO losses of one sample with a state-dependent sigma: same terms as ppo_loss_kernel (rl_kernels.cu), the gradient w.r.t. sigma goes to
// dsigma[M, A] (it flows on into the MCP head) instead of the std parameter's accumulator
struct PpoSigmaArgs {
  const float* mu; const float* sigma; const float* value; const float* actions; const float* old_logp; const float* adv;
  const float* target_values; const float* returns; const float* old_mu; const float* old_sigma;
  float* dmu; float* dsigma; float* dvalue; float* scal;
  int M, A; float clip, value_coef, entropy_coef; int use_clipped_value_loss; float inv_count;
  int split; float inv_count_a, inv_count_b;
};
struct PpoRowSums { float kl, surr_a, surr_b, vl, ent; };

GO2_HD PpoRowSums ppo_sigma_row(const PpoSigmaArgs& p, int i) {
  PpoRowSums out{0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  float lp = 0.0f;
  for (int k = 0; k < p.A; ++k) {
    const long o = (long)i * p.A + k;
    const float m = p.mu[o], s = p.sigma[o], d = p.actions[o] - m;
    lp += -(d * d) / (2.0f * s * s) - logf(s) - 0.9189385332046727f;
    const float om = p.old_mu[o], os = p.old_sigma[o];
    out.kl += logf(s / os + 1.e-5f) + (os * os + (om - m) * (om - m)) / (2.0f * s * s) - 0.5f;
    out.ent += 0.5f + 0.9189385332046727f + logf(s);
  }
  const float A_ = p.adv[i];
  const float ratio = expf(lp - p.old_logp[i]);
  const float s1 = -A_ * ratio, s2 = -A_ * fminf(fmaxf(ratio, 1.0f - p.clip), 1.0f + p.clip);
  if (i < p.split) out.surr_a = fmaxf(s1, s2); else out.surr_b = fmaxf(s1, s2);
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
    out.vl = fmaxf(l1, l2);
    if (l1 >= l2) dv = 2.0f * (v - ret);
    else dv = (diff > -p.clip && diff < p.clip) ? 2.0f * (vc - ret) : 0.0f;
  } else { out.vl = (ret - v) * (ret - v); dv = 2.0f * (v - ret); }
  p.dvalue[i] = p.value_coef * dv * p.inv_count;
  for (int k = 0; k < p.A; ++k) {
    const long o = (long)i * p.A + k;
    const float m = p.mu[o], s = p.sigma[o], d = p.actions[o] - m;
    p.dmu[o] = dlp * d / (s * s);
    p.dsigma[o] = dlp * (d * d / (s * s * s) - 1.0f / s) - p.entropy_coef * p.inv_count / s;
  }
  return out;
}

}  // namespace go2
