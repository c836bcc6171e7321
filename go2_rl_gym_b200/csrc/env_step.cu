// env_step.cu — sm_100a kernels of the fused Go2 environment step and their C ABI (include/go2_b200.h).
// The step body (phases, roles) is env_step_core.cuh; this file holds one kernel per thread map — default "H14" = step_kernel_half: 14 envs per 9-warp
// CTA, half a warp per env + dedicated leg warps, 28 envs resident per SM — the per-env rows read / written coalesced, per-link inertias and all solver
// state staged in shared memory.
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
typedef StepT<1, 0, 6> T1;   // one virtual lane per WIDE thread, both roles in every thread, scratch stride 1 mod 32 words

__global__ void __launch_bounds__(32 * WARPS_PER_CTA, 4)
step_kernel(const Go2EnvConfig* __restrict__ cfg, const Go2Model* __restrict__ mdl, const __grid_constant__ Go2EnvBuffers buf,
            const Go2StepParams* __restrict__ sp, const float* __restrict__ actions) {
  __shared__ WarpSmem smem[WARPS_PER_CTA];
  const int e0 = blockIdx.x * WARPS_PER_CTA;
  StepCtx X{cfg, mdl, &buf, sp, actions, cfg};
  Lane L;
  init_roles(L, threadIdx.x, 0, e0, min(WARPS_PER_CTA, cfg->num_envs - e0), WARPS_PER_CTA);
  if (!L.own) return;
  step_env<T1>(L, smem, X);
}

// Same step, W warps per CTA with dynamic shared memory; LOCKSTEP: phases end in a CTA-wide named barrier, so all warps of the CTA
// stream the same instructions (the step is instruction-fetch bound: ~155 KB of straight-line code)
template <int W, int MINB, int LOCKSTEP>     // LOCKSTEP: 0 none, 1 every phase, 2 substep boundaries only, 3 + three points inside a substep
__global__ void __launch_bounds__(32 * W, MINB)
step_kernel_wide(const Go2EnvConfig* __restrict__ cfg, const Go2Model* __restrict__ mdl, const __grid_constant__ Go2EnvBuffers buf,
                 const Go2StepParams* __restrict__ sp, const float* __restrict__ actions) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  WarpSmem* smem = reinterpret_cast<WarpSmem*>(smem_dyn);
  const int e0 = blockIdx.x * W, n_local = min(W, cfg->num_envs - e0);
  StepCtx X{cfg, mdl, &buf, sp, actions, cfg};
  Lane L;
  init_roles(L, threadIdx.x, 0, e0, n_local, W);
  if (!L.own) return;
  L.nsync = LOCKSTEP == 1 ? 32 * n_local : 0;
  L.ncoarse = LOCKSTEP >= 2 ? 32 * n_local : 0;
  L.nmid = LOCKSTEP == 3 ? L.ncoarse : 0;
  step_env<T1>(L, smem, X);
}

// Leg-warp rotation of the packed maps: CTAs that land on the same SM take consecutive tickets, so their leg warps (ticket % warps) sit on
// different warp schedulers (warp w of a CTA runs on scheduler w % 4).  Results do not depend on the choice (same per-thread code).
__device__ unsigned int g_sm_ticket[1024];
__device__ __forceinline__ int pick_leg_warp(int nwarps, int* slot) {
  if (threadIdx.x == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    *slot = (int)(atomicAdd(&g_sm_ticket[smid & 1023u], 1u) % (unsigned int)nwarps);
  }
  __syncthreads();
  return *slot;
}

// PACKED thread map (env_step_core.cuh): a CTA of 8 warps owns 8 envs; the serial leg recursions of all 8 envs run in warp 0 (one
// (env, leg) item per lane), the base 6x6 factorisations in warps 1-2; every phase ends in a CTA barrier.  The leg code is ~80 % of the
// step's instructions and used 4 of 32 lanes with the warp-per-env map: here it is issued once per 8 envs.
template <int MINB, int ROT>
__global__ void __launch_bounds__(256, MINB)
step_kernel_packed(const Go2EnvConfig* __restrict__ cfg, const Go2Model* __restrict__ mdl, const __grid_constant__ Go2EnvBuffers buf,
                   const Go2StepParams* __restrict__ sp, const float* __restrict__ actions) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  WarpSmem* smem = reinterpret_cast<WarpSmem*>(smem_dyn);
  __shared__ int leg_warp_slot;
  const int e0 = blockIdx.x * 8;
  StepCtx X{cfg, mdl, &buf, sp, actions, cfg};
  Lane L;
  init_roles(L, threadIdx.x, 1, e0, min(8, cfg->num_envs - e0), 8, ROT == 1 ? pick_leg_warp(8, &leg_warp_slot) : ROT == 2 ? (int)((blockIdx.x / 148u) & 3u) : 0);
#if defined(GO2_PHASE_TIMING)
  if (threadIdx.x == 0 && blockIdx.x == GO2_PHASE_TIMING) { go2_ph_count = 1; go2_ph_clock[0] = clock64(); }
#endif
  step_env<T1>(L, smem, X);
#if defined(GO2_PHASE_TIMING)
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == GO2_PHASE_TIMING) { const int k = go2_ph_count; if (k < 512) go2_ph_clock[k] = clock64(); go2_ph_count = k + 1; }
#endif
}

// The packed map with FEWER envs per CTA (E = 4: 128 threads, 4 CTAs / SM at 128 registers): the same 16 envs per SM as P2, but four independent
// barrier groups instead of two, so the serial LEGS phase of one CTA (one busy warp) overlaps the WIDE phases of three others.  The leg warp then
// carries 4 E items (half its lanes at E = 4): more issued instructions, which the 25 %-used issue slots absorb.  Built after round 1's GPU budget
// was spent: bit-identical to the other maps in the host emulation, unmeasured on hardware (mode "Q4").
template <int E, int MINB>
__global__ void __launch_bounds__(32 * E, MINB)
step_kernel_quad(const Go2EnvConfig* __restrict__ cfg, const Go2Model* __restrict__ mdl, const __grid_constant__ Go2EnvBuffers buf,
                 const Go2StepParams* __restrict__ sp, const float* __restrict__ actions) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  WarpSmem* smem = reinterpret_cast<WarpSmem*>(smem_dyn);
  __shared__ int leg_warp_slot;
  const int e0 = blockIdx.x * E;
  StepCtx X{cfg, mdl, &buf, sp, actions, cfg};
  Lane L;
  init_roles(L, threadIdx.x, 1, e0, min(E, cfg->num_envs - e0), E, pick_leg_warp(E, &leg_warp_slot));
  step_env<T1>(L, smem, X);
}

// HALF-WARP map "H14" (env_step_core.cuh: init_roles, packed == 2): a CTA of 9 warps owns 14 envs — 7 WIDE warps (16 threads per env, two items per
// thread) and 2 dedicated LEGS warps — at 112 registers, so that TWO such CTAs (28 envs) are resident per SM and 4096 envs are ONE wave on the 148
// SMs (the packed map holds 16 envs per SM: 1.73 waves, i.e. two CTA-steps back to back).  The warps are specialised at compile time: each role's
// instantiation carries only its own per-thread state and instruction stream.
constexpr int HALF_WARPS = GO2_HALF_ENVS / 2 + 2;
typedef WarpSmemT<7> HalfSmem;
#if !defined(GO2_HALF_MINB)
#define GO2_HALF_MINB 2      /* CTAs per SM the register allocation aims at (1: tuning experiment, up to 224 registers) */
#endif
__global__ void __launch_bounds__(32 * HALF_WARPS, GO2_HALF_MINB)
step_kernel_half(const Go2EnvConfig* __restrict__ cfg, const Go2Model* __restrict__ mdl, const __grid_constant__ Go2EnvBuffers buf,
                 const Go2StepParams* __restrict__ sp, const float* __restrict__ actions, const __grid_constant__ Go2EnvConfig cfgv) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  HalfSmem* smem = reinterpret_cast<HalfSmem*>(smem_dyn);
  const int e0 = blockIdx.x * GO2_HALF_ENVS;
  StepCtx X{cfg, mdl, &buf, sp, actions, &cfgv};
  Lane L;
  init_roles(L, threadIdx.x, 2, e0, min(GO2_HALF_ENVS, cfgv.num_envs - e0), HALF_WARPS);
#if defined(GO2_PHASE_TIMING)
  if (threadIdx.x == 0 && blockIdx.x == GO2_PHASE_TIMING) { go2_ph_count = 1; go2_ph_clock[0] = clock64(); }
#endif
  if (threadIdx.x < 32 * (GO2_HALF_ENVS / 2)) step_env<StepT<2, 1, 7>>(L, smem, X);
  else step_env<StepT<2, 2, 7>>(L, smem, X);
#if defined(GO2_PHASE_TIMING)
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == GO2_PHASE_TIMING) { const int k = go2_ph_count; if (k < 512) go2_ph_clock[k] = clock64(); go2_ph_count = k + 1; }
#endif
}

__global__ void __launch_bounds__(32 * WARPS_PER_CTA)
reset_kernel(const Go2EnvConfig* __restrict__ cfg, const Go2Model* __restrict__ mdl, const __grid_constant__ Go2EnvBuffers buf,
             const Go2StepParams* __restrict__ sp) {
  __shared__ WarpSmem smem[WARPS_PER_CTA];
  const int e0 = blockIdx.x * WARPS_PER_CTA;
  StepCtx X{cfg, mdl, &buf, sp, nullptr, cfg};
  Lane L;
  init_roles(L, threadIdx.x, 0, e0, min(WARPS_PER_CTA, cfg->num_envs - e0), WARPS_PER_CTA);
  if (!L.own) return;
  reset_env_initial<T1>(L, smem, X);
}

__global__ void __launch_bounds__(32 * WARPS_PER_CTA)
substeps_kernel(const Go2EnvConfig* __restrict__ cfg, const Go2Model* __restrict__ mdl, const __grid_constant__ Go2EnvBuffers buf,
                const float* __restrict__ tau, int n) {
  __shared__ WarpSmem smem[WARPS_PER_CTA];
  const int e0 = blockIdx.x * WARPS_PER_CTA;
  StepCtx X{cfg, mdl, &buf, nullptr, nullptr, cfg};
  Lane L;
  init_roles(L, threadIdx.x, 0, e0, min(WARPS_PER_CTA, cfg->num_envs - e0), WARPS_PER_CTA);
  if (!L.own) return;
  substeps_env<T1>(L, smem, X, tau, n);
}

// extras["episode"] (legged_robot.py:229-242): refreshed only when at least one env reset this step; then clear the sums
__global__ void finalize_kernel(const Go2EnvConfig* __restrict__ cfg, float* __restrict__ ep_accum, float* __restrict__ ep_stats,
                                const float* __restrict__ id_counts, const Go2StepParams* __restrict__ sp) {
  const int k = threadIdx.x, slot = sp->ep_slot;
  const float n_reset = ep_accum[GO2_NUM_REW + 10];
  __syncthreads();
  if (n_reset > 0.0f && ep_stats != nullptr) {
    float* st = ep_stats + (size_t)slot * GO2_EP_STATS;
    if (k < GO2_NUM_REW) st[k] = (float)((double)reinterpret_cast<const long long*>(ep_accum + GO2_EP_ACC_FIXED_OFF)[k] / (double)GO2_EP_FIXED_ONE) / n_reset / cfg->max_episode_length_s;
    else if (k == GO2_NUM_REW) st[k] = cfg->mesh_type == 0 ? 0.0f : ep_accum[k] / (float)cfg->num_envs;
    else if (k < GO2_NUM_REW + 10) st[k] = id_counts[k - GO2_NUM_REW - 1] > 0 ? ep_accum[k] / id_counts[k - GO2_NUM_REW - 1] : 0.0f;
    else if (k == GO2_NUM_REW + 10) st[k] = n_reset;
    else if (k == GO2_NUM_REW + 11) st[k] = 1.0f;
  } else if (ep_stats != nullptr && k < GO2_EP_STATS) {   // no reset in this step: the previous step's row is served again (header: GO2_EP_SLOTS)
    ep_stats[(size_t)slot * GO2_EP_STATS + k] = ep_stats[(size_t)((slot + GO2_EP_SLOTS - 1) % GO2_EP_SLOTS) * GO2_EP_STATS + k];
  }
  if (cfg->num_xrew > 0 && k < GO2_NUM_XREW) {   // the extra reward terms' episode means: same rules, their own accumulators / rows (Go2EnvConfig.ext_xrew_log)
    long long* xacc = GO2_EXT_PTR(long long*, cfg, ext_xrew_log);
    float* xst = reinterpret_cast<float*>(xacc + GO2_NUM_XREW);
    if (n_reset > 0.0f) xst[(size_t)slot * GO2_NUM_XREW + k] = (float)((double)xacc[k] / (double)GO2_EP_FIXED_ONE) / n_reset / cfg->max_episode_length_s;
    else xst[(size_t)slot * GO2_NUM_XREW + k] = xst[(size_t)((slot + GO2_EP_SLOTS - 1) % GO2_EP_SLOTS) * GO2_NUM_XREW + k];
    xacc[k] = 0;
  }
  __syncthreads();
  if (k < GO2_EP_STATS + 2) ep_accum[k] = 0.0f;
  if (k < GO2_NUM_REW) reinterpret_cast<long long*>(ep_accum + GO2_EP_ACC_FIXED_OFF)[k] = 0;
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
  Go2StepParams* d_sp = nullptr;  // staging slot of the host-parameter entry points (the kernels read the step parameters from device memory)
  cudaStream_t copy_stream = nullptr;   // device -> host copies of go2_env_step_host_begin
  cudaEvent_t ev_step = nullptr, ev_copied = nullptr;
  bool host_pending = false;
  int grid = 0;
  int step_mode = 8;              // thread map of the step kernel, see go2_env_set_step_mode
};

static int parse_step_mode(const char* m) {
  return !m ? 8 : !strcmp(m, "4") ? 0 : !strcmp(m, "8p") ? 1 : !strcmp(m, "P2") ? 2 : !strcmp(m, "P3") ? 3 : !strcmp(m, "Q4") ? 4 : !strcmp(m, "Q2") ? 5
       : !strcmp(m, "P2r") ? 6 : !strcmp(m, "P2b") ? 7 : !strcmp(m, "H14") ? 8 : -1;
}

namespace go2 {
template <int W, int MINB, int LOCKSTEP>
static int launch_wide(Go2Env* h, const float* actions, const Go2StepParams* sp, cudaStream_t st) {
  const int smem = W * (int)sizeof(WarpSmem);
  GO2_CUDA_OK(cudaFuncSetAttribute(step_kernel_wide<W, MINB, LOCKSTEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  step_kernel_wide<W, MINB, LOCKSTEP><<<(h->cfg.num_envs + W - 1) / W, 32 * W, smem, st>>>(h->d_cfg, h->d_mdl, h->buf, sp, actions);
  return 0;
}
template <int MINB, int ROT>
static int launch_packed(Go2Env* h, const float* actions, const Go2StepParams* sp, cudaStream_t st) {
  const int smem = 8 * (int)sizeof(WarpSmem);
  GO2_CUDA_OK(cudaFuncSetAttribute(step_kernel_packed<MINB, ROT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   // per device, cheap
  step_kernel_packed<MINB, ROT><<<(h->cfg.num_envs + 7) / 8, 256, smem, st>>>(h->d_cfg, h->d_mdl, h->buf, sp, actions);
  return 0;
}
template <int E, int MINB>
static int launch_quad(Go2Env* h, const float* actions, const Go2StepParams* sp, cudaStream_t st) {
  const int smem = E * (int)sizeof(WarpSmem);
  GO2_CUDA_OK(cudaFuncSetAttribute(step_kernel_quad<E, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  step_kernel_quad<E, MINB><<<(h->cfg.num_envs + E - 1) / E, 32 * E, smem, st>>>(h->d_cfg, h->d_mdl, h->buf, sp, actions);
  return 0;
}
static int launch_half(Go2Env* h, const float* actions, const Go2StepParams* sp, cudaStream_t st) {
  const int smem = GO2_HALF_ENVS * (int)sizeof(HalfSmem);
  GO2_CUDA_OK(cudaFuncSetAttribute(step_kernel_half, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  step_kernel_half<<<(h->cfg.num_envs + GO2_HALF_ENVS - 1) / GO2_HALF_ENVS, 32 * HALF_WARPS, smem, st>>>(h->d_cfg, h->d_mdl, h->buf, sp, actions, h->cfg);
  return 0;
}
}  // namespace go2

#if defined(GO2_PHASE_TIMING)
extern "C" int go2_debug_phase_clocks(long long* out, int cap) {   // tuning build only: clock64() after each phase barrier of the last step
  int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, go2::go2_ph_count, sizeof(int));
  n = n < cap ? n : cap;
  cudaMemcpyFromSymbol(out, go2::go2_ph_clock, sizeof(long long) * (n < 512 ? n : 512));
  return n;
}
#endif

extern "C" {

int go2_env_create(const Go2EnvConfig* cfg, const Go2Model* model, const Go2EnvBuffers* bufs, Go2Env** out) {
  if (!cfg || !model || !bufs || !out) return go2::set_error(1, "go2_env_create: null argument");
  if (cfg->num_envs <= 0) return go2::set_error(1, "go2_env_create: num_envs must be positive");
  if (cfg->heading_command && (!GO2_EXT_PTR(const void*, cfg, ext_stop_heading) || !GO2_EXT_PTR(const void*, cfg, ext_heading_ranges)))
    return go2::set_error(1, "go2_env_create: heading_command needs ext_stop_heading and ext_heading_ranges");
  if (cfg->num_xrew > 0 && (!GO2_EXT_PTR(const void*, cfg, ext_xrew_sums) || !GO2_EXT_PTR(const void*, cfg, ext_xrew_state) || !GO2_EXT_PTR(const void*, cfg, ext_xrew_log)))
    return go2::set_error(1, "go2_env_create: extra reward terms need ext_xrew_sums, ext_xrew_state and ext_xrew_log");
  if (cfg->turn_over && (!GO2_EXT_PTR(const void*, cfg, ext_turn_over_timer) || !GO2_EXT_PTR(const void*, cfg, ext_stop_heading)))
    return go2::set_error(1, "go2_env_create: turn_over needs ext_turn_over_timer and ext_stop_heading");
  if (cfg->control_type < 0 || cfg->control_type > 2) return go2::set_error(2, "go2_env_create: control_type must be 0 (P), 1 (V) or 2 (T)");
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
  GO2_CUDA_OK(cudaMalloc(&h->d_sp, sizeof(Go2StepParams)));
  GO2_CUDA_OK(cudaMemcpy(h->d_cfg, cfg, sizeof(Go2EnvConfig), cudaMemcpyHostToDevice));
  GO2_CUDA_OK(cudaMemcpy(h->d_mdl, model, sizeof(Go2Model), cudaMemcpyHostToDevice));
  std::vector<int32_t> ids(cfg->num_envs);
  GO2_CUDA_OK(cudaMemcpy(ids.data(), bufs->terrain_ids, sizeof(int32_t) * cfg->num_envs, cudaMemcpyDeviceToHost));
  float counts[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int e = 0; e < cfg->num_envs; ++e) if (ids[e] >= 0 && ids[e] < 9) counts[ids[e]] += 1.0f;
  GO2_CUDA_OK(cudaMemcpy(h->d_id_counts, counts, sizeof(counts), cudaMemcpyHostToDevice));
  GO2_CUDA_OK(cudaMemset(bufs->ep_accum, 0, sizeof(float) * GO2_EP_ACCUM_FLOATS));
  if (const char* m = getenv("GO2_STEP_MODE")) {
    if (parse_step_mode(m) < 0) { go2_env_destroy(h); return go2::set_error(1, "go2_env_create: unknown GO2_STEP_MODE"); }
    h->step_mode = parse_step_mode(m);
  }
  *out = h;
  return 0;
}

// Thread map of the step kernel (same results bit for bit; tuning / A-B aid).  Default "H14"; the GO2_STEP_MODE environment variable
// presets it at create time.
//   "H14": half-warp map, 14 envs per 288-thread CTA (7 WIDE warps, 16 threads per env + 2 dedicated LEGS warps), 2 CTAs/SM at 96 registers: 28 envs
//   resident per SM, 4096 envs in one wave (measured round 2: 114 us vs 162 us for "P2")
//   "P2": packed map, 8 envs per 256-thread CTA, 2 CTAs/SM (128 registers) · "P2r": the same with the leg warp rotated per SM through an atomic ticket
//   (measured round 2: 167 us vs 162 us — the ticket costs more than spreading the leg streams over the schedulers gains) · "P2b": leg warp = (block index / 148) mod 4, no ticket (co-resident CTAs of the first waves get different schedulers) · "P3": 3 CTAs/SM (80 registers)
//   "Q4" / "Q2": packed map with 4 / 2 envs per 128- / 64-thread CTA, 4 / 8 CTAs/SM (unmeasured: built after round 1's GPU budget was spent)
//   "8p": warp per env, 8 warps per CTA, CTA barrier at substep boundaries (the previous default: 203 us at 4096 envs)
//   "4" : warp per env, 4 warps per CTA, no barrier (the first kernel: 239 us)
int go2_env_set_step_mode(Go2Env* h, const char* mode) {
  if (!h || !mode || parse_step_mode(mode) < 0) return go2::set_error(1, "go2_env_set_step_mode: unknown mode (P2, P2r, P2b, P3, Q4, Q2, H14, 8p, 4)");
  h->step_mode = parse_step_mode(mode);
  return 0;
}

void go2_env_destroy(Go2Env* h) {
  if (!h) return;
  cudaFree(h->d_cfg); cudaFree(h->d_mdl); cudaFree(h->d_actions); cudaFree(h->d_id_counts); cudaFree(h->d_sp);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->ev_step) cudaEventDestroy(h->ev_step);
  if (h->ev_copied) cudaEventDestroy(h->ev_copied);
  delete h;
}

int go2_env_step(Go2Env* h, const float* actions, const Go2StepParams* sp, void* stream) {
  if (!h || !actions || !sp) return go2::set_error(1, "go2_env_step: null argument");
  // stream-ordered upload of the 72-byte parameter block (the source is staged before the call returns)
  GO2_CUDA_OK(cudaMemcpyAsync(h->d_sp, sp, sizeof(Go2StepParams), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return go2_env_step_dev(h, actions, h->d_sp, stream);
}

int go2_env_step_dev(Go2Env* h, const float* actions, const Go2StepParams* sp, void* stream) {
  if (!h || !actions || !sp) return go2::set_error(1, "go2_env_step_dev: null argument");
  if (h->host_pending) return go2::set_error(1, "go2_env_step: a host step opened by go2_env_step_host_begin is still pending (go2_env_step_host_end)");
  cudaStream_t st = (cudaStream_t)stream;
  const int mode = h->step_mode;
  if (mode == 0) go2::step_kernel<<<h->grid, 32 * go2::WARPS_PER_CTA, 0, st>>>(h->d_cfg, h->d_mdl, h->buf, sp, actions);
  else {
    int rc = mode == 1 ? go2::launch_wide<8, 2, 2>(h, actions, sp, st) : mode == 2 ? go2::launch_packed<2, 0>(h, actions, sp, st)
           : mode == 3 ? go2::launch_packed<3, 0>(h, actions, sp, st) : mode == 4 ? go2::launch_quad<4, 4>(h, actions, sp, st)
           : mode == 8 ? go2::launch_half(h, actions, sp, st) : mode == 6 ? go2::launch_packed<2, 1>(h, actions, sp, st) : mode == 7 ? go2::launch_packed<2, 2>(h, actions, sp, st) : go2::launch_quad<2, 8>(h, actions, sp, st);
    if (rc) return rc;
  }
  go2::count_launch();
  go2::finalize_kernel<<<1, 32, 0, st>>>(h->d_cfg, h->buf.ep_accum, h->buf.ep_stats, h->d_id_counts, sp);
  go2::count_launch();
  GO2_CUDA_OK(cudaGetLastError());
  return 0;
}

int go2_env_step_host_begin(Go2Env* h, const float* h_actions, const Go2StepParams* sp, float* h_obs, float* h_priv, float* h_rew,
                            uint8_t* h_reset, void* stream) {
  if (!h || !h_actions || !sp) return go2::set_error(1, "go2_env_step_host: null argument");
  if (h->host_pending) return go2::set_error(1, "go2_env_step_host_begin: the previous host step has not been closed by go2_env_step_host_end");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t N = (size_t)h->cfg.num_envs;
  if (!h->copy_stream) {
    GO2_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    GO2_CUDA_OK(cudaEventCreateWithFlags(&h->ev_step, cudaEventDisableTiming));
    GO2_CUDA_OK(cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming));
  }
  GO2_CUDA_OK(cudaMemcpyAsync(h->d_actions, h_actions, sizeof(float) * GO2_NUM_DOF * N, cudaMemcpyHostToDevice, st));
  int rc = go2_env_step(h, h->d_actions, sp, stream);
  if (rc) return rc;
  // the copies leave on the library's copy stream once the step has run: work the caller enqueues on `stream` after this call overlaps them
  GO2_CUDA_OK(cudaEventRecord(h->ev_step, st));
  GO2_CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->ev_step, 0));
  if (h_rew) GO2_CUDA_OK(cudaMemcpyAsync(h_rew, h->buf.rew_buf, sizeof(float) * N, cudaMemcpyDeviceToHost, h->copy_stream));
  if (h_reset) GO2_CUDA_OK(cudaMemcpyAsync(h_reset, h->buf.reset_buf, N, cudaMemcpyDeviceToHost, h->copy_stream));
  if (h_obs) GO2_CUDA_OK(cudaMemcpyAsync(h_obs, h->buf.obs_buf, sizeof(float) * GO2_NUM_OBS * N, cudaMemcpyDeviceToHost, h->copy_stream));
  if (h_priv) GO2_CUDA_OK(cudaMemcpyAsync(h_priv, h->buf.privileged_obs_buf, sizeof(float) * GO2_NUM_PRIV * N, cudaMemcpyDeviceToHost, h->copy_stream));
  GO2_CUDA_OK(cudaEventRecord(h->ev_copied, h->copy_stream));
  h->host_pending = true;
  return 0;
}

int go2_env_step_host_end(Go2Env* h) {
  if (!h) return go2::set_error(1, "go2_env_step_host_end: null argument");
  if (!h->host_pending) return 0;
  h->host_pending = false;
  GO2_CUDA_OK(cudaEventSynchronize(h->ev_copied));
  return 0;
}

int go2_env_step_host(Go2Env* h, const float* h_actions, const Go2StepParams* sp, float* h_obs, float* h_priv, float* h_rew,
                      uint8_t* h_reset, void* stream) {
  int rc = go2_env_step_host_begin(h, h_actions, sp, h_obs, h_priv, h_rew, h_reset, stream);
  return rc ? rc : go2_env_step_host_end(h);
}

int go2_env_reset_all(Go2Env* h, const Go2StepParams* sp, void* stream) {
  if (!h || !sp) return go2::set_error(1, "go2_env_reset_all: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  GO2_CUDA_OK(cudaMemcpyAsync(h->d_sp, sp, sizeof(Go2StepParams), cudaMemcpyHostToDevice, st));
  go2::reset_kernel<<<h->grid, 32 * go2::WARPS_PER_CTA, 0, st>>>(h->d_cfg, h->d_mdl, h->buf, h->d_sp);
  go2::count_launch();
  GO2_CUDA_OK(cudaMemsetAsync(h->buf.ep_accum, 0, sizeof(float) * GO2_EP_ACCUM_FLOATS, st));
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
