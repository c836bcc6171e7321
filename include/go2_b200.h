/* go2_b200.h — C ABI of the B200-native Go2 environment step + PPO trainer kernels.
 *
 * The reference (wty-yy/go2_rl_gym) is pure Python and has no FFI of its own; its only native
 * boundary is the Isaac Gym tensor API reached through `self.gym.*`
 * (legged_gym/envs/base/legged_robot.py:82-92,107-109,632,705,722,769-787).  This header is the
 * boundary a maintainer binds instead (ctypes stub in INTEGRATION.md): plain pointers and sizes,
 * caller-owned device memory, an int return code, no exceptions, no torch types.
 *
 * Every per-env quantity is its own row-major [num_envs, d] array ("structure of arrays of rows");
 * one warp owns one env and reads each row with lanes 0..d-1, i.e. one coalesced transaction per row.
 * The arrays are exactly the tensors the reference exposes as attributes of LeggedRobot
 * (legged_robot.py:765-859), so the Python shell wraps them zero-copy.
 */
#ifndef GO2_B200_H
#define GO2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GO2_NUM_DOF 12
#define GO2_NUM_DYN 13       /* dynamic bodies after merging fixed joints */
#define GO2_NUM_REPORT 19    /* rigid bodies Isaac Gym reports (dont_collapse links kept) */
#define GO2_NUM_COL 32       /* sphere collider samples, one per lane */
#define GO2_NUM_OBS 45       /* go2_env.py:26-32 */
#define GO2_NUM_PRIV 263     /* go2_env.py:36-47 */
#define GO2_NUM_HEIGHT 187   /* 17 x 11 scan, legged_robot_config.py:26-27 */
#define GO2_NUM_REW 14       /* active reward terms of GO2Cfg (go2_config.py:178-194) */
#define GO2_NUM_CMD 4
#define GO2_INERTIA_STRIDE 10 /* mass, com xyz, Ixx Iyy Izz Ixy Ixz Iyz (about COM, link frame) */
#define GO2_EP_STATS (GO2_NUM_REW + 12) /* rew means [14], terrain level mean, per-terrain-id means [9], n_reset, valid flag */
#define GO2_EP_ACC_FIXED_OFF ((GO2_EP_STATS + 2 + 1) / 2 * 2)
#define GO2_EP_ACCUM_FLOATS (GO2_EP_ACC_FIXED_OFF + 2 * GO2_NUM_REW)
#define GO2_EP_FIXED_ONE 1048576.0f /* 2^20 */
#define GO2_EP_SLOTS 64 /* rows of ep_stats: a ring indexed by Go2StepParams.ep_slot.  A step in which no env reset copies the previous slot's row
                           forward (the reference re-serves its stale extras["episode"], legged_robot.py:229-242 / on_policy_runner.py:145-146); the
                           valid flag stays 0 until the first reset, when the reference's extras has no "episode" key yet */

/* order of reward terms everywhere (scales, curriculum scales, episode_sums columns) */
enum Go2Reward {
  GO2_REW_TRACKING_LIN_VEL = 0, GO2_REW_TRACKING_ANG_VEL, GO2_REW_LIN_VEL_Z, GO2_REW_ANG_VEL_XY,
  GO2_REW_DOF_ACC, GO2_REW_DOF_POWER, GO2_REW_TORQUES, GO2_REW_CORRECT_BASE_HEIGHT, GO2_REW_ACTION_RATE,
  GO2_REW_ACTION_SMOOTHNESS, GO2_REW_COLLISION, GO2_REW_DOF_POS_LIMITS, GO2_REW_FEET_REGULATION,
  GO2_REW_HIP_TO_DEFAULT
};

/* Robot constants shared by all envs (from go2.urdf via tools/gen_go2_model.py). */
typedef struct Go2Model {
  float joint_origin[GO2_NUM_DOF][3]; /* joint frame origin in the parent body frame */
  int32_t joint_axis[GO2_NUM_DOF];    /* 0=x 1=y 2=z */
  float q_lower[GO2_NUM_DOF], q_upper[GO2_NUM_DOF];
  float effort[GO2_NUM_DOF];          /* torque_limits, legged_robot.py:370 */
  float vel_limit[GO2_NUM_DOF];
  float col_pos[GO2_NUM_COL][3];      /* sphere centre in its dynamic body's frame */
  float col_radius[GO2_NUM_COL];
  int32_t col_dyn[GO2_NUM_COL];       /* dynamic body 0..12 */
  int32_t col_report[GO2_NUM_COL];    /* reported body 0..18 (contact_forces row) */
  float foot_offset[4][3];            /* *_foot link origin in the calf frame */
} Go2Model;

/* Static configuration (LeggedRobotCfg / GO2Cfg values the step needs). */
typedef struct Go2EnvConfig {
  int32_t num_envs;          /* envs owned by this handle (this rank's shard) */
  int32_t env_offset;        /* global index of local env 0 (multi-GPU sharding keeps global ids for RNG) */
  uint32_t seed_lo, seed_hi; /* Philox key */
  /* sim / control (legged_robot_config.py:242-259, go2_config.py:77-85) */
  float sim_dt; int32_t decimation; float gravity_z;
  float kp[GO2_NUM_DOF], kd[GO2_NUM_DOF], default_dof_pos[GO2_NUM_DOF];
  float action_scale, clip_actions, clip_obs;
  /* domain randomisation switches and ranges (go2_config.py:39-75) */
  int32_t randomize_action_delay, randomize_motor_strength, randomize_motor_zero_offset, randomize_pd_gains;
  int32_t push_robots, add_noise;
  /* uniform ranges are stored as {lower, upper - lower}: the span is formed in double on the host exactly as
     torch_rand_float(lower, upper) forms it from Python floats, so device draws round like the reference's */
  float motor_strength_range[2], motor_zero_offset_range[2], kp_mult_range[2], kd_mult_range[2];
  int32_t push_interval; float max_push_vel_xy, max_push_ang_vel;
  /* contact / limit solver (physics spec) */
  int32_t solver_iters; float erp, limit_erp, contact_offset, max_depen_vel, bounce_threshold, penetration_slop;
  float terrain_friction, terrain_restitution;
  /* terrain (legged_robot_config.py:15-41) */
  int32_t mesh_type;         /* 0 plane, 1 heightfield (trimesh is served by the heightfield path) */
  int32_t hf_rows, hf_cols;  /* height_samples is [hf_rows(x), hf_cols(y)] int16 */
  float hscale, vscale, border;
  int32_t num_levels, num_types; /* terrain.num_rows, terrain.num_cols */
  float terrain_length;      /* cfg.terrain.terrain_length (8 m) */
  int32_t terrain_curriculum, move_down_by_accumulated_xy_command, custom_origins;
  /* commands (go2_config.py:97-146) */
  float resampling_time; int32_t dynamic_resample_commands;
  float limit_vel_prob; int32_t limit_vel_invert_when_continuous; float limit_ang_vel_at_zero_command_prob;
  /* episode */
  int32_t max_episode_length; float max_episode_length_s; float dt; /* dt = decimation * sim_dt */
  /* rewards: scales already multiplied by dt (legged_robot.py:920), GO2 order (enum Go2Reward) */
  float reward_scales[GO2_NUM_REW];
  float tracking_sigma, base_height_target;
  float soft_dof_limit_lo[GO2_NUM_DOF], soft_dof_limit_hi[GO2_NUM_DOF];
  int32_t dynamic_sigma; float ds_min_lin, ds_max_lin, ds_min_ang, ds_max_ang, ds_max_sigma[9];
  /* observations (go2_env.py:9-53) */
  float obs_scale_lin_vel, obs_scale_ang_vel, obs_scale_dof_pos, obs_scale_dof_vel, obs_scale_height;
  float noise_scale_vec[GO2_NUM_OBS];
  float height_points[GO2_NUM_HEIGHT][2];   /* body-frame xy of the scan grid (legged_robot.py:1172-1186) */
  float base_height_mask[GO2_NUM_HEIGHT];   /* legged_robot.py:790-795 */
  float num_base_height_points;
  float base_init_state[13];                /* pos, quat xyzw, lin vel, ang vel (legged_robot.py:1000) */
  /* relaxation of the Jacobi sweeps.  limit_relax = w > 0 (default 0.5): joint-limit rows step with w / (M^-1)_jj, the exact joint-space
     diagonal the mobility recursion already computes (convergent for w <= 0.5); limit_relax = 0: rows step with D_j (the first solver of
     round 1: over-relaxed when the parent link recoils, soft stops, divergent for > 4 sweeps — kept selectable for A/B only).
     contact_relax scales the contact rows' block step (default 0.7; 1 = plain mass splitting). */
  float limit_relax, contact_relax;
  /* state guard (default 1).  The base twist is clamped to the asset's max_linear_velocity / max_angular_velocity after every substep
     (legged_robot_config.py:131-132; PhysX clamps rigid-body velocities), and an env whose state (root 13, dof_pos 12, dof_vel 12) holds a
     non-finite value after the substeps is put into its initial pose at its origin with zero velocities / torques / contact forces and RESETS
     in this step (reset_buf = 1, time_out_buf = 0): one diverged env can never feed a NaN into the shared policy / value networks. */
  int32_t state_guard; float max_base_lin_vel, max_base_ang_vel;
  /* env switches outside the GO2 defaults (0 = the GO2 defaults):
     control_type 0 'P' / 1 'V' / 2 'T' (legged_robot.py:605-618); only_positive_rewards clips the summed reward at 0 (legged_robot.py:266-267) */
  int32_t control_type, only_positive_rewards;
  /* heading commands (legged_robot.py:411-419,468-472,547-548,581-582): commands[:,3] is a heading target, the yaw-rate command follows it
     every step unless stop_heading is set.  The two per-env arrays are caller-owned device memory like Go2EnvBuffers': stop_heading [N] uint8,
     heading_ranges [N,2] (lower, upper).  Their addresses are kept HERE, as two 32-bit halves each, so that neither the by-value Go2EnvBuffers
     kernel argument nor this struct's 4-byte alignment changes (a pointer member raises it to 8 and alters the default build's code). */
  int32_t heading_command, stop_heading_at_limit;
  uint32_t ext_stop_heading_lo, ext_stop_heading_hi, ext_heading_ranges_lo, ext_heading_ranges_hi;
  /* the reward functions that are INACTIVE in every registered go2 task (13 in legged_robot.py:1236-1441 + go2_env.py:62-68; enum Go2XReward):
     evaluated only when num_xrew > 0, i.e. when cfg.rewards.scales gives one of them a non-zero scale.  xrew_scales are multiplied by dt like
     reward_scales (0 = this term is off: its function is not called, so a stateful term does not advance its state — legged_robot.py:920-938).
     Caller-owned device arrays (addresses as 32-bit halves, see above): xrew_sums [N, GO2_NUM_XREW] episode sums; xrew_state [N, 12]
     (feet_air_time[4], last_contacts[4], last_contacts2[4] as 0 / 1); xrew_log = GO2_NUM_XREW int64 accumulators of the finished episodes'
     sums in 2^-20 fixed point followed by GO2_EP_SLOTS rows of GO2_NUM_XREW float means (the counterpart of ep_accum / ep_stats). */
  int32_t num_xrew;
  float xrew_scales[14];
  float soft_dof_vel_limit, soft_torque_limit, max_contact_force, min_legs_distance;
  uint32_t ext_xrew_sums_lo, ext_xrew_sums_hi, ext_xrew_state_lo, ext_xrew_state_hi, ext_xrew_log_lo, ext_xrew_log_hi;
  /* init_state.turn_over (legged_robot.py:114-115,174-175,257-265,586-590,642-691): robots are reset on their back / side with the given
     proportions (back, side, no flip), contact termination is off, while |roll| > turn_over_roll_threshold every reward term uses its
     turn_over scale instead of its normal one (to_scales / to_xscales, already x dt; terms of turn_over_scales need not be in scales), and
     commands stay zero until the per-env timer [N] (caller-owned, ext_turn_over_timer) has run down. */
  int32_t turn_over;
  float turn_over_proportions[3], turn_over_back_height[2], turn_over_side_height[2];   /* heights as {lower, span} like the other uniform ranges */
  float turn_over_zero_time_back, turn_over_zero_time_side, turn_over_roll_threshold;
  float to_scales[GO2_NUM_REW], to_xscales[14];
  uint32_t ext_turn_over_timer_lo, ext_turn_over_timer_hi;
} Go2EnvConfig;
#define GO2_EXT_PTR(type, cfg, name) ((type)(uintptr_t)(((uint64_t)(cfg)->name##_hi << 32) | (uint64_t)(cfg)->name##_lo))

#define GO2_NUM_XREW 14
enum Go2XReward {
  GO2_XREW_ORIENTATION = 0, GO2_XREW_BASE_HEIGHT, GO2_XREW_DOF_VEL, GO2_XREW_TERMINATION, GO2_XREW_DOF_VEL_LIMITS, GO2_XREW_TORQUE_LIMITS,
  GO2_XREW_FEET_AIR_TIME, GO2_XREW_STUMBLE, GO2_XREW_STAND_STILL, GO2_XREW_FEET_CONTACT_FORCES, GO2_XREW_SIMILAR_TO_DEFAULT, GO2_XREW_UPRIGHT,
  GO2_XREW_LEGS_DISTANCE, GO2_XREW_X_COMMAND_HIP_REGULAR
};
#define GO2_XREW_LOG_BYTES (GO2_NUM_XREW * 8 + GO2_EP_SLOTS * GO2_NUM_XREW * 4)

/* Per-step scalars the host derives from common_step_counter (curricula), no device sync involved. */
typedef struct Go2StepParams {
  uint32_t common_step_counter;   /* value AFTER this step's increment (legged_robot.py:112) */
  float reward_curriculum[GO2_NUM_REW]; /* legged_robot.py:144-168, 1.0 where no curriculum */
  float zero_command_proba;       /* legged_robot.py:556-557 */
  float max_lin_vel;              /* legged_robot.py:442 */
  int32_t ep_slot;                /* row of ep_stats this step writes */
  float xrew_curriculum[14];      /* reward_curriculum of the extra terms (enum Go2XReward), 1.0 where no curriculum */
} Go2StepParams;

/* Caller-owned arrays (device pointers for the CUDA library, host pointers for the oracle). */
typedef struct Go2EnvBuffers {
  /* simulation state */
  float* root_states;      /* [N,13] pos, quat xyzw, world lin vel, world ang vel */
  float* dof_pos;          /* [N,12] */
  float* dof_vel;          /* [N,12] */
  float* torques;          /* [N,12] last substep, after motor strength */
  float* contact_forces;   /* [N,19,3] world, last substep */
  float* feet_pos;         /* [N,4,3] world (rigid_body_states[:, feet, 0:3]) */
  float* feet_vel;         /* [N,4,3] world linear velocity */
  /* policy interface */
  float* actions;          /* [N,12] clipped copy of the step input */
  float* last_actions;     /* [N,12] */
  float* last_last_actions;/* [N,12] */
  float* last_dof_vel;     /* [N,12] */
  float* obs_buf;          /* [N,45] */
  float* privileged_obs_buf;/* [N,263] */
  float* rew_buf;          /* [N] */
  uint8_t* reset_buf;      /* [N] */
  uint8_t* time_out_buf;   /* [N] */
  int32_t* episode_length_buf; /* [N] */
  /* derived base-frame quantities (legged_robot.py:119-125) */
  float* base_lin_vel;     /* [N,3] */
  float* base_ang_vel;     /* [N,3] */
  float* projected_gravity;/* [N,3] */
  float* measured_heights; /* [N,187] */
  /* commands */
  float* commands;         /* [N,4] */
  float* commands_resampling_step; /* [N] */
  float* commands_xy_accumulation; /* [N,2] */
  uint8_t* last_is_limit_vel;      /* [N] */
  float* env_command_ranges;       /* [N,6] x lo/hi, y lo/hi, yaw lo/hi */
  /* terrain curriculum */
  int32_t* terrain_levels; /* [N] */
  int32_t* terrain_types;  /* [N] */
  int32_t* terrain_ids;    /* [N] 0..8 */
  float* env_origins;      /* [N,3] */
  float* max_move_distance;/* [N] */
  const float* terrain_origins;   /* [num_levels,num_types,3] */
  const int16_t* height_samples;  /* [hf_rows,hf_cols] */
  /* per-env randomised properties */
  float* motor_strengths;  /* [N,12] */
  float* motor_zero_offsets;/* [N,12] */
  float* p_gains_multiplier;/* [N,12] */
  float* d_gains_multiplier;/* [N,12] */
  const float* friction_coeffs; /* [N] robot shape friction */
  const float* restitutions;    /* [N] */
  const float* body_inertia;    /* [N,13,10] composite inertials of the dynamic bodies */
  /* logging */
  float* episode_sums;     /* [N,14] */
  float* ep_stats;         /* [ep_slots, GO2_EP_STATS] */
  float* ep_accum;         /* [GO2_EP_ACCUM_FLOATS] scratch for the cross-env sums (zeroed by the step): GO2_EP_STATS + 2 floats (terrain-level sums and
                              counters: integer-valued, so their float atomics are order-independent), then, 8-byte aligned at float index
                              GO2_EP_ACC_FIXED_OFF, GO2_NUM_REW int64 accumulators of the finished episodes' reward sums in 2^-20 fixed point —
                              integer atomics, so the logged means are bit-reproducible run to run whatever order the CTAs finish in */
} Go2EnvBuffers;

/* ---- environment (CUDA library: libgo2b200.so) ------------------------------------------------ */
typedef struct Go2Env Go2Env;

/* Replaces gym.create_sim/prepare_sim + acquire_*_tensor (legged_robot.py:292-310,769-787). 0 on success. */
int go2_env_create(const Go2EnvConfig* cfg, const Go2Model* model, const Go2EnvBuffers* bufs, Go2Env** out);
void go2_env_destroy(Go2Env* env);
/* Thread map of the fused step kernel: "H14" (default; 14 envs per 9-warp CTA on half-warps, dedicated leg warps), "P2" (8 envs packed per CTA), "P3", "Q4" (4 envs packed per CTA), "8p", "4" (warp per env).
 * Same results (bit for bit in the host emulation; each map is its own kernel instantiation on the GPU). */
int go2_env_set_step_mode(Go2Env* env, const char* mode);
/* Replaces LeggedRobot.step (legged_robot.py:60-100): the fused kernel. actions: device [N,12]. */
int go2_env_step(Go2Env* env, const float* actions, const Go2StepParams* sp, void* cuda_stream);
/* Same, with the step parameters already in DEVICE memory (d_sp): nothing in the launch depends on host values, so a whole rollout
 * (OnPolicyRunner.learn's 24-step loop, on_policy_runner.py:136-153) can be captured in one CUDA graph over an array of parameter blocks. */
int go2_env_step_dev(Go2Env* env, const float* actions, const Go2StepParams* d_sp, void* cuda_stream);
/* Same, HOST buffers: H2D of actions, kernel, D2H of obs/priv/rew/reset inside the call (bench e2e). */
int go2_env_step_host(Go2Env* env, const float* h_actions, const Go2StepParams* sp, float* h_obs, float* h_priv,
                      float* h_rew, uint8_t* h_reset, void* cuda_stream);

/* The host-buffer step in two halves, for a caller that has device work of its own to enqueue between a step and the next one (the runner's transition
   bookkeeping and the next policy inference): _begin uploads the actions, launches the step on `stream` and starts the device -> host copies on the
   library's own copy stream; it returns without waiting.  _end blocks until the four host buffers of that step are filled.  Every _begin must be closed
   by one _end before the next _begin / any other step call on the handle (the next step overwrites the device buffers the copies read); between the two
   calls work enqueued on `stream` may READ the env's device buffers.  go2_env_step_host == _begin + _end. */
int go2_env_step_host_begin(Go2Env* env, const float* h_actions, const Go2StepParams* sp, float* h_obs, float* h_priv,
                            float* h_rew, uint8_t* h_reset, void* stream);
int go2_env_step_host_end(Go2Env* env);
/* Replaces reset_idx(arange(N)) at construction (base_task.py:82-86 calls reset_idx then a zero-action step). */
int go2_env_reset_all(Go2Env* env, const Go2StepParams* sp, void* cuda_stream);
/* n bare physics substeps under given joint torques tau [N,12] (parity tests of the dynamics in isolation). */
int go2_env_substeps(Go2Env* env, const float* tau, int n_substeps, void* cuda_stream);
const char* go2_last_error(void);
long long go2_kernel_launch_count(void); /* kernels launched by this library since load */


/* ---- PPO trainer kernels (rsl_rl restated; csrc/rl_kernels.cu, csrc/gemm_tc.cu) ------------------------------
 * All matrices row-major fp32 with explicit leading dimensions; every pointer is device memory. */

/* Y[M,N] = act(X[M,K] W[N,K]^T + b[N]); act: 0 none, 1 ELU   (nn.Linear + nn.ELU, actor_critic.py:58-79) */
int go2_linear_forward_simt(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy, float* Yt, int ldyt, int M, int N, int K, int act, void* stream);
/* same contraction on the tensor cores (tcgen05.mma kind::tf32, TMA-fed; csrc/gemm_tc.cu); Yt (optional) receives Y^T [N,M] */
int go2_linear_forward_tc(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy, float* Yt, int ldyt, int M, int N, int K, int act, void* stream);
/* dX[M,K] = dY[M,N] W[N,K], multiplied by ELU'(act_in) when act_in (the layer input's post-activation) is given */
int go2_linear_dgrad_simt(const float* dY, int lddy, const float* W, int ldw, const float* act_in, int ldact, float* dX, int lddx, float* dXt, int lddxt, int M, int N, int K, void* stream);
/* tensor-core dgrad: Wt = W^T stored [K,N]; act_in_t = the activation transposed [K,M] (either form may be NULL); dXt (optional) receives dX^T [K,M] */
int go2_linear_dgrad_tc(const float* dZ, int lddz, const float* Wt, int ldwt, const float* act_in, int ldact, const float* act_in_t, int ldact_t, float* dX, int lddx, float* dXt, int lddxt, int M, int N, int K, void* stream);
/* dW[N,K] = dY^T X, db[N] = column sums of dY; deterministic split over the M rows through `workspace` */
int go2_linear_wgrad_simt(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, float* db, int M, int N, int K, float* workspace, long workspace_floats, void* stream);
/* tensor-core wgrad from the transposed copies dZt [N,M], Xt [K,M] (contraction over the M rows is then K-major for both operands);
 * db != NULL: Xt has one extra row of ones (row K) and db[N] receives the bias gradient from the same contraction */
int go2_linear_wgrad_tc(const float* dZt, int lddzt, const float* Xt, int ldxt, float* dW, int lddw, float* db, int M, int N, int K, float* workspace, long workspace_floats, void* stream);
/* The same weight gradient straight from the ROW-MAJOR tensors (MN-major tf32 operands through TMA's 128B / 32-byte-atom swizzle):
 * dW[N,K] = dZ[M,N]^T X[M,K]; with db != NULL X has a column of ones at column K (ldx > K) and db[N] = column sums of dZ.
 * Replaces autograd's weight gradient of nn.Linear (rsl_rl/algorithms/ppo.py:175 loss.backward()).  workspace is required. */
int go2_linear_wgrad_tc_rm(const float* dZ, int lddz, const float* X, int ldx, float* dW, int lddw, float* db, int M, int N, int K,
                           float* workspace, long workspace_floats, void* stream);
/* Linear layers with a narrow output (N <= 16, K <= 128; the actor's 12-wide head, actor_critic.py:66): streaming fp32 kernels, one warp per
 * row, warp-shuffle reductions.  forward: Y = X W^T + b; dgrad: dX = (dY W) * ELU'(act_in) (act_in optional); wgrad: dW = dY^T X, db = sum(dY)
 * (deterministic two-stage reduction, workspace >= 296 * (N K + N) floats for full parallelism). */
int go2_linear_forward_smalln(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy, int M, int N, int K, void* stream);
int go2_linear_dgrad_smalln(const float* dY, int lddy, const float* W, int ldw, const float* act_in, int ldact, float* dX, int lddx, int M, int N, int K,
                            void* stream);
int go2_linear_wgrad_smalln(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, float* db, int M, int N, int K, float* workspace,
                            long workspace_floats, void* stream);
/* weight / bias gradient of a 1-wide Linear (the critic's head, actor_critic.py:79): dW[K] = dY^T X, db = sum(dY); one streaming pass over X,
 * deterministic two-stage reduction.  workspace >= 297 * (K + 1) floats for full parallelism. */
int go2_linear_wgrad_rank1(const float* dY, int lddy, const float* X, int ldx, float* dW, float* db, int M, int K, float* workspace,
                           long workspace_floats, void* stream);
/* n <= 8 small pitch-copies (transpose[j] = 0) / transposes (1) in one launch: dst_j[cols_j or rows_j ...] from src_j[rows_j, cols_j].  The operand
 * copies an MLP refreshes after each optimiser step (padded first-layer weight, W^T for dgrad).  The arrays are HOST arrays of device pointers / sizes. */
int go2_refresh_weights(int n, const float* const* src, const int* ldin, float* const* dst, const int* ldout, const int* rows, const int* cols,
                        const int* transpose, void* stream);
/* out[cols,rows] = in[rows,cols]^T */
int go2_transpose(const float* in, int ldin, float* out, int ldout, int rows, int cols, void* stream);
/* Multiply precision of the tensor-core GEMMs (process-wide).  3 (default): 3xTF32 — every fp32 operand is split in shared memory into
   hi = rn_tf32(a) and lo = a - hi and the product accumulates lo_a hi_b + hi_a lo_b + hi_a hi_b in fp32 (TMEM): fp32-class products
   (relative error ~1e-6) matching the reference's sgemm (rsl_rl/algorithms/ppo.py:120-187 runs torch fp32).  1: one tf32 pass (10-bit
   mantissa products; round 1's kernel, kept for A/B; also selected by the environment variable GO2_GEMM=tf32). */
int go2_gemm_set_passes(int passes);
int go2_gemm_get_passes(void);
/* Where the 3xTF32 kernel takes its hi operand from.  0 (default): the raw fp32 word — tcgen05 kind::tf32 ignores the 13 low mantissa bits, so the
   hardware's hi is trunc_tf32(a) and the splitter warps only write lo = rn_tf32(a - trunc_tf32(a)).  1: the staged operand is rewritten with
   rn_tf32(a) and lo = a - hi (A/B check of the assumption; also GO2_GEMM_SPLIT=rewrite). */
int go2_gemm_set_split(int rewrite);
/* 1 (default): 3xTF32 GEMMs with more than 128 rows run on CTA pairs (tcgen05.mma.cta_group::2, 256-row tiles, each SM stages half of B);
   0: the one-CTA persistent kernel everywhere (A/B aid; also GO2_GEMM_PAIR=0).  Same arithmetic, same results up to the summation split. */
int go2_gemm_set_pair(int on);
/* 1 (default): the persistent GEMMs are launched with programmatic stream serialization (their prologue runs under the tail of the previous kernel,
   griddepcontrol.wait orders the memory traffic); 0: plain stream order (A/B aid; also GO2_GEMM_PDL=0). */
int go2_gemm_set_pdl(int on);
/* Profiling aid: counters != NULL makes every persistent tensor-core GEMM launch write, per CTA b, 24 int64 counters at counters[24 b ...]
   (device memory, >= 24 x SM count; 16..19: %globaltimer ns at kernel entry, after the prologue, after griddepcontrol.wait, at exit): 0 producer total, 1 producer waiting for a free stage; 2 MMA thread total, 3 .. waiting for a drained accumulator,
   4 .. for TMA bytes, 5 .. for the lo slot, 6 stages processed; 7 splitter waiting for TMA bytes, 8 .. for a free lo slot, 9 splitter busy (incl. 8);
   10/11 and 12/13 epilogue group 0 / 1 total and waiting for an accumulator, 14 / 15 .. for the ELU' operand (dgrad).  NULL (default) disables it. */
int go2_gemm_set_debug(long long* counters);
/* One-shot SUM all-reduce over NVLink peer memory (csrc/dist_kernels.cu), the data-parallel trainer's per-optimiser-step exchange — replaces the
   reference's single-process optimizer.step() boundary (rsl_rl/algorithms/ppo.py:183-185) when the envs are sharded over GPUs:
   out[off .. off+n) = sum_r peer_data[r][off .. off+n), same summation order on every rank.  peer_data / peer_flags are HOST arrays of `world`
   device pointers (this process's mappings of every rank's symmetric buffer / 16-word zeroed flag block, index = rank); ctr = 2 zeroed uint32
   words of this rank.  off, n multiples of 4 floats.  Collective: every rank calls it with the same arguments in the same order; capturable. */
int go2_allreduce_p2p(const float* const* peer_data, uint32_t* const* peer_flags, float* out, long off, long n, int rank, int world, uint32_t* ctr,
                      void* stream);
/* The same exchange with the result buffers symmetric as well (peer_out: HOST array of `world` device pointers, peer_out[rank] == out): with
   world >= 4 it runs two-shot — rank r sums slice r only (1/world of the peer reads) and stores it into every rank's result. */
int go2_allreduce_p2p2(const float* const* peer_data, uint32_t* const* peer_flags, float* const* peer_out, float* out, long off, long n, int rank,
                       int world, uint32_t* ctr, void* stream);
/* db[N] = column sums of dY[M,N] */
int go2_colsum(const float* dY, int lddy, float* db, int M, int N, float* scratch /* >= 64*N floats */, void* stream);
/* PPO.act tail (ppo.py:94-101): actions = mu + std z (Philox normal), log-prob, mu/sigma rows of the transition */
int go2_sample_actions(const float* mu, const float* std_param, float* actions, float* logp, float* mu_out, float* sigma_out, int N, int A, uint64_t seed, uint32_t step, int env_offset, void* stream);
/* same, the step counter read from device memory (rollout captured in a CUDA graph) */
int go2_sample_actions_dev(const float* mu, const float* std_param, float* actions, float* logp, float* mu_out, float* sigma_out, int N, int A, uint64_t seed, const uint32_t* d_step, int env_offset, void* stream);
/* PPO.process_env_step (ppo.py:104-111): rewards += gamma V time_out; rows of the transition */
int go2_process_env_step(const float* rew, const uint8_t* dones, const uint8_t* time_outs, const float* values, float* rew_out, uint8_t* dones_out, int N, float gamma, const int64_t* perm, void* stream);
/* RolloutStorage.compute_returns (rollout_storage.py:123-134), all [T,N]; stats[2] (double) receives sum / sum of squares of adv */
int go2_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values, float* returns, float* advantages, int T, int N, float gamma, float lam, double* stats, void* stream);
/* advantages = (adv - mean) / (std + 1e-8) with the (possibly all-reduced) stats and global count (rollout_storage.py:136-137) */
int go2_adv_normalize(float* advantages, long n, const double* stats, double global_count, void* stream);
/* mini-batch gather (rollout_storage.py:173-181): dst[i, :width] = src[idx[i], :width] (idx NULL = identity), zero padded to ldd;
 * dst_t (optional) receives the transposed copy [width, n] */
int go2_gather_rows(const float* src, int width, const int64_t* idx, float* dst, int ldd, float* dst_t, long n, void* stream);
/* PPO losses forward + backward (ppo.py:131-171). scal[20]: sum KL, sum surrogate (rows < split), sum value loss, sum entropy, d/d std[<=15],
 * scal[19] = sum surrogate of rows >= split.  CTS: rows [0,split) are teacher samples weighted inv_count_a, the rest student samples weighted
 * inv_count_b (cts.py / moe_cts.py:160-168); plain PPO passes split = M and inv_count_a = inv_count */
int go2_ppo_loss(const float* mu, const float* std_param, const float* value, const float* actions, const float* old_logp, const float* adv,
                 const float* target_values, const float* returns, const float* old_mu, const float* old_sigma, float* dmu, float* dmu_t, float* dvalue,
                 float* scal, int M, int A, float clip, float value_coef, float entropy_coef, int use_clipped_value_loss, float inv_count,
                 int split, float inv_count_a, float inv_count_b, void* stream);
/* KL-adaptive learning rate on the device (ppo.py:139-151); log_out[5]: += value loss, += surrogate loss, = kl, = lr, += entropy */
int go2_kl_adaptive_lr(const float* scal, float count, float desired_kl, float* lr_state, float* log_out, float count_a, float count_b, void* stream);
/* clip_grad_norm_ + Adam.step on a flat parameter vector (ppo.py:176-177); lr_state[4] = {lr, step count, 1-beta1^t, sqrt(1-beta2^t)} lives on the
 * device (the call increments the step), scratch >= 1025 floats */
int go2_adam_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n, float max_grad_norm, float* lr_state, float grad_scale, float* scratch, void* stream);

/* ---- Concurrent Teacher-Student / MoE pieces (csrc/cts_kernels.cu) ------------------------------------------------------ */
/* out[i] = [a[i, :wa] | b[i, :wb]] zero padded to ld; out_t (optional) transposed copy   (actor_critic_moe_cts.py:121,140) */
int go2_concat2(const float* a, int wa, int lda, const float* b, int wb, int ldb, float* out, int ld, float* out_t, long n, void* stream);
/* L2Norm forward / backward (modules/utils.py:24-30) */
int go2_l2norm_forward(const float* x, int ldx, float* y, int ldy, float* norm, long n, int d, void* stream);
int go2_l2norm_backward(const float* dy, int lddy, const float* y, int ldy, const float* norm, float* dx, int lddx, float* dx_t, long n, int d, void* stream);
/* softmax gate + weighted sum of the expert outputs (modules/utils.py:122-126) and its backward incl. the load-balance term (moe_cts.py:210-216) */
/* Expert layer of the mixture modules = Conv1d(E*H -> E*D, kernel 1, groups = E) (rsl_rl/modules/utils.py:83-88): E block-diagonal Linear(H -> D) over the
   backbone's feature blocks, all experts in ONE launch per direction (fp32 FMAs on the CUDA cores; W is [E*D, H] row-major, b [E*D]):
   forward Y[m, eD+d] = b + sum_h X[m, eH+h] W[eD+d, h]; dgrad dX[m, eH+h] = ELU'(act[m, eH+h]) sum_d dY[m, eD+d] W[eD+d, h] (act may be null);
   wgrad dW[eD+d, h] = sum_m dY[m, eD+d] X[m, eH+h], row chunks summed in a fixed order (workspace: go2_grouped_linear_wgrad_workspace floats). */
int go2_grouped_linear_forward(const float* X, long ldx, const float* W, const float* b, float* Y, long ldy, long M, int E, int D, int H, void* stream);
int go2_grouped_linear_dgrad(const float* dY, long lddy, const float* W, const float* act, long ldact, float* dX, long lddx, long M, int E, int D, int H,
                             void* stream);
long go2_grouped_linear_wgrad_workspace(long M, int E, int D, int H);
int go2_grouped_linear_wgrad(const float* dY, long lddy, const float* X, long ldx, float* dW, long M, int E, int D, int H, float* workspace,
                             long workspace_floats, void* stream);
int go2_moe_combine_forward(const float* logits, const float* expert_out, float* gates, float* pre, long n, int E, int D, void* stream);
int go2_moe_combine_backward(const float* dpre, const float* gates, const float* expert_out, float* usage, float lb_coef, float* dexpert_out,
                             float* dexpert_out_t, float* dlogits, float* dlogits_t, long n, int E, int D, void* stream);
/* go2_moe_combine_backward in two halves for the env-sharded trainer (the load-balance term of moe_cts.py:211-214 is a function of the WHOLE
   mini-batch's mean gate usage): usage[e] = scale * mean over this rank's n rows (scale = 1 / world_size; the caller sums it over the ranks),
   then the backward with the given (global) usage. */
int go2_gate_usage(const float* gates, float* usage, long n, int E, float scale, void* stream);
int go2_moe_combine_backward_given_usage(const float* dpre, const float* gates, const float* expert_out, const float* usage, float lb_coef,
                                         float* dexpert_out, float* dexpert_out_t, float* dlogits, float* dlogits_t, long n, int E, int D, void* stream);
/* latent reconstruction loss mean((teacher - student)^2) and its gradient (moe_cts.py:205-207); acc[1] receives the sum of squares */
int go2_latent_loss(const float* student, const float* teacher, float* dstudent, float* acc, long n, int d, void* stream);
/* log[0] += latent loss, log[1] += load-balance loss from usage[E] (NULL -> 0) */
int go2_cts_log(const float* acc, const float* usage, float* log, long count, int E, void* stream);
/* rolling observation history [n, H, d]: zero on done, shift, append (on_policy_runner_cts.py:155-156) */
int go2_history_update(float* history, const float* obs, const uint8_t* dones, long n, int H, int d, void* stream);
int go2_gather_u8(const uint8_t* src, const int64_t* perm, uint8_t* out, long n, void* stream);

/* ---- Multiplicative-compositional-policy actor head (csrc/mcp_kernels.cu; go2_mcp_cts) ------------------------------------------ */
/* ActorMCP.forward (rsl_rl/modules/actor_critic_mcp_cts.py:220-247).  expert_out [n, E*2A]: expert e = [mu_e (A) | log_std_e (A)];
 * gates = sigmoid(logits [n,E]); var_e = exp(2 clamp(log_std_e, -5, 2)) + 1e-9; sigma^2 = 1 / (sum_e gates_e / var_e + 1e-9);
 * mu = sigma^2 sum_e gates_e mu_e / var_e.  All matrices dense.  E <= 16, A <= 16. */
int go2_mcp_compose_forward(const float* expert_out, const float* logits, float* gates, float* mu, float* sigma, long n, int E, int A, void* stream);
/* its backward: (d loss / d mu, d loss / d sigma) [n,A] -> d loss / d expert_out [n, E*2A], d loss / d logits [n,E] (through the sigmoid) */
int go2_mcp_compose_backward(const float* dmu, const float* dsigma, const float* expert_out, const float* gates, float* dexpert_out, float* dlogits, long n,
                             int E, int A, void* stream);
/* go2_sample_actions / go2_sample_actions_dev for a state-dependent sigma [N,A] (Normal(mean, std), actor_critic_mcp_cts.py:146-149);
 * the step counter is read from d_step when it is not NULL.  Same Philox draws as go2_sample_actions. */
int go2_sample_actions_sigma(const float* mu, const float* sigma, float* actions, float* logp, float* mu_out, float* sigma_out, int N, int A, uint64_t seed,
                             uint32_t step, const uint32_t* d_step, int env_offset, void* stream);
/* go2_ppo_loss for a state-dependent sigma [M,A] (rsl_rl/algorithms/mcp_cts.py:133-181): d loss / d sigma goes to dsigma [M,A] (entropy term
 * included); scal[0..3] and scal[19] as in go2_ppo_loss, scal[4..18] = 0 */
int go2_ppo_loss_sigma(const float* mu, const float* sigma, const float* value, const float* actions, const float* old_logp, const float* adv,
                       const float* target_values, const float* returns, const float* old_mu, const float* old_sigma, float* dmu, float* dsigma,
                       float* dvalue, float* scal, int M, int A, float clip, float value_coef, float entropy_coef, int use_clipped_value_loss,
                       float inv_count, int split, float inv_count_a, float inv_count_b, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GO2_B200_H */
