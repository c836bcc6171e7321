"""LeggedRobot — the VecEnv the runners drive, backed by ONE fused sm_100a kernel per step.

Drop-in for `legged_gym.envs.base.legged_robot.LeggedRobot` / `base_task.BaseTask`
(reference: legged_gym/envs/base/legged_robot.py:25-142, base_task.py:11-86; contract: rsl_rl/env/vec_env.py:36-59):
same constructor signature, same public buffers (`obs_buf, privileged_obs_buf, rew_buf, reset_buf, episode_length_buf,
time_out_buf, extras, commands, root_states, dof_pos, dof_vel, torques, contact_forces, base_lin_vel, ...`), same
`step / reset / get_observations / get_privileged_observations / update_reward_curriculum` methods and the attributes
train.py and the runners reach into.  Every buffer is a zero-copy torch view of the device rows
the kernel reads and writes (include/go2_b200.h: Go2EnvBuffers).

There is no CPU path: construction raises if libgo2b200.so is missing or no CUDA device is present.
"""
import ctypes

import numpy as np
import torch

from ... import _abi
from ..env_arrays import EnvArrays, EP_SLOTS, class_to_dict  # noqa: F401
from ...utils.terrain import TERRAIN_NAMES


class LeggedRobot:
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True,
                 env_offset=0, num_envs_global=None):
        self.cfg = cfg
        self.sim_params = sim_params
        self.physics_engine = physics_engine
        self.sim_device = sim_device
        self.headless = headless
        self.height_samples = None
        self.debug_viz = False
        self.init_done = False
        dev = torch.device(sim_device if str(sim_device).startswith("cuda") else "cuda:0")
        if not torch.cuda.is_available():
            raise RuntimeError("go2_rl_gym_b200 needs a CUDA device (B200); there is no CPU fallback for the env step")
        self._lib = _abi.load_library()
        self.device = str(dev)
        torch.cuda.set_device(dev)
        self.num_envs = cfg.env.num_envs
        self.num_obs = cfg.env.num_observations
        self.num_privileged_obs = cfg.env.num_privileged_obs
        self.num_actions = cfg.env.num_actions
        if (self.num_obs, self.num_privileged_obs, self.num_actions) != (_abi.NUM_OBS, _abi.NUM_PRIV, _abi.NUM_DOF):
            raise NotImplementedError("the fused kernel is specialised for the Go2 45 / 263 / 12 layout (go2_env.py:23-53)")
        A = EnvArrays(cfg, dev, num_envs=self.num_envs, env_offset=env_offset, num_envs_global=num_envs_global,
                      seed=getattr(cfg, "seed", 1))
        self._A = A
        T = A.tensors
        # ---- public buffers (same names as the reference's attributes)
        for name in ("root_states", "dof_pos", "dof_vel", "torques", "contact_forces", "actions", "last_actions",
                     "last_last_actions", "last_dof_vel", "obs_buf", "privileged_obs_buf", "rew_buf", "base_lin_vel",
                     "base_ang_vel", "projected_gravity", "measured_heights", "commands", "commands_resampling_step",
                     "commands_xy_accumulation", "terrain_levels", "terrain_types", "terrain_ids", "env_origins",
                     "max_move_distance", "motor_strengths", "motor_zero_offsets", "p_gains_multiplier",
                     "d_gains_multiplier", "feet_pos", "feet_vel"):
            setattr(self, name, T[name])
        self.reset_buf = T["reset_buf"].view(torch.bool)
        self.time_out_buf = T["time_out_buf"].view(torch.bool)
        self.last_is_limit_vel = T["last_is_limit_vel"].view(torch.bool)
        self._episode_length_buf = T["episode_length_buf"]
        self.base_quat = self.root_states[:, 3:7]
        self.base_pos = self.root_states[:, 0:3]
        self.friction_coeffs = T["friction_coeffs"]
        self.episode_sums = {n: T["episode_sums"][:, k] for k, n in enumerate(_abi.REWARD_NAMES) if n in A.reward_names}
        self.episode_sums.update({n: T["xrew_sums"][:, k] for k, n in enumerate(_abi.XREWARD_NAMES) if n in A.reward_names})
        self.turn_over_timer = T["turn_over_timer"]          # legged_robot.py:833
        self.reward_turn_over_scales = A.reward_turn_over_scales
        self.feet_air_time = T["xrew_state"][:, 0:4]            # state of the feet_air_time / base_height reward terms (legged_robot.py:817-818,1248-1251)
        self.last_contacts, self.last_contacts2 = T["xrew_state"][:, 4:8], T["xrew_state"][:, 8:12]
        self.env_command_ranges = {"lin_vel_x": T["env_command_ranges"][:, 0:2], "lin_vel_y": T["env_command_ranges"][:, 2:4],
                                   "ang_vel_yaw": T["env_command_ranges"][:, 4:6], "heading": T["heading_ranges"]}
        self.stop_heading = T["stop_heading"].view(torch.bool)
        self.dt = A.dt
        self.max_episode_length_s = A.max_episode_length_s
        self.max_episode_length = A.max_episode_length
        self.obs_scales = cfg.normalization.obs_scales
        self.reward_scales = A.reward_scales
        self.command_ranges = A.command_ranges
        self.custom_origins = A.custom_origins
        if not A.plane:
            self.terrain = A.terrain
            self.height_samples = T["height_samples"]
            self.terrain_origins = T["terrain_origins"]
            self.max_terrain_level = cfg.terrain.num_rows
        self.num_dof = self.num_dofs = _abi.NUM_DOF
        self.num_bodies = _abi.NUM_REPORT
        self.dof_names = A.model_json["dof_names"]
        body_names = A.model_json["report_bodies"]
        idx = lambda key: torch.tensor([i for i, n in enumerate(body_names) if key in n], dtype=torch.long, device=dev)
        self.feet_indices = idx(cfg.asset.foot_name)
        self.penalised_contact_indices = torch.cat([idx(k) for k in cfg.asset.penalize_contacts_on])
        self.termination_contact_indices = torch.cat([idx(k) for k in cfg.asset.terminate_after_contacts_on])
        self.default_dof_pos = torch.from_numpy(A.default_dof_pos_np).to(dev).unsqueeze(0)
        self.torque_limits = torch.tensor(list(A.model.effort), device=dev)
        self.noise_scale_vec = torch.from_numpy(A.noise_scale_vec_np).to(dev)
        self.add_noise = cfg.noise.add_noise
        self.common_step_counter = 0
        self.num_steps_per_env = 24
        self.extras = {}
        self.reward_curriculum_configs = list(cfg.rewards.curriculum_rewards or [])
        self.reward_curriculum_scales = {c["reward_name"]: c["start_value"] for c in self.reward_curriculum_configs}
        self.zero_command_proba = 0.0
        self._ep_names = ["rew_" + n for n in _abi.REWARD_NAMES if n in A.reward_names]
        self._ep_cols = [k for k, n in enumerate(_abi.REWARD_NAMES) if n in A.reward_names]
        self._xep = [("rew_" + n, k) for k, n in enumerate(_abi.XREWARD_NAMES) if n in A.reward_names]
        self._xrew_stats = T["xrew_log"][_abi.NUM_XREW:].view(torch.float32).view(EP_SLOTS, _abi.NUM_XREW)   # rows behind the 14 int64 accumulators
        # ---- library handle
        h = ctypes.c_void_p()
        _abi.check(self._lib.go2_env_create(ctypes.byref(A.config), ctypes.byref(A.model), ctypes.byref(A.buffers), ctypes.byref(h)), self._lib)
        self._h = h
        self.init_done = True

    @property
    def _stream(self):
        """Kernels are launched on torch's current stream (the capture stream while a CUDA graph is being recorded)."""
        return torch.cuda.current_stream().cuda_stream

    # episode_length_buf is re-ASSIGNED by the runner (on_policy_runner.py:118): keep the device row the kernel owns
    @property
    def episode_length_buf(self):
        return self._episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, value):
        self._episode_length_buf.copy_(value.to(self._episode_length_buf.dtype))

    def __del__(self):
        h, lib = getattr(self, "_h", None), getattr(self, "_lib", None)
        if h and lib:
            lib.go2_env_destroy(h)
            self._h = None

    # ---- VecEnv contract -----------------------------------------------------------------------------------
    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def reset_idx(self, env_ids):
        raise NotImplementedError("resets are fused into the step kernel (legged_robot.py:132-133); use reset()")

    def reset(self):
        """Reset all robots then take one zero-action step (base_task.py:82-86)."""
        sp = self._A.step_params(self.common_step_counter)
        _abi.check(self._lib.go2_env_reset_all(self._h, ctypes.byref(sp), self._stream), self._lib)
        obs, privileged_obs, _, _, _ = self.step(torch.zeros(self.num_envs, self.num_actions, device=self.device))
        return obs, privileged_obs

    def update_reward_curriculum(self, force_update: bool = False):  # legged_robot.py:144-152
        if self.reward_curriculum_configs and (self.common_step_counter % self.num_steps_per_env == 0 or force_update):
            it = self.common_step_counter // self.num_steps_per_env
            for c in self.reward_curriculum_configs:
                self.reward_curriculum_scales[c["reward_name"]] = EnvArrays._scale(c, it)

    def get_current_scale(self, config):
        return EnvArrays._scale(config, self.common_step_counter // self.num_steps_per_env)

    def step(self, actions):
        """legged_robot.py:60-100 as one kernel: clip, action delay, 4x(PD torque + physics substep), post-physics."""
        a = actions.to(device=self.device, dtype=torch.float32)
        if not a.is_contiguous():
            a = a.contiguous()
        self.common_step_counter += 1
        self.update_reward_curriculum()
        slot = self.common_step_counter % EP_SLOTS
        sp = self._A.step_params(self.common_step_counter, ep_slot=slot, reward_curriculum=self.reward_curriculum_scales)
        self.zero_command_proba = sp.zero_command_proba
        _abi.check(self._lib.go2_env_step(self._h, a.data_ptr(), ctypes.byref(sp), self._stream), self._lib)
        self._last_actions_in = a  # keep the input alive until the kernel has consumed it
        self._fill_extras(slot)
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    # ---- rollout with device-resident step parameters (CUDA-graph capturable: nothing in a launch depends on host values) -------
    def begin_rollout(self, T):
        """Upload the parameter blocks of the next T steps (Go2StepParams[T], exactly what T step() calls would pass) and advance the
        host counters past them.  Returns False when this rollout has to step eagerly: a command-range curriculum boundary
        (legged_robot.py:433-446) rewrites env_command_ranges from the host in the middle of the rollout."""
        A = self._A
        it_end = (self.common_step_counter + T) // self.num_steps_per_env
        if any(it_end >= entry["iter"] for entry in A.command_range_curriculum):
            return False
        blocks = (_abi.Go2StepParams * T)()
        slots = []
        for t in range(T):
            self.common_step_counter += 1
            self.update_reward_curriculum()
            slot = self.common_step_counter % EP_SLOTS
            blocks[t] = A.step_params(self.common_step_counter, ep_slot=slot, reward_curriculum=self.reward_curriculum_scales)
            slots.append(slot)
        self.zero_command_proba = blocks[T - 1].zero_command_proba
        nbytes = ctypes.sizeof(_abi.Go2StepParams) * T
        if getattr(self, "_sp_dev", None) is None or self._sp_dev.numel() != nbytes:
            self._sp_dev = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self._sp_dev.copy_(torch.frombuffer(bytearray(blocks), dtype=torch.uint8), non_blocking=False)
        self._rollout_slots = slots
        return True

    def step_dev(self, actions, t):
        """Step t of the rollout opened by begin_rollout(): same kernel as step(), parameters read from the device block t."""
        a = actions
        if a.dtype != torch.float32 or not a.is_contiguous():
            a = a.to(torch.float32).contiguous()
        sp = self._sp_dev.data_ptr() + t * ctypes.sizeof(_abi.Go2StepParams)
        _abi.check(self._lib.go2_env_step_dev(self._h, a.data_ptr(), sp, self._stream), self._lib)
        self._last_actions_in = a
        if self.cfg.env.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def end_rollout(self, fetch=True):
        """extras of the rollout's steps (what step() would have returned in infos['episode'] at each of them): a step without a reset re-serves the
        previous step's statistics (the kernel copies the row forward) and nothing is served before the first reset ever, like the reference's
        extras dict, which has no 'episode' key until then (legged_robot.py:229-242, on_policy_runner.py:145-146).  One D2H read per rollout."""
        assert len(self._rollout_slots) <= EP_SLOTS, "num_steps_per_env must not exceed the episode-statistics ring (GO2_EP_SLOTS)"
        if not fetch:        # un-logged rollout: no host read, the host runs ahead of the device (the statistics stay in the device ring)
            self._fill_extras(self._rollout_slots[-1])
            return []
        valid = self._A.tensors["ep_stats"][self._rollout_slots, _abi.NUM_REW + 11].cpu()
        eps = []
        for slot, v in zip(self._rollout_slots, valid.tolist()):
            self._fill_extras(slot)
            if v > 0:
                eps.append(self.extras["episode"])
        return eps

    def step_host(self, actions_np, obs_out, priv_out, rew_out, reset_out):
        """Same step through HOST buffers (numpy, ideally pinned): H2D, kernel, D2H inside the C call (bench e2e)."""
        self.step_host_begin(actions_np, obs_out, priv_out, rew_out, reset_out)
        self.step_host_end()

    def step_host_begin(self, actions_np, obs_out, priv_out, rew_out, reset_out):
        """(arguments: host torch tensors — pinned for asynchronous copies — or numpy arrays.)  First half of step_host: uploads the actions, launches the step and starts the device -> host copies on the library's copy stream, without
        waiting.  Device work enqueued before step_host_end() (the transition bookkeeping, the next policy inference) overlaps the copies; it may
        read the env's device buffers.  The host arrays are valid after step_host_end()."""
        self.common_step_counter += 1
        self.update_reward_curriculum()
        sp = self._A.step_params(self.common_step_counter, ep_slot=self.common_step_counter % EP_SLOTS,
                                 reward_curriculum=self.reward_curriculum_scales)
        p = lambda x: x.data_ptr() if hasattr(x, "data_ptr") else x.ctypes.data      # host tensors (pinned) or numpy arrays
        _abi.check(self._lib.go2_env_step_host_begin(self._h, p(actions_np), ctypes.byref(sp), p(obs_out), p(priv_out), p(rew_out), p(reset_out),
                                                     self._stream), self._lib)

    def step_host_end(self):
        _abi.check(self._lib.go2_env_step_host_end(self._h), self._lib)

    def _fill_extras(self, slot):
        """extras["episode"] / extras["time_outs"] (legged_robot.py:229-245) as 0-d views of the row the kernel filled.
        The row is only rewritten by a step in which some env reset; otherwise the previous dict is re-served, exactly
        like the reference's stale `self.extras` (on_policy_runner.py:145-146)."""
        row = self._A.tensors["ep_stats"][slot]
        ep = {}
        for name, col in zip(self._ep_names, self._ep_cols):
            ep[name] = row[col]
        for name, col in self._xep:
            ep[name] = self._xrew_stats[slot, col]
        ep["terrain_level_all"] = row[_abi.NUM_REW]
        if not self._A.plane:
            for name, cols in self._A.terrain.name2cols.items():
                ep["terrain_level_" + name] = row[_abi.NUM_REW + 1 + TERRAIN_NAMES.index(name)]
        if self.cfg.commands.curriculum:
            ep["max_command_x"] = self.command_ranges["lin_vel_x"][1]
        self._ep_latest = (ep, row[_abi.NUM_REW + 11])
        # the validity flag of the row tells the runner whether this step produced a fresh dict
        self.extras["episode"] = ep
        self.extras["episode_valid"] = row[_abi.NUM_REW + 11]
        if self.cfg.env.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf

    def render(self, sync_frame_time=True):
        """No viewer: this package is headless by construction (base_task.py:113-140 drives the Isaac Gym viewer)."""

    def set_camera(self, position, lookat):
        """No viewer (legged_robot.py:283-288); kept so that scripts written for the reference run unchanged."""

    def set_step_mode(self, mode):
        """Thread map of the step kernel: "H14" (default), "P2", "P3", "Q4", "8p", "4" — identical results, different speed (tuning / tests)."""
        _abi.check(self._lib.go2_env_set_step_mode(self._h, str(mode).encode()), self._lib)

    def substeps(self, tau, n):
        """n bare physics substeps under given joint torques (tests)."""
        _abi.check(self._lib.go2_env_substeps(self._h, tau.contiguous().data_ptr(), int(n), self._stream), self._lib)
