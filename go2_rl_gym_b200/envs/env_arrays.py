"""Host-side construction of everything the fused step kernel reads: the packed `Go2EnvConfig`, the robot model,
and every per-env array, allocated as torch tensors on one device.

This is the restatement of the reference's init-time path
(legged_gym/envs/base/legged_robot.py: _parse_cfg :1093-1106, _create_envs :952-1052 incl. the friction /
restitution / mass callbacks :320-402, _get_env_origins :1054-1091, _init_buffers :765-859,
_prepare_reward_function :909-940, _get_noise_scale_vec go2_env.py:9-21) as vectorised numpy — no O(N) Python
loop — plus the host half of the curricula that depend only on `common_step_counter`
(update_reward_curriculum :144-168, command_range_curriculum :433-446, zero_command_curriculum :556-557).
The same object serves the CUDA library (device='cuda') and, in tests only, the CPU oracle (device='cpu').
"""
import ctypes
import itertools
import math

import numpy as np
import torch

from .. import _abi
from ..utils import robot_model
from ..utils.cfg_dict import class_to_dict  # noqa: F401
from ..utils.terrain import Terrain, TERRAIN_NAMES


_U8 = ("reset_buf", "time_out_buf", "last_is_limit_vel")
_I32 = ("episode_length_buf", "terrain_levels", "terrain_types", "terrain_ids")
EP_SLOTS = 64          # = GO2_EP_SLOTS (include/go2_b200.h): the finalize kernel copies the previous slot forward on steps without a reset


class EnvArrays:
    def __init__(self, cfg, device, num_envs=None, env_offset=0, num_envs_global=None, seed=None):
        self.cfg = cfg
        self.device = torch.device(device)
        N = int(num_envs if num_envs is not None else cfg.env.num_envs)
        NG = int(num_envs_global if num_envs_global is not None else N)
        self.num_envs, self.env_offset, self.num_envs_global = N, int(env_offset), NG
        seed = int(getattr(cfg, "seed", 1) if seed is None else seed)
        self.seed = seed
        host_rng = np.random.RandomState(seed)
        m = robot_model.load_model_json()
        self.model_json = m
        self.model = robot_model.build_model_struct(m)
        dof_names = m["dof_names"]

        # ---- _parse_cfg
        self.sim_dt = cfg.sim.dt
        self.dt = cfg.control.decimation * cfg.sim.dt
        self.max_episode_length_s = cfg.env.episode_length_s
        self.max_episode_length = int(np.ceil(self.max_episode_length_s / self.dt))
        self.push_interval = int(np.ceil(cfg.domain_rand.push_interval_s / self.dt))
        self.command_ranges = class_to_dict(cfg.commands.ranges)
        self.command_range_curriculum = sorted(list(cfg.commands.command_range_curriculum), key=lambda x: x["iter"], reverse=True)
        self.max_lin_vel = self._max_lin_vel()
        self.num_steps_per_env = 24

        # ---- rewards (scales * dt, zero scales dropped)
        scales = class_to_dict(cfg.rewards.scales)
        active = {k: v for k, v in scales.items() if v != 0}
        unknown = sorted(set(active) - set(_abi.REWARD_NAMES) - set(_abi.XREWARD_NAMES))
        if unknown:      # the reference fails the same way: getattr(self, '_reward_' + name), legged_robot.py:936
            raise AttributeError(f"no reward function for {unknown}: the kernel holds {_abi.REWARD_NAMES + _abi.XREWARD_NAMES}")
        self.reward_scales = {k: v * self.dt for k, v in active.items()}
        # init_state.turn_over: a second scale set used while the robot lies flipped (legged_robot.py:257-265,922-930)
        self.turn_over = bool(cfg.init_state.turn_over)
        to = {k: v for k, v in class_to_dict(cfg.rewards.turn_over_scales).items() if v != 0} if self.turn_over else {}
        unknown = sorted(set(to) - set(_abi.REWARD_NAMES) - set(_abi.XREWARD_NAMES))
        if unknown:
            raise AttributeError(f"no reward function for {unknown} (rewards.turn_over_scales)")
        self.reward_turn_over_scales = {k: v * self.dt for k, v in to.items()}
        # the terms outside the 14 every registered go2 task uses (legged_robot.py:1236-1441, go2_env.py:62-68): evaluated only when switched on
        self.xreward_names = [n for n in _abi.XREWARD_NAMES if n in active or n in to]
        self.reward_names = [n for n in _abi.REWARD_NAMES + _abi.XREWARD_NAMES if n in active or n in to]
        self.reward_curriculum_configs = list(cfg.rewards.curriculum_rewards or [])

        # ---- terrain
        mesh = cfg.terrain.mesh_type
        if mesh not in ("plane", "heightfield", "trimesh"):
            raise ValueError("Terrain mesh type not recognised. Allowed types are [None, plane, heightfield, trimesh]")
        self.plane = mesh == "plane"
        self.terrain = Terrain(cfg.terrain, NG, seed=seed)
        gidx = np.arange(self.env_offset, self.env_offset + N)
        if not self.plane:
            t = self.terrain
            hs = np.ascontiguousarray(t.heightsamples, dtype=np.int16)
            max_init = cfg.terrain.max_init_terrain_level if cfg.terrain.curriculum else cfg.terrain.num_rows - 1
            levels = np.fmod(gidx, max_init + 1).astype(np.int32)
            # torch's own arithmetic (legged_robot.py:1072): int64 arange / python float is a FLOAT32 division, which lands just below the
            # integer at i = k NG / 4 (e.g. env 1024 of 4096 -> type 4, not 5); a float64 floor would differ from the reference there
            types = torch.div(torch.arange(NG), (NG / cfg.terrain.num_cols), rounding_mode="floor").to(torch.long).numpy()[gidx].astype(np.int32)
            cols2id = np.asarray(t.cols2id if len(t.cols2id) else [8] * cfg.terrain.num_cols, dtype=np.int32)
            ids = cols2id[types]
            origins_grid = t.env_origins.astype(np.float32)
            env_origins = origins_grid[levels, types]
            self.custom_origins = True
        else:
            hs = np.zeros((2, 2), dtype=np.int16)
            levels = np.zeros(N, np.int32); types = np.zeros(N, np.int32); ids = np.full(N, 8, np.int32)
            origins_grid = np.zeros((1, 1, 3), np.float32)
            ncols = np.floor(np.sqrt(NG)); nrows = np.ceil(NG / ncols)
            xx, yy = np.meshgrid(np.arange(nrows), np.arange(ncols), indexing="ij")
            env_origins = np.zeros((NG, 3), np.float32)
            env_origins[:, 0] = cfg.env.env_spacing * xx.flatten()[:NG]
            env_origins[:, 1] = cfg.env.env_spacing * yy.flatten()[:NG]
            env_origins = env_origins[gidx]
            self.custom_origins = False
        self.height_samples_np = hs

        # ---- per-env physical randomisation (drawn for the GLOBAL env set, then sliced -> rank independent)
        dr = cfg.domain_rand
        if dr.randomize_friction:
            buckets = host_rng.uniform(dr.friction_range[0], dr.friction_range[1], 64)
            friction = buckets[host_rng.randint(0, 64, NG)]
        else:
            friction = np.full(NG, 1.0)
        rest = host_rng.uniform(dr.restitution_range[0], dr.restitution_range[1], NG) if dr.randomize_restitution else np.zeros(NG)
        add_mass = host_rng.uniform(dr.added_mass_range[0], dr.added_mass_range[1], NG) if dr.randomize_base_mass else None
        ratio = host_rng.uniform(dr.multiplied_link_mass_range[0], dr.multiplied_link_mass_range[1], (NG, 18)) if dr.randomize_link_mass else None
        add_com = host_rng.uniform(dr.added_base_com_range[0], dr.added_base_com_range[1], (NG, 3)) if dr.randomize_base_com else None
        sl = slice(self.env_offset, self.env_offset + N)
        inertia = robot_model.composite_inertials(m, N, None if add_mass is None else add_mass[sl],
                                                  None if ratio is None else ratio[sl], None if add_com is None else add_com[sl])

        # ---- tensors
        T = {}
        dev = self.device

        def z(name, *shape, dtype=torch.float32):
            T[name] = torch.zeros(*shape, dtype=dtype, device=dev)

        for n, d in (("root_states", 13), ("dof_pos", 12), ("dof_vel", 12), ("torques", 12), ("actions", 12),
                     ("last_actions", 12), ("last_last_actions", 12), ("last_dof_vel", 12), ("obs_buf", _abi.NUM_OBS),
                     ("privileged_obs_buf", _abi.NUM_PRIV), ("base_lin_vel", 3), ("base_ang_vel", 3),
                     ("projected_gravity", 3), ("measured_heights", _abi.NUM_HEIGHT), ("commands", _abi.NUM_CMD),
                     ("commands_xy_accumulation", 2), ("env_command_ranges", 6), ("episode_sums", _abi.NUM_REW)):
            z(n, N, d)
        z("contact_forces", N, _abi.NUM_REPORT, 3); z("feet_pos", N, 4, 3); z("feet_vel", N, 4, 3)
        for n in ("rew_buf", "commands_resampling_step", "max_move_distance"):
            z(n, N)
        for n in _U8:
            z(n, N, dtype=torch.uint8)
        z("episode_length_buf", N, dtype=torch.int32)
        z("stop_heading", N, dtype=torch.uint8); z("heading_ranges", N, 2)      # heading commands (Go2EnvConfig.ext_*)
        T["terrain_levels"] = torch.from_numpy(levels).to(dev)
        T["terrain_types"] = torch.from_numpy(types).to(dev)
        T["terrain_ids"] = torch.from_numpy(ids.astype(np.int32)).to(dev)
        T["env_origins"] = torch.from_numpy(np.ascontiguousarray(env_origins, dtype=np.float32)).to(dev)
        T["terrain_origins"] = torch.from_numpy(np.ascontiguousarray(origins_grid)).to(dev)
        T["height_samples"] = torch.from_numpy(hs).to(dev)
        T["motor_strengths"] = torch.ones(N, 12, device=dev)
        T["motor_zero_offsets"] = torch.zeros(N, 12, device=dev)
        T["p_gains_multiplier"] = torch.ones(N, 12, device=dev)
        T["d_gains_multiplier"] = torch.ones(N, 12, device=dev)
        T["friction_coeffs"] = torch.from_numpy(friction[sl].astype(np.float32)).to(dev)
        T["restitutions"] = torch.from_numpy(rest[sl].astype(np.float32)).to(dev)
        T["body_inertia"] = torch.from_numpy(inertia).to(dev)
        z("xrew_sums", N, _abi.NUM_XREW); z("xrew_state", N, 12)               # extra reward terms (Go2EnvConfig.ext_xrew_*)
        z("xrew_log", _abi.XREW_LOG_BYTES // 8, dtype=torch.int64)
        z("turn_over_timer", N)                                                  # init_state.turn_over (Go2EnvConfig.ext_turn_over_timer)
        z("ep_stats", EP_SLOTS, _abi.EP_STATS)
        z("ep_accum", _abi.EP_ACCUM_FLOATS)
        T["projected_gravity"][:, 2] = -1.0
        self.tensors = T
        self.terrain_ids_np = ids
        self._update_env_command_ranges()

        # ---- packed config
        c = _abi.Go2EnvConfig()
        c.num_envs, c.env_offset = N, self.env_offset
        c.seed_lo, c.seed_hi = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
        c.sim_dt, c.decimation, c.gravity_z = cfg.sim.dt, cfg.control.decimation, cfg.sim.gravity[2]
        if cfg.control.control_type not in ("P", "V", "T"):
            raise NameError(f"Unknown controller type: {cfg.control.control_type}")          # legged_robot.py:616-617
        c.control_type = "PVT".index(cfg.control.control_type)      # legged_robot.py:605-618
        default = np.zeros(12, np.float32)
        for j, name in enumerate(dof_names):
            default[j] = cfg.init_state.default_joint_angles[name]
            kp = kd = 0.0
            for key in cfg.control.stiffness:
                if key in name:
                    kp, kd = cfg.control.stiffness[key], cfg.control.damping[key]
            c.kp[j], c.kd[j], c.default_dof_pos[j] = kp, kd, default[j]
        self.default_dof_pos_np = default
        c.action_scale = cfg.control.action_scale
        c.clip_actions, c.clip_obs = cfg.normalization.clip_actions, cfg.normalization.clip_observations
        c.randomize_action_delay = int(dr.randomize_action_delay)
        c.randomize_motor_strength = int(dr.randomize_motor_strength)
        c.randomize_motor_zero_offset = int(dr.randomize_motor_zero_offset)
        c.randomize_pd_gains = int(dr.randomize_pd_gains)
        c.push_robots, c.add_noise = int(dr.push_robots), int(cfg.noise.add_noise)
        for dst, src in ((c.motor_strength_range, dr.motor_strength_range), (c.motor_zero_offset_range, dr.motor_zero_offset_range),
                         (c.kp_mult_range, dr.stiffness_multiplier_range), (c.kd_mult_range, dr.damping_multiplier_range)):
            dst[0], dst[1] = src[0], src[1] - src[0]   # {lower, span}; span formed in double like torch_rand_float does
        c.push_interval, c.max_push_vel_xy, c.max_push_ang_vel = self.push_interval, dr.max_push_vel_xy, dr.max_push_ang_vel
        b200 = getattr(cfg.sim, "b200", None)
        c.solver_iters = getattr(b200, "solver_iterations", 4)
        c.erp, c.limit_erp = getattr(b200, "erp", 0.2), getattr(b200, "limit_erp", 0.8)
        c.contact_offset = cfg.sim.physx.contact_offset
        c.penetration_slop = getattr(b200, "penetration_slop", 0.004)
        c.max_depen_vel = cfg.sim.physx.max_depenetration_velocity
        c.bounce_threshold = cfg.sim.physx.bounce_threshold_velocity
        c.terrain_friction, c.terrain_restitution = cfg.terrain.static_friction, cfg.terrain.restitution
        c.limit_relax, c.contact_relax = getattr(b200, "limit_relax", 0.5), getattr(b200, "contact_relax", 0.7)
        c.state_guard = int(getattr(b200, "state_guard", 1))
        c.max_base_lin_vel, c.max_base_ang_vel = cfg.asset.max_linear_velocity, cfg.asset.max_angular_velocity
        c.mesh_type = 0 if self.plane else 1
        c.hf_rows, c.hf_cols = hs.shape
        c.hscale, c.vscale, c.border = cfg.terrain.horizontal_scale, cfg.terrain.vertical_scale, cfg.terrain.border_size
        c.num_levels, c.num_types = origins_grid.shape[0], origins_grid.shape[1]
        c.terrain_length = cfg.terrain.terrain_length
        c.terrain_curriculum = int(cfg.terrain.curriculum)
        c.move_down_by_accumulated_xy_command = int(cfg.terrain.move_down_by_accumulated_xy_command)
        c.custom_origins = int(self.custom_origins)
        cm = cfg.commands
        if cm.curriculum:      # vestigial in this fork: it rewrites command_ranges["lin_vel_x"], which the sampler no longer reads
            raise NotImplementedError("commands.curriculum is not built (SURVEY 8f-3, DESIGN.md section 6)")
        c.turn_over = int(self.turn_over)
        if self.turn_over:     # legged_robot.py:642-691,585-590; go2_config.py:23-30,125-128
            ini = cfg.init_state
            pr = [float(x) for x in ini.turn_over_proportions]
            c.turn_over_proportions[0], c.turn_over_proportions[1], c.turn_over_proportions[2] = pr[0], pr[0] + pr[1], pr[0] + pr[1] + pr[2]   # cumulative, formed in double
            for dst, key in ((c.turn_over_back_height, "backflip"), (c.turn_over_side_height, "sideflip")):
                lo, hi = ini.turn_over_init_heights[key]
                dst[0], dst[1] = lo, hi - lo                                       # {lower, span in double} like torch_rand_float
            c.turn_over_zero_time_back, c.turn_over_zero_time_side = cm.turn_over_zero_time["backflip"], cm.turn_over_zero_time["sideflip"]
            c.turn_over_roll_threshold = cfg.rewards.turn_over_roll_threshold
            for k, name in enumerate(_abi.REWARD_NAMES):
                c.to_scales[k] = self.reward_turn_over_scales.get(name, 0.0)
            for k, name in enumerate(_abi.XREWARD_NAMES):
                c.to_xscales[k] = self.reward_turn_over_scales.get(name, 0.0)
        c.ext_turn_over_timer_lo, c.ext_turn_over_timer_hi = T["turn_over_timer"].data_ptr() & 0xFFFFFFFF, T["turn_over_timer"].data_ptr() >> 32
        # heading commands (legged_robot.py:411-419)
        c.heading_command, c.stop_heading_at_limit = int(bool(cm.heading_command)), int(bool(getattr(cm, "stop_heading_at_limit", False)))
        for name, key in (("ext_stop_heading", "stop_heading"), ("ext_heading_ranges", "heading_ranges")):       # addresses as 32-bit halves (go2_b200.h)
            setattr(c, name + "_lo", T[key].data_ptr() & 0xFFFFFFFF); setattr(c, name + "_hi", T[key].data_ptr() >> 32)
        c.resampling_time, c.dynamic_resample_commands = cm.resampling_time, int(cm.dynamic_resample_commands)
        c.limit_vel_prob = cm.limit_vel_prob
        c.limit_vel_invert_when_continuous = int(cm.limit_vel_invert_when_continuous)
        c.limit_ang_vel_at_zero_command_prob = cm.limit_ang_vel_at_zero_command_prob
        comb = list(itertools.product(cm.limit_vel["lin_vel_x"], cm.limit_vel["lin_vel_y"], cm.limit_vel["ang_vel_yaw"]))
        assert comb == list(itertools.product([-1, 1], [-1, 1], [-1, 0, 1])), "limit_vel table is baked into the kernel"
        c.max_episode_length, c.max_episode_length_s, c.dt = self.max_episode_length, self.max_episode_length_s, self.dt
        for k, name in enumerate(_abi.REWARD_NAMES):
            c.reward_scales[k] = self.reward_scales.get(name, 0.0)
        c.num_xrew = len(self.xreward_names)
        for k, name in enumerate(_abi.XREWARD_NAMES):
            c.xrew_scales[k] = self.reward_scales.get(name, 0.0)
        rw = cfg.rewards
        c.soft_dof_vel_limit, c.soft_torque_limit = getattr(rw, "soft_dof_vel_limit", 1.0), getattr(rw, "soft_torque_limit", 1.0)
        c.max_contact_force, c.min_legs_distance = getattr(rw, "max_contact_force", 100.0), getattr(rw, "min_legs_distance", 0.1)
        for name, key in (("ext_xrew_sums", "xrew_sums"), ("ext_xrew_state", "xrew_state"), ("ext_xrew_log", "xrew_log")):
            setattr(c, name + "_lo", T[key].data_ptr() & 0xFFFFFFFF); setattr(c, name + "_hi", T[key].data_ptr() >> 32)
        c.only_positive_rewards = int(bool(cfg.rewards.only_positive_rewards))     # off for every go2 task (go2_config.py:159)
        c.tracking_sigma, c.base_height_target = cfg.rewards.tracking_sigma, cfg.rewards.base_height_target
        for j in range(12):
            lo, hi = self.model.q_lower[j], self.model.q_upper[j]
            mid, rng = (lo + hi) / 2, hi - lo
            c.soft_dof_limit_lo[j] = mid - 0.5 * rng * cfg.rewards.soft_dof_pos_limit
            c.soft_dof_limit_hi[j] = mid + 0.5 * rng * cfg.rewards.soft_dof_pos_limit
        ds = cfg.rewards.dynamic_sigma
        c.dynamic_sigma = int(ds is not None)
        if ds is not None:
            c.ds_min_lin, c.ds_max_lin, c.ds_min_ang, c.ds_max_ang = ds["min_lin_vel"], ds["max_lin_vel"], ds["min_ang_vel"], ds["max_ang_vel"]
            for k in range(9):
                c.ds_max_sigma[k] = ds["max_sigma"][k]
        os_ = cfg.normalization.obs_scales
        c.obs_scale_lin_vel, c.obs_scale_ang_vel, c.obs_scale_dof_pos = os_.lin_vel, os_.ang_vel, os_.dof_pos
        c.obs_scale_dof_vel, c.obs_scale_height = os_.dof_vel, os_.height_measurements
        ns, nl = cfg.noise.noise_scales, cfg.noise.noise_level
        nv = np.zeros(_abi.NUM_OBS, np.float32)      # go2_env.py:9-21
        nv[0:3] = ns.ang_vel * nl * os_.ang_vel
        nv[3:6] = ns.gravity * nl
        nv[9:21] = ns.dof_pos * nl * os_.dof_pos
        nv[21:33] = ns.dof_vel * nl * os_.dof_vel
        self.noise_scale_vec_np = nv
        for i in range(_abi.NUM_OBS):
            c.noise_scale_vec[i] = nv[i]
        px, py = cfg.terrain.measured_points_x, cfg.terrain.measured_points_y
        assert len(px) * len(py) == _abi.NUM_HEIGHT, "the fused kernel is specialised for the 17x11 height scan"
        cnt = 0
        i = 0
        for x in px:                                   # torch.meshgrid(x, y) 'ij' flatten, legged_robot.py:1180-1185
            for y in py:
                c.height_points[i][0], c.height_points[i][1] = x, y
                inside = (np.float32(x) >= np.float32(-0.2)) and (np.float32(x) <= np.float32(0.2)) and \
                         (np.float32(y) >= np.float32(-0.15)) and (np.float32(y) <= np.float32(0.15))
                c.base_height_mask[i] = 1.0 if inside else 0.0
                cnt += inside
                i += 1
        c.num_base_height_points = float(cnt)
        base_init = cfg.init_state.pos + cfg.init_state.rot + cfg.init_state.lin_vel + cfg.init_state.ang_vel
        for k in range(13):
            c.base_init_state[k] = base_init[k]
        self.config = c

        b = _abi.Go2EnvBuffers()
        for name in _abi.PTR_FIELDS:
            t = T[name]
            assert t.is_contiguous()
            setattr(b, name, t.data_ptr())
        self.buffers = b

    # ---- host half of the curricula ------------------------------------------------------------------
    def _max_lin_vel(self):
        r = self.command_ranges
        return max(abs(r["lin_vel_x"][0]), abs(r["lin_vel_x"][1]), abs(r["lin_vel_y"][0]), abs(r["lin_vel_y"][1]))

    def _update_env_command_ranges(self):  # legged_robot.py:861-907
        r = self.command_ranges
        table = np.zeros((9, 8), np.float32)
        table[:, 6:8] = r["heading"]                   # per-terrain heading limits apply only with heading commands (:899-907)
        for tid, tr in enumerate(self.cfg.commands.terrain_max_command_ranges):
            for a, key in enumerate(("lin_vel_x", "lin_vel_y", "ang_vel_yaw") + (("heading",) if self.cfg.commands.heading_command else ())):
                table[tid, 2 * a] = max(tr[key][0], r[key][0])
                table[tid, 2 * a + 1] = min(tr[key][1], r[key][1])
        if self.plane:  # no terrain_ids attribute in the reference -> global ranges
            rows = np.tile(np.array([r["lin_vel_x"][0], r["lin_vel_x"][1], r["lin_vel_y"][0], r["lin_vel_y"][1],
                                     r["ang_vel_yaw"][0], r["ang_vel_yaw"][1], r["heading"][0], r["heading"][1]], np.float32), (self.num_envs, 1))
        else:
            rows = table[self.terrain_ids_np]
        self.tensors["env_command_ranges"].copy_(torch.from_numpy(np.ascontiguousarray(rows[:, :6])))
        self.tensors["heading_ranges"].copy_(torch.from_numpy(np.ascontiguousarray(rows[:, 6:8])))

    @staticmethod
    def _scale(config, it):  # get_current_scale, legged_robot.py:154-168
        p = (it - config["start_iter"]) / (config["end_iter"] - config["start_iter"])
        p = max(min(p, 1.0), 0.0)
        return (1.0 - p) * config["start_value"] + p * config["end_value"]

    def reward_curriculum_scales(self, common_step_counter):
        """Scales as the reference holds them at this step: refreshed only when counter % 24 == 0 (:147)."""
        it = (common_step_counter - common_step_counter % self.num_steps_per_env) // self.num_steps_per_env
        return {cfgd["reward_name"]: self._scale(cfgd, it) for cfgd in self.reward_curriculum_configs}

    def step_params(self, common_step_counter, ep_slot=0, reward_curriculum=None):
        """common_step_counter: value after this step's increment."""
        it = common_step_counter // self.num_steps_per_env
        changed = False
        for i in range(len(self.command_range_curriculum) - 1, -1, -1):  # legged_robot.py:433-446
            entry = self.command_range_curriculum[i]
            if it >= entry["iter"]:
                for key in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw", "heading"):
                    self.command_ranges[key] = entry[key]
                self.command_range_curriculum.pop(i)
                changed = True
        if changed:
            self.max_lin_vel = self._max_lin_vel()
            self._update_env_command_ranges()
        rc = reward_curriculum if reward_curriculum is not None else self.reward_curriculum_scales(common_step_counter)
        # everything but the two counters depends on the iteration and the curriculum scales only: the block of the previous call is reused while they
        # stand (the per-step host loops call this 24 times per iteration)
        key = (it, tuple(rc.values()), self.max_lin_vel)
        memo = getattr(self, "_sp_memo", None)
        if memo is not None and memo[0] == key:
            sp = _abi.Go2StepParams.from_buffer_copy(memo[1])
        else:
            sp = _abi.Go2StepParams()
            for k, name in enumerate(_abi.REWARD_NAMES):
                sp.reward_curriculum[k] = rc.get(name, 1.0)
            for k, name in enumerate(_abi.XREWARD_NAMES):
                sp.xrew_curriculum[k] = rc.get(name, 1.0)
            zc = self.cfg.commands.zero_command_curriculum
            sp.zero_command_proba = self._scale(zc, it) if zc is not None else 0.0
            sp.max_lin_vel = self.max_lin_vel
            self._sp_memo = (key, bytes(sp))
        sp.common_step_counter = int(common_step_counter) & 0xFFFFFFFF
        sp.ep_slot = int(ep_slot)
        return sp
