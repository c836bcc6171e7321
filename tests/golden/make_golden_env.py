#!/usr/bin/env python3
"""Generate tests/golden/env_*.npz by running the UNMODIFIED reference env Python over the oracle's physics.

Runs only in the build container (needs /root/reference).  It imports legged_gym from /root/reference through the
isaacgym stand-in in tests/ref_stub, builds a `Go2Robot` object without Isaac Gym (its `self.gym` is a fake whose
`simulate()` advances the state with the oracle's physics substep — the "inner boundary" of SURVEY 8b), replaces
the torch random draws the reference makes (legged_robot.py:72,197-206,628,645,698,703,718,719,1166,469-575;
isaacgym_utils.py:40; go2_env.py:53) by the build's Philox draws keyed (env, step, stream), and records K full
`LeggedRobot.step()` calls.  tests/test_oracle_golden.py replays the same initial state through the oracle's own
post-physics restatement and compares: this pins PD torques, action delay, prelude, command resampling, height scan,
termination, all 14 rewards, reset / terrain curriculum, pushes, both observation vectors and their ordering quirks
against the reference's own code.  (gym.simulate itself — PhysX — stays unpinned.)

Usage: python tests/golden/make_golden_env.py   (writes tests/golden/env_rough.npz, env_plane.npz, env_cmdcur.npz)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
# the reference tree first: the repo root carries import shims that are also called legged_gym / rsl_rl
sys.path[:0] = [os.path.join(ROOT, "tests", "ref_stub"), "/root/reference", "/root/reference/rsl_rl", ROOT]

from go2_rl_gym_b200 import _abi  # noqa: E402
from go2_rl_gym_b200.envs.env_arrays import EnvArrays  # noqa: E402
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg as MyGO2Cfg  # noqa: E402
from oracle.oracle import OracleEnv  # noqa: E402

ST_DELAY, ST_NOISE, ST_PUSH, ST_RESET_DR, ST_RESET_STATE, ST_CMD_CB, ST_CMD_RESET = range(7)


def philox(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all args uint32 arrays (broadcastable)."""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    M = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & M
        n1 = p1 & M
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & M
        n3 = p0 & M
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & M
        k1 = (k1 + np.uint64(0xBB67AE85)) & M
    return np.stack([c0, c1, c2, c3], -1).astype(np.uint32)


def u01(x):
    return torch.from_numpy(((x >> 8).astype(np.float32) * np.float32(1.0 / 16777216.0)))


class Draws:
    """Philox draws in the layout the oracle / kernel use (oracle/go2_oracle.cpp, enum Stream)."""

    def __init__(self, seed, env_offset=0):
        self.k0, self.k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
        self.env_offset = env_offset
        self.step = 0

    def raw(self, ids, stream, block):
        ids = np.asarray(ids, dtype=np.uint32) + np.uint32(self.env_offset)
        return philox(ids, np.uint32(self.step), np.uint32(stream), np.uint32(block), self.k0, self.k1)  # [n,4]

    def u(self, ids, stream, block, comp):
        return u01(self.raw(ids, stream, block)[:, comp])

    def i(self, ids, stream, block, comp, mod):
        return torch.from_numpy((self.raw(ids, stream, block)[:, comp] % np.uint32(mod)).astype(np.int64))


def _ids(t):
    return t.cpu().numpy().astype(np.int64)


def install_rng(draws, num_envs):
    """Replace the random functions the reference env calls, dispatching on the call site."""
    import legged_gym.envs.base.legged_robot as LR
    import legged_gym.envs.go2.go2_env as GE
    import legged_gym.utils.isaacgym_utils as IU

    def in_reset():
        f = sys._getframe(2)
        while f is not None:
            if f.f_code.co_name == "reset_idx":
                return True
            f = f.f_back
        return False

    def torch_rand_float(lower, upper, shape, device):
        fr = sys._getframe(1)
        line = fr.f_lineno
        all_ids = np.arange(num_envs)
        if line in (197, 198, 199, 202, 205, 206):
            comp = {197: 0, 198: 0, 199: 0, 202: 1, 205: 2, 206: 3}[line]
            ids = _ids(fr.f_locals["env_ids"])
            u = torch.stack([draws.u(ids, ST_RESET_DR, j, comp) for j in range(12)], 1)
        elif line == 628:
            ids = _ids(fr.f_locals["env_ids"])
            u = torch.stack([draws.u(ids, ST_RESET_STATE, j // 4, j % 4) for j in range(12)], 1)
        elif line == 645:
            ids = _ids(fr.f_locals["env_ids"])
            u = draws.u(ids, ST_RESET_STATE, 3, 0).unsqueeze(1)
        elif line in (661, 672):    # init_state.turn_over: initial heights of the robots reset on their back / side (draws of the flipped envs only)
            ids = _ids(fr.f_locals["env_ids"][fr.f_locals["back_mask" if line == 661 else "side_mask"]])
            u = draws.u(ids, ST_RESET_STATE, 6, 0 if line == 661 else 1).unsqueeze(1)
        elif line == 698:
            ids = _ids(fr.f_locals["env_ids"])
            u = torch.stack([draws.u(ids, ST_RESET_STATE, 3, 1), draws.u(ids, ST_RESET_STATE, 3, 2)], 1)
        elif line == 703:
            ids = _ids(fr.f_locals["env_ids"])
            u = torch.stack([draws.u(ids, ST_RESET_STATE, 4, k) for k in range(4)] + [draws.u(ids, ST_RESET_STATE, 5, k) for k in range(2)], 1)
        elif line == 718:
            u = torch.stack([draws.u(all_ids, ST_PUSH, 0, 0), draws.u(all_ids, ST_PUSH, 0, 1)], 1)
        elif line == 719:
            u = torch.stack([draws.u(all_ids, ST_PUSH, 0, 2), draws.u(all_ids, ST_PUSH, 0, 3), draws.u(all_ids, ST_PUSH, 1, 0)], 1)
        else:
            raise RuntimeError(f"unmapped torch_rand_float call site legged_robot.py:{line}")
        assert tuple(u.shape) == tuple(int(s) for s in shape), (line, u.shape, shape)
        return (upper - lower) * u + lower

    class TorchProxy(types.ModuleType):
        def __init__(self, where):
            super().__init__("torch")
            self._where = where

        def __getattr__(self, name):
            return getattr(torch, name)

        def rand(self, *shape, device=None):
            fr = sys._getframe(1)
            line, fn = fr.f_lineno, fr.f_code.co_name
            if self._where == "IU":
                ids = _ids(fr.f_locals["env_ids"])
                if fn == "sample_single_interval":      # commands.dynamic_resample_commands = False (legged_robot.py:479-504)
                    assert line == 53, line
                    comp = {479: 0, 485: 1, 492: 2, 499: 2}[fr.f_back.f_lineno]
                else:
                    assert fn == "sample_disjoint_intervals" and line == 40, (fn, line)
                    comp = {454: 0, 461: 1}[fr.f_back.f_lineno]
                return draws.u(ids, ST_CMD_RESET if in_reset() else ST_CMD_CB, 0, comp)
            stream = ST_CMD_RESET if in_reset() else ST_CMD_CB
            if line in (469, 474):      # the yaw draw; with heading commands the same draw feeds the heading target (:469)
                return draws.u(_ids(fr.f_locals["env_ids"]), stream, 0, 2)
            if line == 509:
                return draws.u(_ids(fr.f_locals["env_ids"]), stream, 0, 3)
            if line == 654:     # turn_over: which flip (legged_robot.py:654)
                return draws.u(_ids(fr.f_locals["env_ids"]), ST_RESET_STATE, 5, 2)
            if line == 674:     # turn_over: which side (:674)
                return draws.u(_ids(fr.f_locals["env_ids"][fr.f_locals["side_ids"]]), ST_RESET_STATE, 5, 3)
            if line == 571:
                return draws.u(_ids(fr.f_locals["zero_env_ids"]), stream, 1, 1)
            if line == 575:
                return draws.u(_ids(fr.f_locals["add_ang_env_ids"]), stream, 1, 2)
            raise RuntimeError(f"unmapped torch.rand call site {fn}:{line}")

        def randint(self, low, high, size, device=None):
            fr = sys._getframe(1)
            line = fr.f_lineno
            if line == 72:
                return draws.i(np.arange(num_envs), ST_DELAY, 0, 0, high).view(num_envs, 1)
            if line == 525:
                stream = ST_CMD_RESET if in_reset() else ST_CMD_CB
                return draws.i(_ids(fr.f_locals["change_lim_env_ids"]), stream, 1, 0, high)
            raise RuntimeError(f"unmapped torch.randint call site :{line}")

        def randint_like(self, t, high):
            fr = sys._getframe(1)
            assert fr.f_lineno in (1165, 1166, 1167), fr.f_lineno
            return draws.i(_ids(fr.f_locals["env_ids"]), ST_RESET_STATE, 3, 3, high)

        def rand_like(self, t):
            fr = sys._getframe(1)
            assert fr.f_code.co_name == "compute_observations"
            n = t.shape[1]
            return torch.stack([draws.u(np.arange(num_envs), ST_NOISE, i // 4, i % 4) for i in range(n)], 1)

    LR.torch = TorchProxy("LR")
    GE.torch = TorchProxy("GE")
    IU.torch = TorchProxy("IU")
    LR.torch_rand_float = torch_rand_float


class FakeGym:
    """Owns the simulated state (an EnvArrays + OracleEnv); the reference's tensors are separate and refreshed from it."""

    def __init__(self, arrays, oracle):
        self.A, self.O = arrays, oracle
        self.ref = None
        self.tau = None

    def __getattr__(self, name):  # every other gym call is a no-op
        return lambda *a, **k: None

    def set_dof_actuation_force_tensor(self, sim, t):
        self.tau = t

    def simulate(self, sim):
        self.O.substeps(self.tau.view(self.A.num_envs, 12), 1)

    def refresh_dof_state_tensor(self, sim):
        r, T = self.ref, self.A.tensors
        if not hasattr(r, 'dof_state'):
            return
        ds = r.dof_state.view(self.A.num_envs, 12, 2)
        ds[..., 0] = T["dof_pos"]
        ds[..., 1] = T["dof_vel"]

    def refresh_actor_root_state_tensor(self, sim):
        if not hasattr(self.ref, 'root_states'):
            return
        self.ref.root_states.copy_(self.A.tensors["root_states"])

    def refresh_net_contact_force_tensor(self, sim):
        if not hasattr(self.ref, 'contact_forces'):
            return
        self.ref.contact_forces.copy_(self.A.tensors["contact_forces"])

    def refresh_rigid_body_state_tensor(self, sim):
        if not hasattr(self.ref, 'rigid_body_states'):
            return
        self.O.feet()
        rb = self.ref.rigid_body_states.view(self.A.num_envs, 19, 13)
        rb[:, [6, 10, 14, 18], 0:3] = self.A.tensors["feet_pos"]
        rb[:, [6, 10, 14, 18], 7:10] = self.A.tensors["feet_vel"]

    def set_dof_state_tensor_indexed(self, sim, t, ids, n):
        ids = ids.long()
        ds = t.view(self.A.num_envs, 12, 2)
        self.A.tensors["dof_pos"][ids] = ds[ids, :, 0]
        self.A.tensors["dof_vel"][ids] = ds[ids, :, 1]

    def set_actor_root_state_tensor_indexed(self, sim, t, ids, n):
        ids = ids.long()
        self.A.tensors["root_states"][ids] = t[ids]


def build_reference_env(A, oracle, ref_cfg):
    """A `Go2Robot` (reference class) wired to the fake gym, mirroring what __init__/_create_envs would have set."""
    from legged_gym.envs.go2.go2_env import Go2Robot
    from legged_gym.envs.base.legged_robot import LeggedRobot
    N = A.num_envs
    T = A.tensors
    r = object.__new__(Go2Robot)
    r.cfg = ref_cfg
    r.sim_params = types.SimpleNamespace(dt=ref_cfg.sim.dt)
    r.height_samples = None
    r.debug_viz = False
    r.init_done = False
    LeggedRobot._parse_cfg(r, ref_cfg)
    r.gym = FakeGym(A, oracle)
    r.gym.ref = r
    r.sim = None
    r.device = "cpu"
    r.headless = True
    r.viewer = None
    r.num_envs, r.num_obs, r.num_privileged_obs, r.num_actions = N, 45, 263, 12
    r.obs_buf = torch.zeros(N, 45)
    r.rew_buf = torch.zeros(N)
    r.reset_buf = torch.ones(N, dtype=torch.long)
    r.episode_length_buf = torch.zeros(N, dtype=torch.long)
    r.time_out_buf = torch.zeros(N, dtype=torch.bool)
    r.privileged_obs_buf = torch.zeros(N, 263)
    r.extras = {}
    r.up_axis_idx = 2
    # --- what _create_envs leaves behind
    m = A.model_json
    r.num_dof = r.num_dofs = 12
    r.num_bodies = 19
    r.dof_names = m["dof_names"]
    body_names = m["report_bodies"]
    idx = lambda key: torch.tensor([i for i, n in enumerate(body_names) if key in n], dtype=torch.long)
    r.feet_indices = idx("foot")
    r.penalised_contact_indices = torch.cat([idx("thigh"), idx("calf")])
    r.termination_contact_indices = idx("base")
    r.motor_zero_offsets = torch.zeros(N, 12)
    r.p_gains_multiplier = torch.ones(N, 12)
    r.d_gains_multiplier = torch.ones(N, 12)
    if ref_cfg.rewards.dynamic_sigma:        # legged_robot.py:1009-1011
        r.dynamic_sigma_cfg = ref_cfg.rewards.dynamic_sigma
        r.terrain_max_sigmas = torch.tensor(r.dynamic_sigma_cfg["max_sigma"])
    if not A.plane:
        r.terrain = A.terrain
        r.height_samples = torch.tensor(A.terrain.heightsamples).view(A.terrain.tot_rows, A.terrain.tot_cols)
    LeggedRobot._get_env_origins(r)
    props = np.zeros(12, dtype=[("lower", np.float32), ("upper", np.float32), ("velocity", np.float32), ("effort", np.float32)])
    for i, j in enumerate(m["joints"]):  # the structured array gym.get_asset_dof_properties returns
        props[i] = (j["lower"], j["upper"], j["velocity"], j["effort"])
    LeggedRobot._process_dof_props(r, props, 0)
    base_init = ref_cfg.init_state.pos + ref_cfg.init_state.rot + ref_cfg.init_state.lin_vel + ref_cfg.init_state.ang_vel
    r.base_init_state = torch.tensor(base_init, dtype=torch.float)
    # --- tensors the fake gym "acquires"
    root = torch.zeros(N, 13)
    dof_state = torch.zeros(N * 12, 2)
    contact = torch.zeros(N * 19, 3)
    rigid = torch.zeros(N * 19, 13)
    acquire = {"acquire_actor_root_state_tensor": root, "acquire_dof_state_tensor": dof_state,
               "acquire_net_contact_force_tensor": contact, "acquire_rigid_body_state_tensor": rigid}
    for k, v in acquire.items():
        setattr(r.gym, k, (lambda vv: (lambda sim: vv))(v))
    root[:, 6] = 1.0
    Go2Robot._init_buffers(r)
    LeggedRobot._prepare_reward_function(r)
    r.init_done = True
    r.reward_curriculum_scales = {}
    r.reward_curriculum_configs = ref_cfg.rewards.curriculum_rewards
    for c in r.reward_curriculum_configs:
        r.reward_curriculum_scales[c["reward_name"]] = c["start_value"]
    r.num_steps_per_env = 24
    r.render = lambda: None
    return r


STATE_KEYS = ["root_states", "dof_pos", "dof_vel", "actions", "last_actions", "last_last_actions", "last_dof_vel", "torques",
              "commands", "commands_resampling_step", "commands_xy_accumulation", "last_is_limit_vel", "episode_length_buf",
              "terrain_levels", "env_origins", "max_move_distance", "motor_strengths", "motor_zero_offsets",
              "p_gains_multiplier", "d_gains_multiplier", "episode_sums", "friction_coeffs", "restitutions", "body_inertia",
              "contact_forces", "xrew_sums", "xrew_state", "turn_over_timer"]
OUT_KEYS = ["obs_buf", "privileged_obs_buf", "rew_buf", "reset_buf", "time_out_buf", "root_states", "dof_pos", "dof_vel", "torques",
            "commands", "commands_resampling_step", "commands_xy_accumulation", "last_is_limit_vel", "episode_length_buf",
            "terrain_levels", "env_origins", "max_move_distance", "motor_strengths", "motor_zero_offsets", "p_gains_multiplier",
            "d_gains_multiplier", "episode_sums", "base_lin_vel", "base_ang_vel", "projected_gravity", "measured_heights",
            "last_actions", "last_last_actions", "last_dof_vel", "contact_forces", "feet_pos", "feet_vel"]


def load_state_into_reference(r, A):
    """Copy the oracle-side state S0 into the reference env's own buffers."""
    T = A.tensors
    N = A.num_envs
    r.root_states.copy_(T["root_states"])
    ds = r.dof_state.view(N, 12, 2)
    ds[..., 0] = T["dof_pos"]; ds[..., 1] = T["dof_vel"]
    for k in ("actions", "last_actions", "last_dof_vel", "torques", "commands", "commands_resampling_step",
              "commands_xy_accumulation", "max_move_distance", "motor_strengths", "motor_zero_offsets",
              "p_gains_multiplier", "d_gains_multiplier"):
        getattr(r, k).copy_(T[k])
    r.last_last_actions = T["last_last_actions"].clone()
    r.last_is_limit_vel.copy_(T["last_is_limit_vel"].bool())
    r.stop_heading.copy_(T["stop_heading"].bool())
    r.episode_length_buf.copy_(T["episode_length_buf"].long())
    if not A.plane:
        r.terrain_levels.copy_(T["terrain_levels"].long())
        r.env_origins.copy_(T["env_origins"])
    for k, name in enumerate(_abi.REWARD_NAMES):
        r.episode_sums[name].copy_(T["episode_sums"][:, k])
    r.contact_forces.copy_(T["contact_forces"])
    for k, name in enumerate(_abi.XREWARD_NAMES):      # the reward terms outside the GO2 defaults, when the case switches them on
        if name in r.episode_sums:
            r.episode_sums[name].copy_(T["xrew_sums"][:, k])
    r.turn_over_timer.copy_(T["turn_over_timer"])
    r.feet_air_time.copy_(T["xrew_state"][:, 0:4])
    r.last_contacts.copy_(T["xrew_state"][:, 4:8].bool())
    r.last_contacts2 = T["xrew_state"][:, 8:12].bool().clone()        # created lazily by _reward_base_height (legged_robot.py:1248-1249)


def _apply(cfg, overrides):
    """overrides: {"domain_rand.push_robots": False, ...} set on a config tree by dotted path."""
    for path, val in (overrides or {}).items():
        node = cfg
        parts = path.split(".")
        for p in parts[:-1]:
            node = getattr(node, p)
        assert hasattr(node, parts[-1]) or parts[:-1] in (["rewards", "scales"], ["rewards", "turn_over_scales"]), path      # reward scales may be added (the reference looks the function up by name)
        setattr(node, parts[-1], val)


def make_case(name, plane, N=48, K=6, seed=7, start_counter=24 * 700, control_type="P", only_positive=False, heading=False, overrides=None, b200=None):
    torch.manual_seed(seed)
    cfg = MyGO2Cfg()
    cfg.env.num_envs = N
    cfg.terrain.mesh_type = "plane" if plane else "heightfield"
    cfg.seed = seed
    cfg.control.control_type, cfg.rewards.only_positive_rewards = control_type, only_positive     # switches outside the GO2 defaults
    cfg.commands.heading_command = heading
    _apply(cfg, overrides)
    for k, v in (b200 or {}).items():       # knobs of this package's contact solver (no reference counterpart: the oracle physics serves both sides)
        setattr(cfg.sim.b200, k, v)
    if heading:
        # the reference clips the masked yaw command with UNMASKED [N] bounds (legged_robot.py:415-419): it raises a shape error as soon as one
        # env holds its heading (stop_heading), so its heading mode only runs while no env ever stops: no stop at limits, no yaw kick at zero commands
        cfg.commands.stop_heading_at_limit, cfg.commands.limit_ang_vel_at_zero_command_prob = False, 0.0
    A = EnvArrays(cfg, "cpu", seed=seed)
    O = OracleEnv(A)
    O.common_step_counter = start_counter
    O.reset_all()
    T = A.tensors
    g = torch.Generator().manual_seed(seed)
    # warm-up with the oracle's full step so that states are mid-episode and contacts are established
    for _ in range(30):
        O.step(0.5 * torch.randn(N, 12, generator=g))
    # provoke every branch within the recorded window
    ep = torch.randint(0, 1200, (N,), generator=g).int()
    ep[0:4] = torch.tensor([1249, 1250, 1251, 1248]).int()          # time-outs (ep_len > 1250) and the max-1 guard
    ep[4:8] = torch.tensor([198, 199, 399, 200]).int()              # pushes at ep_len % 200 == 0
    T["episode_length_buf"].copy_(ep)
    T["commands_resampling_step"][8:20] = torch.tensor([1., 2., 3., 1., 2., 3., 4., 5., 1., 1., 2., 2.])
    T["last_is_limit_vel"][8:14] = 1
    T["max_move_distance"][0:4] = torch.tensor([5.0, 0.1, 4.5, 0.0])
    T["commands_xy_accumulation"][0:4] = torch.tensor([[1.0, 1.0], [3.0, 0.5], [0.2, 0.1], [2.0, 2.0]])
    if not plane:
        T["terrain_levels"][0] = 9                                   # overflow -> random level
        T["terrain_levels"][1] = 0                                   # clamp at 0
    # two robots lying on their side / upside down -> base contact termination
    T["root_states"][20, 3:7] = torch.tensor([0.7071, 0.0, 0.0, 0.7071]); T["root_states"][20, 2] = T["env_origins"][20, 2] + 0.12
    T["root_states"][21, 3:7] = torch.tensor([1.0, 0.0, 0.0, 0.0]); T["root_states"][21, 2] = T["env_origins"][21, 2] + 0.15
    if heading:      # mid-episode state of the heading mode: targets drawn
        T["commands"][:, 3] = (torch.rand(N, generator=g) * 2 - 1) * 3.14
    S0 = {k: T[k].clone() for k in STATE_KEYS + ["stop_heading"]}
    actions = 0.8 * torch.randn(K, N, 12, generator=g)
    actions[2, 5] = 150.0                                            # exercises clip_actions

    # ---- reference run
    import legged_gym.envs  # noqa: F401  (registers tasks; proves the import path)
    from legged_gym.envs.go2.go2_config import GO2Cfg as RefGO2Cfg
    ref_cfg = RefGO2Cfg()
    ref_cfg.env.num_envs = N
    ref_cfg.terrain.mesh_type = cfg.terrain.mesh_type
    ref_cfg.control.control_type, ref_cfg.rewards.only_positive_rewards = control_type, only_positive
    ref_cfg.commands.heading_command = heading
    _apply(ref_cfg, overrides)
    if heading:
        ref_cfg.commands.stop_heading_at_limit, ref_cfg.commands.limit_ang_vel_at_zero_command_prob = False, 0.0
    draws = Draws(seed)
    install_rng(draws, N)
    r = build_reference_env(A, O, ref_cfg)
    load_state_into_reference(r, A)
    r.common_step_counter = start_counter + 30
    r.update_reward_curriculum(force_update=True)
    r.zero_command_proba = A.step_params(r.common_step_counter).zero_command_proba
    outs = []
    for k in range(K):
        draws.step = r.common_step_counter + 1
        obs, priv, rew, reset, extras = r.step(actions[k].clone())
        rec = {"obs_buf": obs, "privileged_obs_buf": priv, "rew_buf": rew, "reset_buf": reset.to(torch.uint8),
               "time_out_buf": r.time_out_buf.to(torch.uint8), "root_states": T["root_states"], "dof_pos": T["dof_pos"],
               "dof_vel": T["dof_vel"], "torques": r.torques, "commands": r.commands,
               "commands_resampling_step": r.commands_resampling_step, "commands_xy_accumulation": r.commands_xy_accumulation,
               "last_is_limit_vel": r.last_is_limit_vel.to(torch.uint8), "stop_heading": r.stop_heading.to(torch.uint8),
               "episode_length_buf": r.episode_length_buf.int(),
               "max_move_distance": r.max_move_distance, "motor_strengths": r.motor_strengths,
               "motor_zero_offsets": r.motor_zero_offsets, "p_gains_multiplier": r.p_gains_multiplier,
               "d_gains_multiplier": r.d_gains_multiplier,
               "episode_sums": torch.stack([r.episode_sums[n] for n in _abi.REWARD_NAMES], 1),
               "turn_over_timer": r.turn_over_timer,
               "xrew_sums": torch.stack([r.episode_sums[n] if n in r.episode_sums else torch.zeros(N) for n in _abi.XREWARD_NAMES], 1),
               "xrew_state": torch.cat([r.feet_air_time, r.last_contacts.float(), getattr(r, "last_contacts2", torch.zeros(N, 4)).float()], 1),
               "base_lin_vel": r.base_lin_vel, "base_ang_vel": r.base_ang_vel, "projected_gravity": r.projected_gravity,
               "measured_heights": r.measured_heights if not plane else torch.zeros(N, 187),
               "last_actions": r.last_actions, "last_last_actions": r.last_last_actions, "last_dof_vel": r.last_dof_vel,
               "contact_forces": r.contact_forces, "env_origins": r.env_origins,
               "terrain_levels": (r.terrain_levels.int() if not plane else torch.zeros(N, dtype=torch.int32))}
        ep = extras.get("episode", {})
        rec["ep_rew"] = torch.tensor([float(ep.get("rew_" + n, float("nan"))) for n in _abi.REWARD_NAMES])
        rec["ep_xrew"] = torch.tensor([float(ep.get("rew_" + n, float("nan"))) for n in _abi.XREWARD_NAMES])
        rec["ep_terrain_level_all"] = torch.tensor(float(ep.get("terrain_level_all", float("nan"))))
        outs.append({kk: vv.clone().numpy() for kk, vv in rec.items()})
    save = {"meta_N": N, "meta_K": K, "meta_seed": seed, "meta_plane": int(plane), "meta_start_counter": start_counter + 30,
            "meta_control_type": "PVT".index(control_type), "meta_only_positive": int(only_positive), "meta_heading": int(heading), "meta_overrides": np.array(repr(dict(overrides or {}, **{"sim.b200." + k: v for k, v in (b200 or {}).items()}))),
            "actions": actions.numpy()}
    for k, v in S0.items():
        save["s0_" + k] = v.numpy()
    for i, o in enumerate(outs):
        for k, v in o.items():
            save[f"out{i}_{k}"] = v
    path = os.path.join(HERE, f"env_{name}.npz")
    np.savez_compressed(path, **save)
    print("wrote", path, f"{os.path.getsize(path) / 1024:.0f} KiB; resets per step:", [int(o['reset_buf'].sum()) for o in outs])
    return path


if __name__ == "__main__":
    if "--switches" in sys.argv:      # env switches outside the GO2 defaults: velocity / torque control, only_positive_rewards
        make_case("ctrl_v_pos", plane=False, N=32, K=4, seed=13, control_type="V", only_positive=True)
        make_case("ctrl_t", plane=True, N=32, K=3, seed=14, control_type="T")
        make_case("heading", plane=False, N=48, K=6, seed=15, heading=True)
        sys.exit(0)
    if "--xrew" in sys.argv:          # every reward function the registered go2 tasks leave off (legged_robot.py:1236-1441, go2_env.py:62-68), switched on
        xr = {"orientation": -0.2, "base_height": -1.0, "dof_vel": -1e-4, "termination": -2.0, "dof_vel_limits": -0.5, "torque_limits": -0.01,
              "feet_air_time": 1.0, "stumble": -0.5, "stand_still": -0.1, "feet_contact_forces": -0.01, "similar_to_default": -0.02, "upright": 0.3,
              "legs_distance": -1.5, "x_command_hip_regular": -0.05}
        ov = {"rewards.scales." + k: v for k, v in xr.items()}
        ov.update({"rewards.max_contact_force": 20.0, "rewards.soft_dof_vel_limit": 0.1, "rewards.soft_torque_limit": 0.3, "rewards.min_legs_distance": 0.25,
                   "rewards.base_height_target": 0.3})
        make_case("xrew", plane=False, N=48, K=8, seed=21, overrides=ov)
        # init_state.turn_over (legged_robot.py:114-115,174-175,257-265,585-590,642-691): flipped resets, no contact termination, the turn_over reward scales
        # (an extra term, a default term and the default upright = 1) while |roll| > pi / 4, zero commands while the timer runs
        to = {"init_state.turn_over": True, "init_state.turn_over_proportions": [0.3, 0.4, 0.3], "rewards.turn_over_scales.torques": -2e-4,
              "rewards.turn_over_scales.dof_vel": -1e-3, "rewards.scales.orientation": -0.2, "env.episode_length_s": 0.3}
        make_case("turn_over", plane=False, N=48, K=8, seed=23, overrides=to)
        make_case("xrew_pos", plane=True, N=32, K=5, seed=22, only_positive=True, overrides={k: v for k, v in ov.items() if "hip_regular" not in k})
        sys.exit(0)
    make_case("rough", plane=False)
    make_case("plane", plane=True, N=32, K=4)
    # crosses learning iteration 20 000: the command-range curriculum widens the ranges (go2_config.py:112-124, legged_robot.py:433-446)
    make_case("cmdcur", plane=False, N=32, K=5, seed=9, start_counter=24 * 20000 - 33)
