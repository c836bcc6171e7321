"""The CUDA library behind the same call sequence as OracleEnv / EmuEnv (GPU tests call the product through its C ABI)."""
import ctypes as C

import torch

from go2_rl_gym_b200 import _abi

STATE = ["root_states", "dof_pos", "dof_vel", "actions", "last_actions", "last_last_actions", "last_dof_vel", "torques",
         "commands", "commands_resampling_step", "commands_xy_accumulation", "last_is_limit_vel", "episode_length_buf",
         "terrain_levels", "env_origins", "max_move_distance", "motor_strengths", "motor_zero_offsets",
         "p_gains_multiplier", "d_gains_multiplier", "episode_sums", "contact_forces", "reset_buf", "time_out_buf",
         "stop_heading", "xrew_sums", "xrew_state", "turn_over_timer"]


class CudaEnv:
    def __init__(self, arrays, mode=None, lib_path=None):
        assert arrays.device.type == "cuda"
        self.A = arrays
        self.lib = _abi.load_library(lib_path)
        self.h = C.c_void_p()
        _abi.check(self.lib.go2_env_create(C.byref(arrays.config), C.byref(arrays.model), C.byref(arrays.buffers), C.byref(self.h)), self.lib)
        self.common_step_counter = 0
        if mode is not None:
            _abi.check(self.lib.go2_env_set_step_mode(self.h, mode.encode()), self.lib)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.go2_env_destroy(self.h)
            self.h = None

    def reset_all(self):
        sp = self.A.step_params(self.common_step_counter)
        _abi.check(self.lib.go2_env_reset_all(self.h, C.byref(sp), None), self.lib)

    def step(self, actions, reward_curriculum=None):
        self.common_step_counter += 1
        sp = self.A.step_params(self.common_step_counter, ep_slot=self.common_step_counter % 64, reward_curriculum=reward_curriculum)
        a = actions.to(self.A.device).contiguous().float()
        _abi.check(self.lib.go2_env_step(self.h, a.data_ptr(), C.byref(sp), None), self.lib)
        torch.cuda.synchronize()
        return sp

    def substeps(self, tau, n):
        t = tau.to(self.A.device).contiguous().float()
        _abi.check(self.lib.go2_env_substeps(self.h, t.data_ptr(), n, None), self.lib)
        torch.cuda.synchronize()


def copy_state(src_tensors, dst_tensors):
    for k in STATE:
        dst_tensors[k].copy_(src_tensors[k].to(dst_tensors[k].device))
