"""ctypes wrapper of the host emulation of the CUDA step kernel (test tooling)."""
import ctypes as C
import os
import subprocess

from go2_rl_gym_b200 import _abi

_DIR = os.path.dirname(os.path.abspath(__file__))


def load():
    subprocess.check_call(["make", "-C", _DIR, "-s"])
    lib = C.CDLL(os.path.join(_DIR, "libgo2emu.so"))
    P = C.POINTER
    lib.go2_emu_step.argtypes = [P(_abi.Go2EnvConfig), P(_abi.Go2Model), P(_abi.Go2EnvBuffers), C.c_void_p, P(_abi.Go2StepParams)]
    lib.go2_emu_reset_all.argtypes = [P(_abi.Go2EnvConfig), P(_abi.Go2Model), P(_abi.Go2EnvBuffers), P(_abi.Go2StepParams)]
    lib.go2_emu_substeps.argtypes = [P(_abi.Go2EnvConfig), P(_abi.Go2Model), P(_abi.Go2EnvBuffers), C.c_void_p, C.c_int]
    return lib


class EmuEnv:
    def __init__(self, arrays, packed=False):
        """packed: True / 8 = emulate the 8-envs-per-CTA thread map (LEGS / COLS items of 8 envs share warps), 4 = the 4-envs-per-CTA map ("Q4"),
        False = warp per env."""
        assert arrays.device.type == "cpu"
        self.A = arrays
        self.lib = load()
        self.packed = int(packed)
        self.common_step_counter = 0

    def reset_all(self):
        sp = self.A.step_params(self.common_step_counter)
        self.lib.go2_emu_set_packed(self.packed)
        self.lib.go2_emu_reset_all(C.byref(self.A.config), C.byref(self.A.model), C.byref(self.A.buffers), C.byref(sp))

    def step(self, actions, reward_curriculum=None):
        self.common_step_counter += 1
        sp = self.A.step_params(self.common_step_counter, ep_slot=self.common_step_counter % 64, reward_curriculum=reward_curriculum)
        a = actions.contiguous().float()
        self.lib.go2_emu_set_packed(self.packed)
        self.lib.go2_emu_step(C.byref(self.A.config), C.byref(self.A.model), C.byref(self.A.buffers), a.data_ptr(), C.byref(sp))
        return sp

    def substeps(self, tau, n):
        t = tau.contiguous().float()
        self.lib.go2_emu_set_packed(self.packed)
        self.lib.go2_emu_substeps(C.byref(self.A.config), C.byref(self.A.model), C.byref(self.A.buffers), t.data_ptr(), n)
