"""ctypes wrapper of the CPU oracle (test infrastructure; see go2_oracle.cpp header)."""
import ctypes as C
import os
import subprocess

from go2_rl_gym_b200 import _abi

_DIR = os.path.dirname(os.path.abspath(__file__))


def build():
    subprocess.check_call(["make", "-C", _DIR, "-s"])


def load(double=False):
    path = os.path.join(_DIR, "libgo2oracle_f64.so" if double else "libgo2oracle.so")
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    P = C.POINTER
    lib.go2_oracle_step.argtypes = [P(_abi.Go2EnvConfig), P(_abi.Go2Model), P(_abi.Go2EnvBuffers), C.c_void_p, P(_abi.Go2StepParams)]
    lib.go2_oracle_reset_all.argtypes = [P(_abi.Go2EnvConfig), P(_abi.Go2Model), P(_abi.Go2EnvBuffers), P(_abi.Go2StepParams)]
    lib.go2_oracle_substeps.argtypes = [P(_abi.Go2EnvConfig), P(_abi.Go2Model), P(_abi.Go2EnvBuffers), C.c_void_p, C.c_int]
    lib.go2_oracle_feet.argtypes = [P(_abi.Go2EnvConfig), P(_abi.Go2Model), P(_abi.Go2EnvBuffers)]
    lib.go2_oracle_philox.argtypes = [C.c_uint32] * 6 + [P(C.c_uint32)]
    return lib


class OracleEnv:
    """The oracle behind the same call sequence the product env uses (tests / cpu_baseline only)."""

    def __init__(self, arrays, double=False):
        assert arrays.device.type == "cpu"
        self.A = arrays
        self.lib = load(double)
        self.common_step_counter = 0

    def reset_all(self):
        sp = self.A.step_params(self.common_step_counter)
        self.lib.go2_oracle_reset_all(C.byref(self.A.config), C.byref(self.A.model), C.byref(self.A.buffers), C.byref(sp))

    def step(self, actions, reward_curriculum=None):
        self.common_step_counter += 1
        sp = self.A.step_params(self.common_step_counter, ep_slot=self.common_step_counter % 64, reward_curriculum=reward_curriculum)
        a = actions.contiguous().float()
        self.lib.go2_oracle_step(C.byref(self.A.config), C.byref(self.A.model), C.byref(self.A.buffers), a.data_ptr(), C.byref(sp))
        return sp

    def feet(self):
        self.lib.go2_oracle_feet(C.byref(self.A.config), C.byref(self.A.model), C.byref(self.A.buffers))

    def substeps(self, tau, n):
        t = tau.contiguous().float()
        self.lib.go2_oracle_substeps(C.byref(self.A.config), C.byref(self.A.model), C.byref(self.A.buffers), t.data_ptr(), n)
