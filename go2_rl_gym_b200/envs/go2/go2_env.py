"""Go2Robot (reference: legged_gym/envs/go2/go2_env.py:7-68).  The Go2 observation layout (45 / 263), its noise vector
and the two Go2 reward terms are part of the fused kernel (csrc/env_step_core.cuh), so the subclass adds nothing."""
from ..base.legged_robot import LeggedRobot


class Go2Robot(LeggedRobot):
    pass
