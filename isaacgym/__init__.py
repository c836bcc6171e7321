"""Import shim: `import isaacgym` is the first import of the reference's entry scripts (legged_gym/scripts/train.py:6, play.py:6) because Isaac Gym
must be loaded before torch.  This package replaces the simulator itself (the fused step kernel behind include/go2_b200.h), so there is nothing to
load: the module exists so that the UNMODIFIED reference scripts run against this repo's `legged_gym` / `rsl_rl` packages:

    PYTHONPATH=/path/to/this/repo python /path/to/reference/legged_gym/scripts/train.py --task=go2 --headless

It deliberately exposes no gymapi / gymtorch / gymutil: nothing on the product path calls the Isaac Gym API (~45 `self.gym.*`
calls are replaced by the C ABI).  The tests' stand-in for running the REFERENCE's own env code lives in tests/ref_stub/ and is separate."""
__all__ = []
