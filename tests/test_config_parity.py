"""Config layer parity: the nested-class config trees of every registered go2 task (env + train cfgs, flattened
by class_to_dict) must equal the REFERENCE's own trees, imported from /root/reference through the isaacgym stand-in.  Runs in the build
container only (the reference tree does not travel); in a subprocess because the reference package is also called `legged_gym`."""
import json
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, json, io, contextlib
sys.path[:0] = [ROOT + "/tests/ref_stub", REF, REF + "/rsl_rl", ROOT]
with contextlib.redirect_stdout(io.StringIO()):
    import legged_gym.envs.go2.go2_config as ref
    from legged_gym.utils.helpers import class_to_dict as ref_c2d
    import importlib
    mine = importlib.import_module("go2_rl_gym_b200.envs.go2.go2_config")
    from go2_rl_gym_b200.utils.cfg_dict import class_to_dict as c2d
assert "reference" in ref.__file__
def diff(a, b, path=""):
    out = []
    for k in sorted(set(a) | set(b)):
        if k not in a: out.append([path + k, "only here", repr(b[k])])
        elif k not in b: out.append([path + k, "only in reference", repr(a[k])])
        elif isinstance(a[k], dict) and isinstance(b[k], dict): out += diff(a[k], b[k], path + k + ".")
        elif a[k] != b[k]: out.append([path + k, repr(a[k]), repr(b[k])])
    return out
res = {}
for name in ("GO2Cfg", "GO2CfgPPO", "GO2CfgCTS", "GO2CfgMoECTS", "GO2CfgMoENGCTS", "GO2CfgMCPCTS", "GO2CfgACMoECTS", "GO2CfgDualMoECTS"):
    res[name] = diff(ref_c2d(getattr(ref, name)()), c2d(getattr(mine, name)()))
print("RESULT" + json.dumps(res))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "legged_gym")), reason="needs the reference tree (runs in the build container)")
def test_config_trees_equal_the_reference():
    code = SCRIPT.replace("ROOT", repr(ROOT)).replace("REF", repr(REF))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    assert line, out.stderr[-2000:]
    res = json.loads(line[0][len("RESULT"):])
    # the only additions allowed: this package's own physics-solver block (the reference delegates those to PhysX)
    for name, d in res.items():
        d = [x for x in d if not x[0].startswith("sim.b200")]
        assert not d, (name, d[:10])
    assert any(x[0].startswith("sim.b200") for x in res["GO2Cfg"])


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "legged_gym")), reason="needs the reference tree (runs in the build container)")
def test_registered_tasks_equal_the_reference():
    src = open(os.path.join(REF, "legged_gym", "envs", "__init__.py")).read()
    import re
    ref_tasks = set(re.findall(r'task_registry\.register\(\s*"([a-z0-9_]+)"', src))
    from go2_rl_gym_b200.envs import task_registry
    mine = set(task_registry.task_classes)
    assert mine == ref_tasks and len(mine) == 7, (mine, ref_tasks)       # legged_gym/envs/__init__.py:9-15


def test_command_range_curriculum_after_a_late_resume():
    """A run resumed past BOTH command-range boundaries (go2_config.py:112-124) trains on the ranges of the LAST boundary — what an uninterrupted run
    would be using — whereas the reference's backward walk inside `_resample_commands` (legged_robot.py:433-446) ends on the earliest entry
    (DESIGN.md section 6).  An uninterrupted run crosses the boundaries one at a time and agrees with the reference at each."""
    from go2_rl_gym_b200.envs.env_arrays import EnvArrays
    from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
    cfg = GO2Cfg(); cfg.env.num_envs = 8; cfg.terrain.mesh_type = "plane"
    entries = sorted(cfg.commands.command_range_curriculum, key=lambda e: e["iter"])
    assert len(entries) >= 2
    late = EnvArrays(cfg, "cpu", seed=1)
    late.step_params(24 * (entries[-1]["iter"] + 5))                          # resumed after the last boundary
    for key in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw"):
        assert list(late.command_ranges[key]) == list(entries[-1][key]), key
    walk = EnvArrays(cfg, "cpu", seed=1)
    for e in entries:                                                           # uninterrupted: one boundary at a time
        walk.step_params(24 * e["iter"])
        for key in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw"):
            assert list(walk.command_ranges[key]) == list(e[key]), (e["iter"], key)
    assert late.tensors["env_command_ranges"].equal(walk.tensors["env_command_ranges"])


def test_step_params_memo_equals_a_fresh_block():
    """EnvArrays.step_params reuses the previous call's block while the iteration and the curriculum scales stand (the host-buffer loop calls it every
    step): the memoised block must equal a freshly built one byte for byte, across iteration boundaries, reward-curriculum changes and the
    command-range curriculum boundary at learning iteration 20 000."""
    from go2_rl_gym_b200.envs.env_arrays import EnvArrays
    from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
    cfg = GO2Cfg(); cfg.env.num_envs = 16; cfg.terrain.mesh_type = "plane"
    A, B = EnvArrays(cfg, "cpu", seed=1), EnvArrays(cfg, "cpu", seed=1)
    for c in list(range(24 * 99, 24 * 101 + 3)) + list(range(24 * 19999, 24 * 20001 + 2)):
        rc = {"lin_vel_z": 1.0 - (c // 24 % 7) / 7.0}
        a = A.step_params(c, ep_slot=c % 64, reward_curriculum=rc)
        B._sp_memo = None
        b = B.step_params(c, ep_slot=c % 64, reward_curriculum=rc)
        assert bytes(a) == bytes(b), c
        assert a.common_step_counter == c and a.ep_slot == c % 64
