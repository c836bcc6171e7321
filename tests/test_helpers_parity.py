"""Host helpers against the reference's own functions (legged_gym/utils/helpers.py:12-125, imported from /root/reference through the isaacgym
stand-in, in a subprocess because the reference package is also called `legged_gym`): class_to_dict, update_cfg_from_args, get_load_path.
Build container only."""
import json
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, json, io, os, contextlib, tempfile, types
sys.path[:0] = [ROOT + "/tests/ref_stub", REF, REF + "/rsl_rl", ROOT]
with contextlib.redirect_stdout(io.StringIO()):
    import legged_gym.envs.go2.go2_config as ref_cfg
    import legged_gym.utils.helpers as ref_h
    import importlib
    my_cfg = importlib.import_module("go2_rl_gym_b200.envs.go2.go2_config")
    my_h = importlib.import_module("go2_rl_gym_b200.utils.helpers")
out = {}
# update_cfg_from_args on every override the scripts use
def ns(**kw):
    base = dict(num_envs=None, seed=None, max_iterations=None, resume=False, experiment_name=None, run_name=None, load_run=None, checkpoint=None,
                robogauge=None, robogauge_port=None)
    base.update(kw)
    return types.SimpleNamespace(**base)
cases = [ns(), ns(num_envs=77, seed=5, max_iterations=9), ns(resume=True, experiment_name="e", run_name="r", load_run="Oct01_x", checkpoint=1500),
         ns(robogauge=True, robogauge_port=1234)]
res = []
for a in cases:
    re_, rt = ref_h.update_cfg_from_args(ref_cfg.GO2Cfg(), ref_cfg.GO2CfgMoECTS(), a)
    me_, mt = my_h.update_cfg_from_args(my_cfg.GO2Cfg(), my_cfg.GO2CfgMoECTS(), a)
    pick = lambda e, t: [e.env.num_envs, t.seed, t.runner.max_iterations, t.runner.resume, t.runner.experiment_name, t.runner.run_name, t.runner.load_run,
                         t.runner.checkpoint, t.robogauge.enabled, t.robogauge.port]
    res.append([pick(re_, rt), pick(me_, mt)])
out["update_cfg"] = res
# get_load_path: latest run / selected run, latest / selected checkpoint, the 'exported' directory is skipped
with tempfile.TemporaryDirectory() as d:
    for run, models in (("Sep30_10-00-00_", [0, 500, 1000]), ("Oct01_09-00-00_a", [0, 50, 10000, 9500]), ("exported", [])):
        os.makedirs(os.path.join(d, run))
        for m in models:
            open(os.path.join(d, run, f"model_{m}.pt"), "w").close()
    q = []
    for kw in (dict(), dict(load_run="Sep30_10-00-00_"), dict(checkpoint=50), dict(load_run="Sep30_10-00-00_", checkpoint=500)):
        q.append([os.path.relpath(ref_h.get_load_path(d, **kw), d), os.path.relpath(my_h.get_load_path(d, **kw), d)])
    out["load_path"] = q
    try:
        ref_h.get_load_path(os.path.join(d, "nothing"))
        out["ref_raises"] = False
    except ValueError:
        out["ref_raises"] = True
    try:
        my_h.get_load_path(os.path.join(d, "nothing"))
        out["mine_raises"] = False
    except ValueError:
        out["mine_raises"] = True
# class_to_dict of a nested config
out["c2d_equal"] = ref_h.class_to_dict(ref_cfg.GO2CfgCTS()) == my_h.class_to_dict(my_cfg.GO2CfgCTS())
# legged_gym.utils.math / Logger (math.py:7-27, logger.py:5-40)
import torch
import legged_gym.utils.math as ref_m
import legged_gym.utils.logger as ref_l
my_m = importlib.import_module("go2_rl_gym_b200.utils.math")
my_l = importlib.import_module("go2_rl_gym_b200.utils.logger")
g = torch.Generator().manual_seed(0)
q = torch.nn.functional.normalize(torch.randn(64, 4, generator=g), dim=-1); v = torch.randn(64, 3, generator=g); ang = 20 * torch.randn(257, generator=g)
torch.manual_seed(3); a = ref_m.torch_rand_sqrt_float(-1.5, 2.0, (9, 2), "cpu")
torch.manual_seed(3); b = my_m.torch_rand_sqrt_float(-1.5, 2.0, (9, 2), "cpu")
inplace = ang.clone(); ret = my_m.wrap_to_pi(inplace)
out["math_equal"] = [bool(torch.allclose(ref_m.quat_apply_yaw(q, v), my_m.quat_apply_yaw(q, v), atol=1e-6)), bool(torch.allclose(ref_m.wrap_to_pi(ang.clone()), ret, atol=1e-6)) and ret.data_ptr() == inplace.data_ptr(),
                     bool(torch.allclose(a, b, atol=1e-7))]
logs = []
for L in (ref_l.Logger(0.02), my_l.Logger(0.02)):
    L.log_states({"x": 1.0, "y": 2.0}); L.log_state("x", 3.0)
    L.log_rewards({"rew_a": torch.tensor(0.5), "rew_b": torch.tensor(-1.0), "terrain_level": torch.tensor(3.0)}, 4)
    L.log_rewards({"rew_a": torch.tensor(1.5)}, 2)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        L.print_rewards()
    logs.append([dict(L.state_log), dict(L.rew_log), L.num_episodes, buf.getvalue()])
    L.reset()
    logs[-1].append([len(L.state_log), len(L.rew_log)])
out["logger_equal"] = logs[0] == logs[1]
print("RESULT" + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "legged_gym")), reason="needs the reference tree (runs in the build container)")
def test_helpers_behave_like_the_reference():
    code = SCRIPT.replace("ROOT", repr(ROOT)).replace("REF", repr(REF))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    assert line, out.stderr[-3000:]
    res = json.loads(line[0][len("RESULT"):])
    for ref, mine in res["update_cfg"]:
        assert ref == mine, (ref, mine)
    for ref, mine in res["load_path"]:
        assert ref == mine, (ref, mine)
    assert res["load_path"][0][1] == os.path.join("Sep30_10-00-00_", "model_1000.pt")      # runs sort by name: 'Sep30' > 'Oct01'
    assert res["ref_raises"] and res["mine_raises"]
    assert res["c2d_equal"]
    assert res["math_equal"] == [True, True, True] and res["logger_equal"]
