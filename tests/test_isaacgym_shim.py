"""The reference's own entry script, unmodified, resolves every import against this repo (missing item 8): `import isaacgym` (shim),
`from legged_gym.envs import *`, `from legged_gym.utils import get_args, task_registry`.  Runs in a subprocess: other tests put tests/ref_stub's
Isaac Gym stand-in (for the reference's ENV code) on sys.path under the same module name."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TRAIN = "/root/reference/legged_gym/scripts/train.py"


def _run(code):
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd="/tmp")


def test_import_isaacgym_resolves_to_the_shim():
    r = _run("import isaacgym, os; print(os.path.dirname(isaacgym.__file__))")
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == os.path.join(ROOT, "isaacgym")


@pytest.mark.skipif(not os.path.exists(REF_TRAIN), reason="reference tree not present")
def test_reference_train_script_imports_and_builds_its_arguments_unmodified():
    code = (f"import runpy, sys; sys.argv = ['train.py', '--task=go2_moe_cts', '--num_envs', '64', '--headless', '--max_iterations', '3']\n"
            f"ns = runpy.run_path({REF_TRAIN!r}, run_name='ref_train')\n"
            "import legged_gym, rsl_rl, isaacgym\n"
            "args = ns['get_args']()\n"
            "cfgs = ns['task_registry'].get_cfgs(args.task)\n"
            "print(callable(ns['train']), args.task, args.num_envs, type(cfgs[1]).__name__, legged_gym.__file__)")
    r = _run(code)
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout.strip().splitlines()[-1].split()
    assert out[0] == "True" and out[1] == "go2_moe_cts" and out[2] == "64"
    assert out[4].startswith(ROOT)          # the reference script is running against THIS repo's legged_gym package
