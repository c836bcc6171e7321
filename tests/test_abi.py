"""The C-ABI library loads and exports every symbol include/go2_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["libgo2b200.so"])
def test_library_exports_header_symbols(name):
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "go2_rl_gym_b200", name))
    hdr = open(os.path.join(ROOT, "include", "go2_b200.h")).read()
    names = sorted(set(re.findall(r"\b(go2_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 8
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C compiler's layout."""
    import subprocess, tempfile
    from go2_rl_gym_b200 import _abi
    src = '#include <stdio.h>\n#include "go2_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(Go2Model), sizeof(Go2EnvConfig), sizeof(Go2StepParams), sizeof(Go2EnvBuffers));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        out = subprocess.check_output([os.path.join(d, "s")]).split()
    got = [ctypes.sizeof(x) for x in (_abi.Go2Model, _abi.Go2EnvConfig, _abi.Go2StepParams, _abi.Go2EnvBuffers)]
    assert got == [int(x) for x in out]


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
    from go2_rl_gym_b200.envs.go2.go2_env import Go2Robot
    cfg = GO2Cfg(); cfg.env.num_envs = 4
    with pytest.raises(RuntimeError):
        Go2Robot(cfg, None, None, "cuda:0", True)


def test_create_accepts_every_solver_setting_and_env_switch():
    """ONE library carries the convergent solver, the state guard, control types V / T, only_positive_rewards and heading commands: go2_env_create
    gets past the settings to the topology check of the (empty) model for each of them (checked before any CUDA call, so it runs without a GPU),
    and still rejects what no build serves."""
    import __graft_entry__ as ge
    ge.build()
    from go2_rl_gym_b200 import _abi
    lib = _abi.load_library()
    assert not os.path.exists(os.path.join(ROOT, "go2_rl_gym_b200", "libgo2b200_relaxed.so"))
    mdl, buf = _abi.Go2Model(), _abi.Go2EnvBuffers()
    h = ctypes.c_void_p()
    for kw in (dict(limit_relax=0.5, contact_relax=0.7, state_guard=1), dict(limit_relax=0.0, contact_relax=1.0, state_guard=0),
               dict(control_type=1), dict(control_type=2, only_positive_rewards=1)):
        cfg = _abi.Go2EnvConfig()
        cfg.num_envs = 8
        for k, v in kw.items():
            setattr(cfg, k, v)
        rc = lib.go2_env_create(ctypes.byref(cfg), ctypes.byref(mdl), ctypes.byref(buf), ctypes.byref(h))
        assert rc != 0 and b"joint axes" in lib.go2_last_error(), (kw, lib.go2_last_error())
    cfg = _abi.Go2EnvConfig()
    cfg.num_envs, cfg.control_type = 8, 3
    assert lib.go2_env_create(ctypes.byref(cfg), ctypes.byref(mdl), ctypes.byref(buf), ctypes.byref(h)) != 0 and b"control_type" in lib.go2_last_error()
    cfg.control_type, cfg.heading_command = 0, 1            # heading commands need their two per-env arrays
    assert lib.go2_env_create(ctypes.byref(cfg), ctypes.byref(mdl), ctypes.byref(buf), ctypes.byref(h)) != 0 and b"ext_stop_heading" in lib.go2_last_error()


@pytest.mark.parametrize("name", ["libgo2b200.so"])
def test_every_bound_trainer_entry_point_is_exported_and_declared(name):
    """rl/_ops.py binds its whole signature table when the library is first used: one missing symbol would take the trainer (and bench.py) down."""
    import __graft_entry__ as ge
    ge.build()
    from go2_rl_gym_b200.rl import _ops
    lib = ctypes.CDLL(os.path.join(ROOT, "go2_rl_gym_b200", name))
    hdr = open(os.path.join(ROOT, "include", "go2_b200.h")).read()
    for sym, sig in _ops._SIGS.items():
        assert hasattr(lib, sym), sym
        m = re.search(r"\bint\s+" + sym + r"\s*\(([^;]*)\)\s*;", hdr)
        assert m, f"{sym} is not declared in include/go2_b200.h"
        params = [a.strip() for a in m.group(1).split(",") if a.strip()]
        assert len(params) == len(sig), (sym, m.group(1))                       # argument count incl. the stream
        for prm, ct in zip(params, sig):                                           # ... and each argument's C type against its ctypes type
            prm = re.sub(r"/\*.*?\*/", "", prm).strip()
            if "*" in prm:
                want = ctypes.c_void_p
            else:
                base = " ".join(prm.split()[:-1]).replace("const ", "").strip()
                want = {"int": ctypes.c_int, "long": ctypes.c_long, "float": ctypes.c_float, "double": ctypes.c_double, "uint64_t": ctypes.c_uint64,
                        "uint32_t": ctypes.c_uint32}[base]
            assert ct is want, (sym, prm, ct)
