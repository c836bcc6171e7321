"""The C-ABI library loads and exports every symbol include/go2_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["libgo2b200.so", "libgo2b200_relaxed.so"])
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


def test_create_rejects_solver_settings_the_build_does_not_carry():
    """The sm_100a library of this round is built without the relaxed solver (GO2_RELAXED_SOLVER=0): go2_env_create must refuse
    limit_relax / contact_relax instead of silently running the first solver (checked before any CUDA call, so it runs without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    from go2_rl_gym_b200 import _abi
    lib = _abi.load_library()
    cfg, mdl, buf = _abi.Go2EnvConfig(), _abi.Go2Model(), _abi.Go2EnvBuffers()
    cfg.num_envs, cfg.limit_relax, cfg.contact_relax = 8, 0.5, 0.7
    h = ctypes.c_void_p()
    rc = lib.go2_env_create(ctypes.byref(cfg), ctypes.byref(mdl), ctypes.byref(buf), ctypes.byref(h))
    assert rc != 0 and b"relaxed solver" in lib.go2_last_error()
    cfg.limit_relax, cfg.contact_relax, cfg.state_guard = 0.0, 1.0, 1
    rc = lib.go2_env_create(ctypes.byref(cfg), ctypes.byref(mdl), ctypes.byref(buf), ctypes.byref(h))
    assert rc != 0 and b"state guard" in lib.go2_last_error()
    # the second build (libgo2b200_relaxed.so) accepts them: it gets past this check to the topology check of the (empty) model
    relaxed = _abi.load_library(os.path.join(ROOT, "go2_rl_gym_b200", "libgo2b200_relaxed.so"))
    cfg.limit_relax, cfg.contact_relax = 0.5, 0.7
    rc = relaxed.go2_env_create(ctypes.byref(cfg), ctypes.byref(mdl), ctypes.byref(buf), ctypes.byref(h))
    assert rc != 0 and b"joint axes" in relaxed.go2_last_error()


@pytest.mark.parametrize("name", ["libgo2b200.so", "libgo2b200_relaxed.so"])
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
