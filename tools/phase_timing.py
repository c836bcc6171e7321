#!/usr/bin/env python3
"""Per-phase cycle counts of one CTA of the packed step kernel (tuning aid; needs tools/build_timing_lib.sh).
GO2_B200_LIB=go2_rl_gym_b200/libgo2b200_timing.so python tools/phase_timing.py [--num_envs 4096]"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from go2_rl_gym_b200 import _abi
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
from go2_rl_gym_b200.envs.go2.go2_env import Go2Robot

ap = argparse.ArgumentParser()
ap.add_argument("--num_envs", type=int, default=4096)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--mode", default="H14")
ap.add_argument("--raw", action="store_true", help="print every interval between CTA barriers")
args = ap.parse_args()
cfg = GO2Cfg(); cfg.env.num_envs = args.num_envs; cfg.terrain.mesh_type = "heightfield"
env = Go2Robot(cfg, None, None, "cuda:0", True)
env.set_step_mode(args.mode)
env.reset()
lib = _abi.load_library()
lib.go2_debug_phase_clocks.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_longlong * 512)()
acc, n_ok = None, 0
for i in range(args.steps):
    env.step(0.5 * torch.randn(args.num_envs, 12, device="cuda"))
    n = lib.go2_debug_phase_clocks(buf, 512)
    d = [buf[k + 1] - buf[k] for k in range(n - 1)]
    if i >= 5:
        acc = d if acc is None else [a + b for a, b in zip(acc, d)]
        n_ok += 1
d = [a / n_ok for a in acc]
# CTA barriers of the packed map only sit at role switches: per substep  [WIDE: (load | S12) + torques + S1] | [LEGS: ABA passes, base
# inverse, mobility] | [WIDE: narrow phase + contact solve 0] | [LEGS: sweep 0] | 3 x ([WIDE: contact solve] | [LEGS: sweep]) ; after the last
# one [WIDE: S12 + post-physics + store] up to the kernel's end
names = []
for sb in range(4):
    names += [f"s{sb}.wide_pre", f"s{sb}.legs_aba", f"s{sb}.narrow_contact0", f"s{sb}.legs_sweep"]
    for it in range(3):
        names += [f"s{sb}.contact", f"s{sb}.legs_sweep"]
post = ["post.S12_last", "post.state_guard", "post.feet_kin", "post.heights_prelude", "post.partials_jterm", "post.termination", "post.fterm", "post.reward",
        "post.episode_sums", "post.reset", "post.push_atomics", "post.obsrow", "post.obs_out", "post.lact", "post.store_state", "post.cta_exit"]
names += [post[k] if k < len(post) else f"post{k}" for k in range(len(d) - len(names))]
tot = sum(d)
if args.raw:
    for nm, v in zip(names, d):
        print(f"  {nm:20s} {v:9.0f}")
print(f"phases {len(d)}  total {tot:.0f} cycles per CTA-step")
agg = {}
for nm, v in zip(names, d):
    key = nm.split(".")[-1] if nm.startswith("s") else "post"
    agg[key] = agg.get(key, 0.0) + v
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"{k:12s} {v:10.0f} cycles  {100 * v / tot:5.1f} %")
