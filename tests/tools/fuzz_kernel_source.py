#!/usr/bin/env python3
"""Fuzz the step kernel's OWN SOURCE (csrc/env_step_core.cuh compiled with g++, tests/emu) against the CPU oracle: random env count (partially
filled CTAs), thread map (warp per env / P2 / Q4), terrain or plane, config switches off the training defaults, control type, heading commands,
relaxed solver and state guard; one step at a time from the oracle's state (flags exact, floats within tests/golden_util.TOL).  Complements
tests/tools/fuzz_reference_parity.py (reference vs oracle): together they tie the shipped kernel logic to the reference's Python.

Usage: python tests/tools/fuzz_kernel_source.py [--seeds 0:40] [--steps 25]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "tools")]
from go2_rl_gym_b200.envs.env_arrays import EnvArrays  # noqa: E402
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg  # noqa: E402
from cuda_util import copy_state  # noqa: E402
from emu.emu import EmuEnv  # noqa: E402
from golden_util import TOL  # noqa: E402
from oracle.oracle import OracleEnv  # noqa: E402
from fuzz_reference_parity import SWITCHES  # noqa: E402

EXACT = ("reset_buf", "time_out_buf", "episode_length_buf", "terrain_levels", "last_is_limit_vel", "stop_heading")
FLOATS = ("obs_buf", "privileged_obs_buf", "rew_buf", "root_states", "dof_pos", "dof_vel", "commands", "episode_sums", "torques", "xrew_sums", "xrew_state",
          "turn_over_timer")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="0:40")
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--f3", action="store_true", help="also draw the reward functions outside the GO2 defaults and init_state.turn_over (fuzz_reference_parity.f3_overrides)")
    args = ap.parse_args()
    lo, hi = (int(x) for x in args.seeds.split(":"))
    n_bad = 0
    for seed in range(lo, hi):
        rng = np.random.default_rng(seed)
        N = int(rng.integers(9, 90))
        packed = [False, True, 4, 14][int(rng.integers(0, 4))]
        cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.seed = seed
        cfg.terrain.mesh_type = "plane" if rng.integers(0, 4) == 0 else "heightfield"
        ov = {}
        for path, vals in SWITCHES.items():
            if rng.integers(0, 3) == 0:
                ov[path] = vals[int(rng.integers(0, len(vals)))]
        if args.f3:
            from fuzz_reference_parity import f3_overrides
            ov.update(f3_overrides(rng))
        for path, val in ov.items():
            node, parts = cfg, path.split(".")
            for p in parts[:-1]:
                node = getattr(node, p)
            setattr(node, parts[-1], val)
        cfg.commands.heading_command = bool(rng.integers(0, 4) == 0)
        cfg.control.control_type = "PPPVT"[int(rng.integers(0, 5))]
        cfg.rewards.only_positive_rewards = bool(rng.integers(0, 5) == 0)
        relaxed = bool(rng.integers(0, 2)) or cfg.control.control_type != "P"      # V / T slam the joints into their stops: the first solver diverges there
        if relaxed:
            cfg.sim.b200.limit_relax, cfg.sim.b200.contact_relax, cfg.sim.b200.limit_erp = 0.5, 0.7, 0.8
            cfg.sim.b200.state_guard = int(rng.integers(0, 2))
        tol = dict(TOL)
        if cfg.control.control_type == "V":      # kd (qd - last_qd) / sim_dt amplifies operation-order noise (tests/test_emu_cpu.py)
            tol.update(dof_vel=(1e-3, 0.3), torques=(1e-3, 0.3), privileged_obs_buf=(1e-3, 3e-2), obs_buf=(1e-3, 3e-2), root_states=(1e-4, 5e-3),
                       dof_pos=(1e-4, 2e-3), rew_buf=(1e-3, 2e-3), episode_sums=(1e-3, 2e-3))
        Ac, Ae = EnvArrays(cfg, "cpu", seed=seed), EnvArrays(cfg, "cpu", seed=seed)
        orc, env = OracleEnv(Ac), EmuEnv(Ae, packed=packed)
        orc.common_step_counter = env.common_step_counter = int(24 * rng.integers(10, 60000))
        orc.reset_all(); env.reset_all()
        g = torch.Generator().manual_seed(seed)
        Ac.tensors["episode_length_buf"].copy_(torch.randint(0, int(Ac.max_episode_length), (N,), generator=g).int())
        copy_state(Ac.tensors, Ae.tensors)
        bad, n_reset = None, 0
        scale = float(rng.choice([0.3, 0.6, 1.0]))
        for step in range(args.steps):
            a = scale * torch.randn(N, 12, generator=g)
            orc.step(a); env.step(a)
            for k in EXACT:
                if not torch.equal(Ac.tensors[k], Ae.tensors[k]):
                    bad = (step, k, "exact")
            for k in FLOATS:
                rtol, atol = tol.get(k, tol["default"])
                x, y = Ae.tensors[k].numpy(), Ac.tensors[k].numpy()
                if not np.allclose(x, y, rtol=rtol, atol=atol, equal_nan=True):
                    bad = (step, k, float(np.nanmax(np.abs(x - y))))
            if bad:
                break
            n_reset += int(Ac.tensors["reset_buf"].sum())
            copy_state(Ac.tensors, Ae.tensors)
            Ae.tensors["stop_heading"].copy_(Ac.tensors["stop_heading"])
        n_bad += bad is not None
        print(f"seed {seed}: N={N} map={packed} {cfg.terrain.mesh_type} ctrl={cfg.control.control_type} heading={cfg.commands.heading_command} relaxed={relaxed} "
              f"guard={cfg.sim.b200.state_guard} act={scale} switches={len(ov)} resets={n_reset} -> {'OK' if bad is None else 'MISMATCH ' + str(bad)}", flush=True)
    print("mismatching cases:", n_bad)
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
