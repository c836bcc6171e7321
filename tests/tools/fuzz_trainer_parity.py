#!/usr/bin/env python3
"""Fuzz the trainer's host logic: random CTS-family variant, widths, env counts (incl. counts that do not divide evenly into mini-batches) and
algorithm options (clipped / unclipped value loss, fixed / adaptive schedule, entropy and load-balance coefficients, epochs, mini-batches), the
REFERENCE's rsl_rl run side by side with this package's classes over the emulated C ABI (tests/cts_util.side_by_side).  Build container only.

Usage: python tests/tools/fuzz_trainer_parity.py [--seeds 0:30]"""
import argparse
import os
import sys
import traceback

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import emu_rl  # noqa: E402
from cts_util import side_by_side, side_by_side_ppo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="0:30")
    args = ap.parse_args()
    lo, hi = (int(x) for x in args.seeds.split(":"))
    n_bad = 0
    for seed in range(lo, hi):
        rng = np.random.default_rng(seed)
        variant = ["ppo", "cts", "moe_cts", "moe_ng_cts", "ac_moe_cts", "dual_moe_cts", "mcp_cts"][int(rng.integers(0, 7))]
        w = lambda: int(rng.choice([16, 32, 64, 136, 160]))
        policy = dict(actor_hidden_dims=[w(), w(), w()], critic_hidden_dims=[w(), w(), w()], teacher_encoder_hidden_dims=[w(), w()],
                      student_encoder_hidden_dims=[w(), w()] + ([w()] if variant in ("moe_cts", "dual_moe_cts") else []),
                      latent_dim=int(rng.choice([16, 32])), activation="elu", norm_type="l2norm")
        if variant != "mcp_cts":
            policy["init_noise_std"] = float(rng.choice([0.5, 1.0]))
        if variant in ("moe_cts", "ac_moe_cts", "dual_moe_cts"):
            policy["expert_num"] = int(rng.choice([4, 8]))
        if variant in ("moe_ng_cts", "mcp_cts"):
            policy["student_expert_num"] = int(rng.choice([4, 8]))
        alg = dict(value_loss_coef=float(rng.choice([0.5, 1.0])), use_clipped_value_loss=bool(rng.integers(0, 2)), clip_param=0.2,
                   entropy_coef=float(rng.choice([0.0, 0.01])), num_learning_epochs=int(rng.integers(1, 4)), num_mini_batches=int(rng.choice([1, 2, 4])),
                   learning_rate=1e-3, student_encoder_learning_rate=float(rng.choice([1e-3, 3e-4])), schedule=str(rng.choice(["adaptive", "fixed"])),
                   gamma=0.99, lam=0.95, desired_kl=0.01, max_grad_norm=float(rng.choice([0.5, 1.0])), teacher_env_ratio=0.75)
        if variant in ("moe_cts", "moe_ng_cts", "ac_moe_cts", "dual_moe_cts"):
            alg["load_balance_coef"] = float(rng.choice([0.0, 0.01, 0.1]))
        N, T = int(rng.choice([8, 12, 16, 20, 32])), int(rng.choice([3, 4, 6, 8]))
        gemm = str(rng.choice(["tc", "simt"]))
        mp = pytest.MonkeyPatch()
        try:
            emu_rl.install(mp)
            mp.setenv("GO2_GEMM", gemm)
            if variant == "ppo":
                pol = dict(actor_hidden_dims=policy["actor_hidden_dims"], critic_hidden_dims=policy["critic_hidden_dims"], activation="elu",
                           init_noise_std=policy["init_noise_std"])
                akw = {k: v for k, v in alg.items() if k not in ("student_encoder_learning_rate", "teacher_env_ratio", "load_balance_coef")}
                r = side_by_side_ppo(pol, akw, N=N, T=T, seed=seed, monkeypatch=mp)
            else:
                r = side_by_side(variant, policy, alg, N=N, T=T, seed=seed, monkeypatch=mp)
            ok = r["act"] < 5e-5 and r["returns"] < 5e-5 and r["adv"] < 5e-4 and r["loss"] < 5e-4 and r["lr"] < 1e-9 and r["update_rel"] < 2e-2
            msg = " ".join(f"{k}={v:.1e}" for k, v in r.items())
        except Exception as e:  # noqa: BLE001 - report and continue
            ok, msg = False, f"{type(e).__name__}: {e}\\n" + traceback.format_exc(limit=4)
        finally:
            mp.undo()
        n_bad += not ok
        print(f"seed {seed}: {variant} N={N} T={T} nb={alg['num_mini_batches']} ep={alg['num_learning_epochs']} {gemm} clipV={alg['use_clipped_value_loss']} "
              f"{alg['schedule']} -> {'OK ' if ok else 'MISMATCH '}{msg}", flush=True)
    print("mismatching cases:", n_bad)
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
