#!/usr/bin/env python3
"""Fuzz the post-physics parity: for a range of seeds / sizes / start iterations, run the UNMODIFIED reference env Python over the oracle physics
(the harness of tests/golden/make_golden_env.py) and the oracle's own restatement from the same state, and compare every recorded step with the
tight tolerances of tests/golden_util.py.  Build container only (needs /root/reference).  The committed fixtures are three such cases; this tool
widens the net (it found the float32 terrain-column assignment at env k N / 4).

Usage: python tools/fuzz_reference_parity.py [--seeds 20:40] [--N 64] [--K 8]"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import make_golden_env as H  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="20:30")
    ap.add_argument("--N", type=int, default=64)
    ap.add_argument("--K", type=int, default=8)
    ap.add_argument("--control_types", action="store_true", help="also draw control_type V / T (violent: the first contact solver can diverge to NaN there, and V control amplifies rounding; expect tolerance-level mismatches)")
    args = ap.parse_args()
    lo, hi = (int(x) for x in args.seeds.split(":"))
    import golden_util as GU
    from oracle.oracle import OracleEnv
    tmp = tempfile.mkdtemp()
    H.HERE = tmp                       # fixtures of this run go to a scratch directory
    GU.GOLDEN = tmp
    n_bad = 0
    for seed in range(lo, hi):
        rng = np.random.default_rng(seed)
        plane = bool(rng.integers(0, 4) == 0)
        N = int(args.N + 4 * rng.integers(0, 8))
        start = int(24 * rng.integers(10, 60000) - rng.integers(0, 24))
        heading = bool(rng.integers(0, 5) == 0)
        ctrl = "PPPVT"[int(rng.integers(0, 5))] if args.control_types else "P"
        name = f"fuzz{seed}"
        H.make_case(name, plane=plane, N=N, K=args.K, seed=seed, start_counter=start, control_type=ctrl, heading=heading)
        z, A = GU.load_case(name)
        O = OracleEnv(A)
        O.common_step_counter = int(z["meta_start_counter"])
        actions = torch.from_numpy(z["actions"])
        bad_all = []
        for i in range(int(z["meta_K"])):
            O.step(actions[i])
            bad = GU.compare_step(z, i, A.tensors, tol=GU.TOL_TIGHT)
            if bad:
                bad_all.append((i, bad))
                break
        n_bad += bool(bad_all)
        print(f"seed {seed}: N={N} plane={plane} start={start} ctrl={ctrl} heading={heading} resets={[int(z[f'out{i}_reset_buf'].sum()) for i in range(args.K)]} "
              f"-> {'OK' if not bad_all else 'MISMATCH ' + str(bad_all)[:600]}", flush=True)
    print("mismatching cases:", n_bad)
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
