"""The oracle's post-physics restatement against the reference's own Python (fixtures made by
tests/golden/make_golden_env.py from /root/reference, legged_robot.py:60-142 et al.)."""
import numpy as np
import pytest
import torch

from golden_util import load_case, compare_step, TOL_TIGHT
from oracle.oracle import OracleEnv


@pytest.mark.parametrize("name", ["rough", "plane", "cmdcur", "ctrl_v_pos", "ctrl_t", "heading"])
def test_oracle_matches_reference_env(name):
    z, A = load_case(name)
    O = OracleEnv(A)
    O.common_step_counter = int(z["meta_start_counter"])
    actions = torch.from_numpy(z["actions"])
    n_reset = 0
    for i in range(int(z["meta_K"])):
        sp = O.step(actions[i])
        bad = compare_step(z, i, A.tensors, tol=TOL_TIGHT)
        assert not bad, f"step {i}: {bad}"
        n_reset += int(z[f"out{i}_reset_buf"].sum())
        ep = z[f"out{i}_ep_rew"]
        if z[f"out{i}_reset_buf"].sum() > 0:
            st = A.tensors["ep_stats"][sp.ep_slot].numpy()
            assert np.allclose(st[:14], ep, rtol=1e-4, atol=1e-6)
            assert np.isclose(st[14], float(z[f"out{i}_ep_terrain_level_all"]), atol=1e-6)
    assert n_reset >= 3, "fixture must exercise resets"
    if name == "ctrl_v_pos":  # velocity control + only_positive_rewards (legged_robot.py:612-613,266-267): the clip is really exercised
        assert all(float(z[f"out{i}_rew_buf"].min()) >= 0.0 for i in range(int(z["meta_K"]))) and bool((z["out0_rew_buf"] == 0.0).any())
    if name == "cmdcur":      # the recorded window crosses learning iteration 20 000: the widened command ranges are in force at its end
        assert float(A.tensors["env_command_ranges"][:, 1].max()) == 1.0 and float(A.tensors["env_command_ranges"][:, 5].max()) == 1.5
