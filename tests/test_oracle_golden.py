"""The oracle's post-physics restatement against the reference's own Python (fixtures made by
tests/golden/make_golden_env.py from /root/reference, legged_robot.py:60-142 et al.)."""
import numpy as np
import pytest
import torch

from golden_util import load_case, compare_step, TOL_TIGHT
from oracle.oracle import OracleEnv


@pytest.mark.parametrize("name", ["rough", "plane", "cmdcur", "ctrl_v_pos", "ctrl_t", "heading", "xrew", "xrew_pos", "turn_over"])
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
            if name.startswith("xrew") or name == "turn_over":    # extras["episode"]["rew_<name>"] of the terms outside the GO2 defaults
                from go2_rl_gym_b200 import _abi
                xst = A.tensors["xrew_log"][_abi.NUM_XREW:].view(torch.float32).view(-1, _abi.NUM_XREW)[sp.ep_slot].numpy()
                exr = z[f"out{i}_ep_xrew"]
                on = ~np.isnan(exr)
                assert on.sum() >= (13 if name.startswith("xrew") else 1) and np.allclose(xst[on], exr[on], rtol=1e-4, atol=1e-6), (xst, exr)
    assert n_reset >= 3, "fixture must exercise resets"
    if name == "turn_over":         # all three initial poses were drawn, timers are running, flipped robots are paid by the turn_over scales
        rs, tt = z["out0_root_states"], z["out0_turn_over_timer"]
        assert (np.abs(rs[:, 3]) > 0.9).any() and ((np.abs(rs[:, 3]) > 0.4) & (np.abs(rs[:, 3]) < 0.8)).any() and (np.abs(rs[:, 3]) < 1e-6).any()
        assert set(np.round(tt, 2).tolist()) >= {0.0, 3.0, 5.0} and bool((z["out0_commands"][tt > 0][:, :3] == 0).all())
    if name.startswith("xrew"):     # every stateful extra term really moved: air time accumulated, first contacts rewarded, the termination term fired
        K = int(z["meta_K"])
        assert float(z[f"out{K - 1}_xrew_state"][:, :4].max()) > 0.0 and any(float(np.abs(z[f"out{i}_xrew_sums"][:, 6]).max()) > 0 for i in range(K))
        assert any(float(np.abs(z[f"out{i}_xrew_sums"][:, 3]).max()) > 0 or z[f"out{i}_reset_buf"].sum() > 0 for i in range(K))
    if name == "ctrl_v_pos":  # velocity control + only_positive_rewards (legged_robot.py:612-613,266-267): the clip is really exercised
        assert all(float(z[f"out{i}_rew_buf"].min()) >= 0.0 for i in range(int(z["meta_K"]))) and bool((z["out0_rew_buf"] == 0.0).any())
    if name == "cmdcur":      # the recorded window crosses learning iteration 20 000: the widened command ranges are in force at its end
        assert float(A.tensors["env_command_ranges"][:, 1].max()) == 1.0 and float(A.tensors["env_command_ranges"][:, 5].max()) == 1.5
