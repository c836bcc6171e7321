"""The RL oracle (oracle/rl_oracle.py) against the reference's own rsl_rl (fixture from tests/golden/make_golden_rl.py)."""
import os

import numpy as np
import torch

from oracle import rl_oracle as R
from golden.rl_cfg import CFG

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rl_ppo.npz"))


def _t(k):
    return torch.from_numpy(Z[k])


def test_forward_and_gae_match_reference():
    sd = {k[4:]: _t(k) for k in Z.files if k.startswith("sd0_")}
    obs, priv = _t("st_observations"), _t("st_privileged_observations")
    mu = R.mlp_forward(sd, "actor", obs)
    assert torch.allclose(mu, _t("st_mu"), atol=1e-6)
    v = R.mlp_forward(sd, "critic", priv)
    assert torch.allclose(v, _t("st_values"), atol=1e-6)
    lp = R.log_prob(mu, sd["std"], _t("st_actions"))
    assert torch.allclose(lp.unsqueeze(-1), _t("st_actions_log_prob"), atol=1e-5)
    last_v = R.mlp_forward(sd, "critic", _t("in_priv")[-1])
    ret, adv = R.gae(_t("st_rewards"), _t("st_values"), _t("st_dones"), last_v, CFG["gamma"], CFG["lam"])
    assert torch.allclose(ret, _t("st_returns"), atol=1e-5)
    assert torch.allclose(adv, _t("st_advantages"), atol=1e-5)
    # time-out bootstrap of process_env_step (ppo.py:109)
    r = _t("in_rew") + CFG["gamma"] * (_t("st_values").squeeze(-1) * _t("in_touts").float())
    assert torch.allclose(r.unsqueeze(-1), _t("st_rewards"), atol=1e-6)


def test_update_matches_reference():
    sd = {k[4:]: _t(k) for k in Z.files if k.startswith("sd0_")}
    f = lambda k: _t("st_" + k).flatten(0, 1)
    data = {"obs": f("observations"), "critic_obs": f("privileged_observations"), "actions": f("actions"), "values": f("values"),
            "returns": f("returns"), "old_logp": f("actions_log_prob"), "adv": f("advantages"), "old_mu": f("mu"), "old_sigma": f("sigma")}
    new, mvl, msl, lr, _ = R.ppo_update(sd, data, _t("perm"), CFG)
    assert abs(mvl - float(Z["mean_value_loss"])) < 1e-6 and abs(msl - float(Z["mean_surrogate_loss"])) < 1e-6
    assert abs(lr - float(Z["lr"])) < 1e-12
    for k, v in new.items():
        assert torch.allclose(v, _t("sd1_" + k), atol=2e-6), k
