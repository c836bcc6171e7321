"""Host logic of the trainer (PPO / CTS / MoE-CTS algorithm classes, MlpEngine, flat parameter storage) on CPU tensors, with the C ABI's
trainer entry points replaced by tests/emu_rl.py, against the fixtures made by the REFERENCE's rsl_rl (tests/golden/rl_*.npz).

The same comparisons run against the real kernels in tests/test_gpu_rl.py / tests/test_gpu_cts.py (-m gpu); here both GEMM wirings
("simt" and "tc": padded operands, W^T copies, ones columns, narrow-head kernels) are driven with exact fp32 arithmetic, so the bars are
the strict-fp32 ones for both."""
import os

import numpy as np
import pytest
import torch

import emu_rl

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
CTS_VARIANTS = ["moe_cts", "cts", "moe_ng_cts", "ac_moe_cts", "dual_moe_cts", "mcp_cts"]
STORAGE_KEYS = ("observations", "privileged_observations", "history", "actions", "rewards", "dones", "values", "returns", "advantages",
                "actions_log_prob", "mu", "sigma")


def test_emulated_philox_known_answers():
    for ctr, key, exp in KAT:
        assert tuple(int(x) for x in emu_rl.philox(*ctr, *key)) == exp


def test_product_refuses_to_run_without_the_library(monkeypatch):
    """No silent fallback: on a host without CUDA the trainer's first kernel call must raise."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from go2_rl_gym_b200.rl import _ops
    monkeypatch.setattr(_ops, "_lib", None)
    with pytest.raises(Exception):
        _ops.call("go2_colsum", 0, 0, 0, 0, 0, 0)


def _load(name):
    return np.load(os.path.join(G, f"rl_{name}.npz"))


def _check_update(model, Z, strict=True):
    num = den = 0.0
    for k, v in model.state_dict().items():
        r, o = torch.from_numpy(Z["sd1_" + k]), torch.from_numpy(Z["sd0_" + k])
        num += float(((v - o) - (r - o)).pow(2).sum()); den += float((r - o).pow(2).sum())
        if strict:      # Adam divides by sqrt(v): a handful of near-zero-gradient elements amplify summation-order differences
            e = (v - r).abs()
            assert float((e > 3e-5 + 1e-3 * r.abs()).float().mean()) <= 1e-3 and float(e.max()) < 5e-4, (k, float(e.max()))
    return (num / den) ** 0.5


@pytest.mark.parametrize("gemm", ["simt", "tc"])
def test_ppo_update_host_logic_matches_reference_fixture(gemm, monkeypatch):
    emu_rl.install(monkeypatch)
    monkeypatch.setenv("GO2_GEMM", gemm)
    Z = _load("ppo")
    from golden.rl_cfg import CFG
    from go2_rl_gym_b200.rl.algorithms import PPO
    from go2_rl_gym_b200.rl.modules import ActorCritic
    T, N = Z["st_rewards"].shape[:2]
    ac = ActorCritic(45, 263, 12, actor_hidden_dims=[64, 32, 16], critic_hidden_dims=[64, 32, 16])
    ac.load_state_dict({k[4:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("sd0_")})
    alg = PPO(ac, device="cpu", **CFG)
    alg.init_storage(N, T, [45], [263], [12])
    st = alg.storage
    for k in STORAGE_KEYS:
        if k != "history":
            getattr(st, k).copy_(torch.from_numpy(Z["st_" + k]))
    mvl, msl = alg.update(indices=torch.from_numpy(Z["perm"]))
    assert abs(mvl - float(Z["mean_value_loss"])) < 1e-4 and abs(msl - float(Z["mean_surrogate_loss"])) < 1e-4
    assert abs(alg.learning_rate - float(Z["lr"])) < 1e-9
    assert _check_update(ac, Z) < 1e-3


def test_gae_host_logic_matches_reference_fixture(monkeypatch):
    emu_rl.install(monkeypatch)
    Z = _load("ppo")
    from go2_rl_gym_b200.rl.storage import RolloutStorage
    from oracle import rl_oracle as R
    T, N = Z["st_rewards"].shape[:2]
    st = RolloutStorage(N, T, [45], [263], [12], device="cpu")
    for k in ("rewards", "values", "dones"):
        getattr(st, k).copy_(torch.from_numpy(Z["st_" + k]))
    sd = {k[4:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("sd0_")}
    st.compute_returns(R.mlp_forward(sd, "critic", torch.from_numpy(Z["in_priv"][-1])), 0.99, 0.95)
    assert torch.allclose(st.returns, torch.from_numpy(Z["st_returns"]), atol=1e-5)
    assert torch.allclose(st.advantages, torch.from_numpy(Z["st_advantages"]), atol=2e-5)


def make_cts(variant, Z, device):
    """(model, algorithm, T, N) of a CTS-family variant with the fixture's initial weights (shared with tests/test_gpu_cts.py)."""
    from golden import cts_cfg as cc
    from go2_rl_gym_b200.rl import algorithms as A, modules as Mo
    T, N = Z["st_rewards"].shape[:2]
    model_cls, alg_cls, policy, alg_kw = {
        "moe_cts": (Mo.ActorCriticMoECTS, A.MoECTS, cc.POLICY, cc.ALG),
        "moe_ng_cts": (Mo.ActorCriticMoENGCTS, A.MoENGCTS, cc.POLICY_NG, cc.ALG),
        "cts": (Mo.ActorCriticCTS, A.CTS, cc.POLICY_CTS, cc.ALG_CTS),
        "ac_moe_cts": (getattr(Mo, "ActorCriticACMoECTS", None), getattr(A, "ACMoECTS", None), getattr(cc, "POLICY_AC", None), cc.ALG),
        "dual_moe_cts": (getattr(Mo, "ActorCriticDualMoECTS", None), getattr(A, "DualMoECTS", None), getattr(cc, "POLICY_DUAL", None), cc.ALG),
        "mcp_cts": (getattr(Mo, "ActorCriticMCPCTS", None), getattr(A, "MCPCTS", None), getattr(cc, "POLICY_MCP", None), cc.ALG_CTS),
    }[variant]
    model = model_cls(45, 263, 12, N, 5, **policy)
    model.load_state_dict({k[4:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("sd0_")})
    alg = alg_cls(model, N, 5, device=device, **alg_kw)
    alg.init_storage(N, T, [45], [263], [12])
    return model, alg, T, N


def _variant_or_skip(variant):
    if not os.path.exists(os.path.join(G, f"rl_{variant}.npz")):
        pytest.skip(f"no fixture for {variant}")
    return _load(variant)


@pytest.mark.parametrize("variant", CTS_VARIANTS)
@pytest.mark.parametrize("gemm", ["simt", "tc"])
def test_cts_act_host_logic_matches_reference(gemm, variant, monkeypatch):
    Z = _variant_or_skip(variant)
    emu_rl.install(monkeypatch)
    monkeypatch.setenv("GO2_GEMM", gemm)
    model, alg, T, N = make_cts(variant, Z, "cpu")
    t = lambda k: torch.from_numpy(Z[k])
    a = alg.act(t("in_obs")[0], t("in_priv")[0], t("in_hist")[0])
    st = alg.storage
    assert torch.allclose(st.mu[0], t("st_mu")[0], atol=2e-5)
    assert torch.allclose(st.sigma[0], t("st_sigma")[0], atol=2e-5)
    assert torch.allclose(st.values[0], t("st_values")[0], atol=2e-5)
    assert torch.equal(st.observations[0], t("st_observations")[0])
    assert torch.equal(st.history[0], t("st_history")[0])
    assert torch.equal(a[alg.perm], st.actions[0])
    # the stored log-prob is the log-density of the stored action under the stored (mu, sigma)
    lp = torch.distributions.Normal(st.mu[0], st.sigma[0]).log_prob(st.actions[0]).sum(-1)
    assert torch.allclose(st.actions_log_prob[0].squeeze(-1), lp, atol=1e-4)
    alg.process_env_step(t("in_rew")[0], t("in_dones")[0], {"time_outs": t("in_touts")[0]})
    ti, si = alg.teacher_env_idxs, alg.student_env_idxs
    assert torch.equal(st.dones[0].squeeze(-1).bool(), torch.cat([t("in_dones")[0][ti], t("in_dones")[0][si]]))


@pytest.mark.parametrize("variant", CTS_VARIANTS)
@pytest.mark.parametrize("gemm", ["simt", "tc"])
def test_cts_returns_and_update_host_logic_match_reference(gemm, variant, monkeypatch):
    """compute_returns on the stored rollout, then both passes of update() with the fixture's permutations."""
    Z = _variant_or_skip(variant)
    emu_rl.install(monkeypatch)
    monkeypatch.setenv("GO2_GEMM", gemm)
    model, alg, T, N = make_cts(variant, Z, "cpu")
    st = alg.storage
    t = lambda k: torch.from_numpy(Z[k])
    for k in STORAGE_KEYS:
        getattr(st, k).copy_(t("st_" + k))
    st.returns.zero_(); st.advantages.zero_()
    st.step = T
    last = (t("in_obs")[T], t("in_priv")[T], t("in_hist")[T])
    alg.compute_returns(*(last if variant in ("ac_moe_cts", "dual_moe_cts") else last[1:]))
    assert torch.allclose(st.returns, t("st_returns"), atol=2e-5)
    assert torch.allclose(st.advantages, t("st_advantages"), atol=1e-4)
    for k in ("returns", "advantages"):
        getattr(st, k).copy_(t("st_" + k))
    losses = alg.update(t("tperm"), t("sperm"))
    assert len(losses) == len(Z["losses"])
    for a, b in zip(losses, Z["losses"]):
        assert abs(a - b) < 2e-4 * max(1.0, abs(b)), (losses, Z["losses"])
    assert abs(alg.learning_rate - float(Z["lr"])) < 1e-9
    assert _check_update(model, Z) < 2e-3
