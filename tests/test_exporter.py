"""Policy export (go2_rl_gym_b200/utils/exporter.py vs legged_gym/utils/exporter.py:13-192): the TorchScript module built from a
policy's state dict must (a) reproduce a plain fp32 evaluation of the same weights, (b) reproduce the REFERENCE's own exporter on
the reference's own module when /root/reference is present, and (c) reproduce the trained policy the reference ships
(deploy/pre_train/go2/go2_cts_150k.pt) from its weights alone."""
import importlib.util
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from go2_rl_gym_b200.utils import exporter as ex

REF = "/root/reference"
has_ref = os.path.isdir(os.path.join(REF, "rsl_rl"))


def _obs_seq(n=9, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(1, 45, generator=g) for _ in range(n)]


def test_ppo_actor_export_roundtrip(tmp_path):
    from go2_rl_gym_b200.rl.modules import ActorCritic
    torch.manual_seed(0)
    ac = ActorCritic(45, 263, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128])
    path = ex.export_policy_as_jit(ac, str(tmp_path))
    m = torch.jit.load(path)
    sd = ac.state_dict()
    for x in _obs_seq(3):
        h = x
        for i in (0, 2, 4, 6):
            h = F.linear(h, sd[f"actor.{i}.weight"], sd[f"actor.{i}.bias"])
            if i < 6:
                h = F.elu(h)
        assert torch.allclose(m(x), h, atol=1e-6)
    ex.export_policy_as_pkl(ac, str(tmp_path))
    assert set(torch.load(os.path.join(tmp_path, "policy.pkl"))) == set(sd)


def test_moe_cts_export_matches_manual_forward(tmp_path):
    from go2_rl_gym_b200.rl.modules import ActorCriticMoECTS
    torch.manual_seed(1)
    pol = ActorCriticMoECTS(45, 263, 12, 4, 5)
    m = torch.jit.load(ex.export_policy_as_jit(pol, str(tmp_path)))
    sd = pol.state_dict()
    P = "student_moe_encoder.moe."
    hist = torch.zeros(1, 5, 45)

    def mlp(prefix, idx, x, last_act=False):
        for n, i in enumerate(idx):
            x = F.linear(x, sd[f"{prefix}.{i}.weight"], sd[f"{prefix}.{i}.bias"])
            if n < len(idx) - 1 or last_act:
                x = F.elu(x)
        return x
    for x in _obs_seq(8):
        hist = torch.cat([hist[:, 1:], x[:, None]], 1)
        h = hist.flatten(1)
        w = torch.softmax(mlp(P + "gating_network.0.network", (0, 2, 4), h), -1)
        feat = mlp(P + "experts.backbone.network", (0, 2, 4), h, last_act=True)
        outs = F.conv1d(feat[:, :, None], sd[P + "experts.experts.weight"], sd[P + "experts.experts.bias"], groups=8)[:, :, 0].reshape(1, 8, 32)
        lat = F.normalize((w[:, :, None] * outs).sum(1), dim=-1)
        a = mlp("actor.network", (0, 2, 4, 6), torch.cat([lat, x], 1))
        act, (weights, latent) = m(x)
        assert torch.allclose(act, a, atol=1e-5) and torch.allclose(weights, w, atol=1e-6) and torch.allclose(latent, lat, atol=1e-6)
    m.reset()
    a0, _ = m(_obs_seq(1, 5)[0])
    m2 = torch.jit.load(os.path.join(tmp_path, "policy.pt"))
    b0, _ = m2(_obs_seq(1, 5)[0])
    assert torch.equal(a0, b0)          # reset() restores the zero history of a freshly loaded module


@pytest.mark.skipif(not has_ref, reason="needs the reference tree (runs in the build container)")
def test_moe_cts_export_matches_reference_exporter(tmp_path):
    """Reference ActorCriticMoECTS -> reference exporter vs the same weights in this package's module -> this exporter."""
    sys.path.insert(0, os.path.join(REF, "rsl_rl"))
    try:
        from rsl_rl.modules.actor_critic_moe_cts import ActorCriticMoECTS as RefMoE
    finally:
        sys.path.pop(0)
    spec = importlib.util.spec_from_file_location("ref_exporter", os.path.join(REF, "legged_gym", "utils", "exporter.py"))
    ref_ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_ex)
    from go2_rl_gym_b200.rl.modules import ActorCriticMoECTS
    torch.manual_seed(3)
    ref = RefMoE(45, 263, 12, 4, 5)
    mine = ActorCriticMoECTS(45, 263, 12, 4, 5)
    mine.load_state_dict(ref.state_dict())
    ref_ex.export_policy_as_jit(ref, str(tmp_path / "ref"))
    ex.export_policy_as_jit(mine, str(tmp_path / "mine"))
    mr, mm = torch.jit.load(str(tmp_path / "ref" / "policy.pt")), torch.jit.load(str(tmp_path / "mine" / "policy.pt"))
    for x in _obs_seq(9, 2):
        (ar, (wr, lr)), (am, (wm, lm)) = mr(x), mm(x)
        assert torch.allclose(ar, am, atol=1e-5) and torch.allclose(wr, wm, atol=1e-6) and torch.allclose(lr, lm, atol=1e-6)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "deploy/pre_train/go2/go2_cts_150k.pt")), reason="needs the reference's shipped policy")
def test_cts_export_reproduces_shipped_policy(tmp_path):
    """Known-answer test: the weights of the policy the reference ships, re-exported by this package, act like the original file."""
    shipped = torch.jit.load(os.path.join(REF, "deploy/pre_train/go2/go2_cts_150k.pt"), map_location="cpu")

    class Holder(torch.nn.Module):          # a state-dict carrier with the ActorCriticCTS key layout
        is_recurrent = False

        def __init__(self, sd):
            super().__init__()
            self._sd = sd
            self.history = torch.zeros(1, 5, 45)

        def state_dict(self, *a, **k):
            return self._sd
    m = torch.jit.load(ex.export_policy_as_jit(Holder({k: v.clone() for k, v in shipped.state_dict().items()}), str(tmp_path)))
    for x in _obs_seq(12, 4):
        ref = shipped(x)
        ref = ref[0] if isinstance(ref, tuple) else ref
        act, (none, latent) = m(x)
        assert none is None and latent.shape == (1, 32) and abs(float(latent.norm()) - 1.0) < 1e-5
        assert torch.allclose(act, ref, atol=1e-5)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "deploy/pre_train/go2/go2_moe_cts_137k_0.6739.pt")), reason="needs the reference's shipped policy")
def test_moe_ng_cts_export_reproduces_shipped_policy(tmp_path):
    """Known-answer test: the trained MoE policy the reference ships has the no-goal variant's parameter layout (experts_backbone /
    experts_hidden / experts_out / gating_network) with experts that still read the full 225-wide history (its scripted forward ignores the
    masked input).  Its weights loaded into THIS package's ActorCriticMoENGCTS (mask = keep every column) and re-exported reproduce the
    original file's actions, gate weights and latents."""
    shipped = torch.jit.load(os.path.join(REF, "deploy/pre_train/go2/go2_moe_cts_137k_0.6739.pt"), map_location="cpu")
    from go2_rl_gym_b200.rl.modules import ActorCriticMoENGCTS
    assert shipped.state_dict()["student_moe_encoder.experts_backbone.0.weight"].shape[1] == 45 * int(shipped.history_length)
    mask = [True] * 45
    pol = ActorCriticMoENGCTS(45, 263, 12, 1, int(shipped.history_length), mask)
    sd = pol.state_dict()
    sd.update({k: v.clone() for k, v in shipped.state_dict().items()})          # student encoder + actor; teacher / critic stay random (unused)
    pol.load_state_dict(sd)
    m = torch.jit.load(ex.export_policy_as_jit(pol, str(tmp_path)))
    for x in _obs_seq(12, 7):
        act_r, (w_r, lat_r) = shipped(x)
        act, (w, lat) = m(x)
        assert torch.allclose(act, act_r, atol=1e-5) and torch.allclose(w, w_r, atol=1e-6) and torch.allclose(lat, lat_r, atol=1e-6)


@pytest.mark.skipif(not has_ref, reason="needs the reference tree (runs in the build container)")
def test_moe_ng_cts_export_matches_reference_exporter(tmp_path):
    """Reference ActorCriticMoENGCTS (GO2 no-goal mask) -> reference exporter vs the same weights in this package's module -> this exporter."""
    sys.path.insert(0, os.path.join(REF, "rsl_rl"))
    try:
        from rsl_rl.modules.actor_critic_moe_ng_cts import ActorCriticMoENGCTS as RefNG
    finally:
        sys.path.pop(0)
    spec = importlib.util.spec_from_file_location("ref_exporter", os.path.join(REF, "legged_gym", "utils", "exporter.py"))
    ref_ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_ex)
    from go2_rl_gym_b200.rl.modules import ActorCriticMoENGCTS
    mask = [True] * 6 + [False] * 3 + [True] * 36
    torch.manual_seed(5)
    ref = RefNG(45, 263, 12, 4, 5, mask)
    mine = ActorCriticMoENGCTS(45, 263, 12, 4, 5, mask)
    mine.load_state_dict(ref.state_dict())
    ref_ex.export_policy_as_jit(ref, str(tmp_path / "ref"))
    ex.export_policy_as_jit(mine, str(tmp_path / "mine"))
    mr, mm = torch.jit.load(str(tmp_path / "ref" / "policy.pt")), torch.jit.load(str(tmp_path / "mine" / "policy.pt"))
    for x in _obs_seq(9, 3):
        (ar, (wr, lr)), (am, (wm, lm)) = mr(x), mm(x)
        assert torch.allclose(ar, am, atol=1e-5) and torch.allclose(wr, wm, atol=1e-6) and torch.allclose(lr, lm, atol=1e-6)


@pytest.mark.skipif(not has_ref, reason="needs the reference tree (runs in the build container)")
@pytest.mark.parametrize("variant", ["mcp_cts", "ac_moe_cts", "dual_moe_cts"])
def test_ablation_variants_export_matches_reference_exporter(variant, tmp_path):
    """The three CTS ablation variants: the reference's module -> the reference's exporter vs THIS exporter on the same module (it only reads
    the state dict, the history shape and the no-goal mask)."""
    import contextlib, io
    sys.path.insert(0, os.path.join(REF, "rsl_rl"))
    try:
        from rsl_rl.modules.actor_critic_mcp_cts import ActorCriticMCPCTS
        from rsl_rl.modules.actor_critic_ac_moe_cts import ActorCriticACMoECTS
        from rsl_rl.modules.actor_critic_dual_moe_cts import ActorCriticDualMoECTS
    finally:
        sys.path.pop(0)
    spec = importlib.util.spec_from_file_location("ref_exporter", os.path.join(REF, "legged_gym", "utils", "exporter.py"))
    ref_ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_ex)
    torch.manual_seed(11)
    with contextlib.redirect_stdout(io.StringIO()):
        if variant == "mcp_cts":
            ref = ActorCriticMCPCTS(45, 263, 12, 4, 5, obs_no_goal_mask=[True] * 6 + [False] * 3 + [True] * 36)
        elif variant == "ac_moe_cts":
            ref = ActorCriticACMoECTS(45, 263, 12, 4, 5)
        else:
            ref = ActorCriticDualMoECTS(45, 263, 12, 4, 5)
    ref_ex.export_policy_as_jit(ref, str(tmp_path / "ref"))
    ex.export_policy_as_jit(ref, str(tmp_path / "mine"))
    mr, mm = torch.jit.load(str(tmp_path / "ref" / "policy.pt")), torch.jit.load(str(tmp_path / "mine" / "policy.pt"))
    for x in _obs_seq(9, 6):
        (ar, extra_r), (am, extra_m) = mr(x), mm(x)
        assert torch.allclose(ar, am, atol=1e-5), variant
        assert len(extra_r) == len(extra_m)
        for a, b in zip(extra_r, extra_m):
            assert torch.allclose(a, b, atol=1e-5), variant
