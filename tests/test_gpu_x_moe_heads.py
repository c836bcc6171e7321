"""GPU parity of the MoE-actor CTS variants (go2_ac_moe_cts, go2_dual_moe_cts, go2_mcp_cts) against fixtures made by the REFERENCE's rsl_rl
(tests/golden/make_golden_cts.py --variant ...), through the C ABI.

These variants were wired after this round's GPU budget was spent: their host logic is pinned on the CPU (tests/test_emu_rl_cpu.py, same
fixtures, the C ABI replaced by tests/emu_rl.py) and every kernel they launch is covered at its own shapes by tests/test_gpu_rl.py /
tests/test_gpu_cts.py; this file is their first run on hardware, hence its place after the verified files in collection order."""
import os

import numpy as np
import pytest
import torch

from cts_util import STORAGE_KEYS, make_cts

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANTS = [v for v in ("ac_moe_cts", "dual_moe_cts", "mcp_cts") if os.path.exists(os.path.join(G, f"rl_{v}.npz"))]
NEEDS_OBS = ("ac_moe_cts", "dual_moe_cts")
# tc = the DEFAULT tensor-core path (3xTF32 tcgen05 GEMMs, fp32-class products): the same bars as the strict-fp32 CUDA-core path (simt)


def _z(variant):
    return np.load(os.path.join(G, f"rl_{variant}.npz"))


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("gemm", ["simt", "tc"])
def test_variant_act_and_returns_match_reference(gemm, variant, monkeypatch):
    monkeypatch.setenv("GO2_GEMM", gemm)
    Z = _z(variant)
    t = lambda k: torch.from_numpy(Z[k]).cuda()
    model, alg, T, N = make_cts(variant, Z, "cuda")
    tol = tol_a = 2e-5
    a = alg.act(t("in_obs")[0], t("in_priv")[0], t("in_hist")[0])
    st = alg.storage
    assert torch.allclose(st.mu[0], t("st_mu")[0], atol=tol_a)
    assert torch.allclose(st.sigma[0], t("st_sigma")[0], atol=tol_a, rtol=tol_a)
    assert torch.allclose(st.values[0], t("st_values")[0], atol=tol)
    assert torch.equal(st.observations[0], t("st_observations")[0])
    assert torch.equal(a[alg.perm], st.actions[0])
    lp = torch.distributions.Normal(st.mu[0], st.sigma[0]).log_prob(st.actions[0]).sum(-1)
    assert torch.allclose(st.actions_log_prob[0].squeeze(-1), lp, atol=1e-4)
    for k in STORAGE_KEYS:
        getattr(st, k).copy_(t("st_" + k))
    st.step = T
    last = (t("in_obs")[T], t("in_priv")[T], t("in_hist")[T])
    alg.compute_returns(*(last if variant in NEEDS_OBS else last[1:]))
    assert torch.allclose(st.returns, t("st_returns"), atol=10 * tol)
    assert torch.allclose(st.advantages, t("st_advantages"), atol=50 * tol)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("gemm", ["simt", "tc"])
def test_variant_update_matches_reference(gemm, variant, monkeypatch):
    """Both passes of update() (ac_moe_cts.py:144-277, dual_moe_cts.py, mcp_cts.py).  Same bars as tests/test_gpu_cts.py for both GEMM paths: update
    within 2e-3 relative, losses within 2e-4, same learning-rate path."""
    monkeypatch.setenv("GO2_GEMM", gemm)
    Z = _z(variant)
    t = lambda k: torch.from_numpy(Z[k]).cuda()
    model, alg, T, N = make_cts(variant, Z, "cuda")
    for k in STORAGE_KEYS:
        getattr(alg.storage, k).copy_(t("st_" + k))
    losses = alg.update(t("tperm"), t("sperm"))
    assert len(losses) == len(Z["losses"])
    tol = 2e-4
    for a, b in zip(losses, Z["losses"]):
        assert abs(a - b) < tol * max(1.0, abs(b)), (losses, Z["losses"])
    assert abs(alg.learning_rate - float(Z["lr"])) < 1e-9
    num = den = 0.0
    for k, v in model.state_dict().items():
        r, o = torch.from_numpy(Z["sd1_" + k]), torch.from_numpy(Z["sd0_" + k])
        num += float(((v.cpu() - o) - (r - o)).pow(2).sum()); den += float((r - o).pow(2).sum())
    rel = (num / den) ** 0.5
    print(f"[{gemm}] {variant} update: relative error of the update = {rel:.3e}")
    assert rel < 2e-3


@pytest.mark.parametrize("task", ["go2_" + v for v in VARIANTS])
def test_variant_runner_two_iterations(task, tmp_path):
    """train.py's call sequence for the registered task: two logged iterations (graph-replayed rollout), checkpoint, resume, inference policy."""
    from go2_rl_gym_b200.envs import task_registry
    from go2_rl_gym_b200.utils import get_args
    args = get_args(["--task", task, "--num_envs", "256", "--headless"])
    env, _ = task_registry.make_env(task, args)
    runner, _ = task_registry.make_alg_runner(env, task, args, log_root=str(tmp_path))
    runner.learn(2, init_at_random_ep_len=True)
    sd = torch.load(os.path.join(runner.log_dir, "model_2.pt"), weights_only=False)
    assert {"model_state_dict", "optimizer1_state_dict", "optimizer2_state_dict", "iter", "infos"} == set(sd)
    assert all(torch.isfinite(v).all() for v in sd["model_state_dict"].values())
    runner2, _ = task_registry.make_alg_runner(env, task, args, log_root=None)
    runner2.load(os.path.join(runner.log_dir, "model_2.pt"))
    assert runner2.current_learning_iteration == 2
    for (k, a), b in zip(runner.alg.model.state_dict().items(), runner2.alg.model.state_dict().values()):
        assert torch.equal(a, b), k
    a = runner.get_inference_policy()(env.get_observations())
    assert a.shape == (256, 12) and torch.isfinite(a).all()


@pytest.mark.parametrize("task", ["go2_" + v for v in VARIANTS])
def test_variant_play_loop_and_policy_export(task, tmp_path):
    """legged_gym/scripts/play.py for the registered task (see tests/test_gpu_rl.py::test_play_loop_and_policy_export)."""
    import importlib.util
    import math
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("play_script", os.path.join(root, "legged_gym", "scripts", "play.py"))
    play_script = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(play_script)
    from go2_rl_gym_b200.envs import task_registry
    from go2_rl_gym_b200.utils import get_args
    args = get_args(["--task", task, "--num_envs", "64", "--headless"])
    env, _ = task_registry.make_env(task, args)
    runner, _ = task_registry.make_alg_runner(env, task, args, log_root=None)
    stats = play_script.play(get_args(["--task", task, "--num_envs", "49", "--headless"]), num_steps=30, runner=runner, export_dir=str(tmp_path))
    assert stats["steps"] == 30 and all(math.isfinite(v) for v in stats.values())
    out = torch.jit.load(os.path.join(tmp_path, "policy.pt"))(torch.randn(1, 45))
    out = out[0] if isinstance(out, tuple) else out
    assert out.shape == (1, 12) and torch.isfinite(out).all()
