"""Shared set-up of the CTS-family parity tests (CPU: tests/test_emu_rl_cpu.py; GPU: tests/test_gpu_x_moe_heads.py)."""
import torch

STORAGE_KEYS = ("observations", "privileged_observations", "history", "actions", "rewards", "dones", "values", "returns", "advantages",
                "actions_log_prob", "mu", "sigma")


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


REF_RSL = "/root/reference/rsl_rl"


def side_by_side(variant, policy, alg_kw, N, T, seed, monkeypatch, H=5, iters=1):
    """Run the REFERENCE's module + algorithm (imported from /root/reference/rsl_rl; build container only) and this package's over the emulated C
    ABI on the same random data: T x (act, process_env_step), compute_returns, update with shared permutations.  Returns the worst deviations:
    act (mu / sigma / value), returns, adv, loss (relative), lr, update_rel (relative error of the parameter update)."""
    import sys
    monkeypatch.syspath_prepend(REF_RSL)
    for k in [k for k in sys.modules if k == "rsl_rl" or k.startswith("rsl_rl.")]:
        monkeypatch.delitem(sys.modules, k)
    import contextlib
    import io
    import rsl_rl.storage.rollout_storage_cts as RS
    from golden.cts_cfg import NO_GOAL_MASK
    from go2_rl_gym_b200.rl import algorithms as A, modules as Mo
    import importlib
    names = {"cts": ("cts", "CTS", "actor_critic_cts", "ActorCriticCTS"), "moe_cts": ("moe_cts", "MoECTS", "actor_critic_moe_cts", "ActorCriticMoECTS"),
             "moe_ng_cts": ("moe_ng_cts", "MoENGCTS", "actor_critic_moe_ng_cts", "ActorCriticMoENGCTS"),
             "ac_moe_cts": ("ac_moe_cts", "ACMoECTS", "actor_critic_ac_moe_cts", "ActorCriticACMoECTS"),
             "dual_moe_cts": ("dual_moe_cts", "DualMoECTS", "actor_critic_dual_moe_cts", "ActorCriticDualMoECTS"),
             "mcp_cts": ("mcp_cts", "MCPCTS", "actor_critic_mcp_cts", "ActorCriticMCPCTS")}[variant]
    RefAlg = getattr(importlib.import_module("rsl_rl.algorithms." + names[0]), names[1])
    ref_mod = importlib.import_module("rsl_rl.modules." + names[2])
    RefModel = getattr(ref_mod, names[3])
    ours_m, ours_a = getattr(Mo, names[3]), getattr(A, names[1])
    needs_obs = variant in ("ac_moe_cts", "dual_moe_cts")
    policy = dict(policy)
    if variant in ("mcp_cts", "moe_ng_cts"):
        policy.setdefault("obs_no_goal_mask", NO_GOAL_MASK)
    if variant == "mcp_cts":
        policy.setdefault("actor_hidden_dims", [512, 256, 128])       # GO2CfgMCPCTS
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        if variant == "cts":      # the reference allocates this module's history on 'cuda' (actor_critic_cts.py:48)
            zeros = torch.zeros
            ref_mod.torch.zeros = lambda *a, **kw: zeros(*a, **{k: v for k, v in kw.items() if k != "device"})
            try:
                ref = RefModel(45, 263, 12, N, H, **policy)
            finally:
                ref_mod.torch.zeros = zeros
        else:
            ref = RefModel(45, 263, 12, N, H, **policy)
        model = ours_m(45, 263, 12, N, H, **policy)
    assert [k for k, _ in model.named_parameters()] == [k for k, _ in ref.named_parameters()]
    model.load_state_dict(ref.state_dict())
    ralg, alg = RefAlg(ref, N, H, device="cpu", **alg_kw), ours_a(model, N, H, device="cpu", **alg_kw)
    ralg.init_storage(N, T, [45], [263], [12]); alg.init_storage(N, T, [45], [263], [12])
    g = torch.Generator().manual_seed(seed + 1)
    out = {"act": 0.0, "returns": 0.0, "adv": 0.0, "loss": 0.0}
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    for it in range(iters):          # consecutive iterations on the SAME instances: Adam moments / step counts, the learning rate and the storage carry over
        _one_iteration(ralg, alg, RS, needs_obs, N, T, H, g, out, monkeypatch)
    out["lr"] = abs(alg.learning_rate - ralg.learning_rate)
    num = den = 0.0
    for (k, v), r in zip(model.state_dict().items(), ref.state_dict().values()):
        num += float(((v - sd0[k]) - (r - sd0[k])).pow(2).sum()); den += float((r - sd0[k]).pow(2).sum())
    out["update_rel"] = (num / max(den, 1e-30)) ** 0.5
    return out


def _one_iteration(ralg, alg, RS, needs_obs, N, T, H, g, out, monkeypatch):
    obs, priv, hist = torch.randn(T + 1, N, 45, generator=g), torch.randn(T + 1, N, 263, generator=g), torch.randn(T + 1, N, H * 45, generator=g)
    rew, dones = 0.1 * torch.randn(T, N, generator=g), torch.rand(T, N, generator=g) < 0.05
    with torch.inference_mode():
        for t in range(T):
            ralg.act(obs[t], priv[t], hist[t])
            alg.act(obs[t], priv[t], hist[t])
            for a, b in ((alg.storage.mu[t], ralg.transition.action_mean), (alg.storage.sigma[t], ralg.transition.action_sigma),
                         (alg.storage.values[t], ralg.transition.values)):
                out["act"] = max(out["act"], float((a - b).abs().max()))
            for k in ("actions", "actions_log_prob"):       # same actions on both sides from here on
                getattr(alg.storage, k)[t].copy_(getattr(ralg.transition, k).view_as(getattr(alg.storage, k)[t]))
            ralg.process_env_step(rew[t], dones[t], {"time_outs": dones[t]})
            alg.process_env_step(rew[t], dones[t], {"time_outs": dones[t]})
        last = (obs[T], priv[T], hist[T])
        ralg.compute_returns(*(last if needs_obs else last[1:]))
        alg.compute_returns(*(last if needs_obs else last[1:]))
    out["returns"] = max(out["returns"], float((alg.storage.returns - ralg.storage.returns).abs().max()))
    out["adv"] = max(out["adv"], float((alg.storage.advantages - ralg.storage.advantages).abs().max()))
    nt, ns = alg.teacher_num_envs * T, alg.student_num_envs * T
    tperm, sperm = torch.randperm(nt, generator=g), torch.randperm(ns, generator=g)
    queue = [tperm.clone(), sperm.clone()]
    with monkeypatch.context() as mp:
        mp.setattr(RS.torch, "randperm", lambda n, **kw: queue.pop(0))
        rl = ralg.update()
    ol = alg.update(tperm, sperm)
    assert len(ol) == len(rl)
    out["loss"] = max(out["loss"], max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(ol, rl)))


def side_by_side_ppo(policy, alg_kw, N, T, seed, monkeypatch):
    """side_by_side for the plain PPO task: the reference's ActorCritic + PPO + RolloutStorage against this package's over the emulated C ABI."""
    import contextlib
    import importlib
    import io
    import sys
    monkeypatch.syspath_prepend(REF_RSL)
    for k in [k for k in sys.modules if k == "rsl_rl" or k.startswith("rsl_rl.")]:
        monkeypatch.delitem(sys.modules, k)
    RS = importlib.import_module("rsl_rl.storage.rollout_storage")
    RefPPO = importlib.import_module("rsl_rl.algorithms.ppo").PPO
    RefAC = importlib.import_module("rsl_rl.modules.actor_critic").ActorCritic
    from go2_rl_gym_b200.rl.algorithms import PPO
    from go2_rl_gym_b200.rl.modules import ActorCritic
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefAC(45, 263, 12, **policy)
        ac = ActorCritic(45, 263, 12, **policy)
    ac.load_state_dict(ref.state_dict())
    ralg, alg = RefPPO(ref, device="cpu", **alg_kw), PPO(ac, device="cpu", **alg_kw)
    ralg.init_storage(N, T, [45], [263], [12]); alg.init_storage(N, T, [45], [263], [12])
    g = torch.Generator().manual_seed(seed + 1)
    obs, priv = torch.randn(T + 1, N, 45, generator=g), torch.randn(T + 1, N, 263, generator=g)
    rew, dones = 0.1 * torch.randn(T, N, generator=g), torch.rand(T, N, generator=g) < 0.05
    out = {"act": 0.0}
    with torch.inference_mode():
        for t in range(T):
            ralg.act(obs[t], priv[t])
            alg.act(obs[t], priv[t])
            for a, b in ((alg.storage.mu[t], ralg.transition.action_mean), (alg.storage.sigma[t], ralg.transition.action_sigma),
                         (alg.storage.values[t], ralg.transition.values)):
                out["act"] = max(out["act"], float((a - b).abs().max()))
            for k in ("actions", "actions_log_prob"):
                getattr(alg.storage, k)[t].copy_(getattr(ralg.transition, k).view_as(getattr(alg.storage, k)[t]))
            ralg.process_env_step(rew[t], dones[t], {"time_outs": dones[t]})
            alg.process_env_step(rew[t], dones[t], {"time_outs": dones[t]})
        ralg.compute_returns(priv[T])
        alg.compute_returns(priv[T])
    out["returns"] = float((alg.storage.returns - ralg.storage.returns).abs().max())
    out["adv"] = float((alg.storage.advantages - ralg.storage.advantages).abs().max())
    nb = alg_kw["num_mini_batches"]
    perm = torch.randperm(nb * (N * T // nb), generator=g)
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    with monkeypatch.context() as mp:
        mp.setattr(RS.torch, "randperm", lambda n, **kw: perm.clone())
        rl = ralg.update()
    ol = alg.update(indices=perm)
    out["loss"] = max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(ol, rl))
    out["lr"] = abs(alg.learning_rate - ralg.learning_rate)
    num = den = 0.0
    for (k, v), r in zip(ac.state_dict().items(), ref.state_dict().values()):
        num += float(((v - sd0[k]) - (r - sd0[k])).pow(2).sum()); den += float((r - sd0[k]).pow(2).sum())
    out["update_rel"] = (num / max(den, 1e-30)) ** 0.5
    return out
