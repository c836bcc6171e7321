"""The multi-GPU trainer path on CPU: two gloo ranks, each with HALF of a reference-made fixture's envs, run the real
compute_returns() + update() of the algorithm classes (C ABI emulated by tests/emu_rl.py) — the per-optimiser-step all-reduce of the flat
gradient + scalar tail, the all-reduced advantage statistics, the 1 / world_size scaling of the student pass, the mean gate usage of the
load-balance terms summed over the ranks — and must end (a) bit-identical on
both ranks and (b) equal, to summation-order tolerance, to ONE process that runs the whole batch with the mini-batches composed of the same samples."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import emu_rl
from cts_util import STORAGE_KEYS

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NB = 4          # mini-batches (golden/cts_cfg.py, golden/rl_cfg.py)


def _build(variant, Z, device, N, env_offset, lb0):
    """(model-or-actor_critic, algorithm) for `variant` over N envs starting at global env id env_offset."""
    import contextlib
    import io
    T = Z["st_rewards"].shape[0]
    with contextlib.redirect_stdout(io.StringIO()):
        if variant == "ppo":
            from golden.rl_cfg import CFG
            from go2_rl_gym_b200.rl.algorithms import PPO
            from go2_rl_gym_b200.rl.modules import ActorCritic
            model = ActorCritic(45, 263, 12, actor_hidden_dims=[64, 32, 16], critic_hidden_dims=[64, 32, 16])
            model.load_state_dict({k[4:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("sd0_")})
            alg = PPO(model, device=device, env_offset=env_offset, **CFG)
            alg.init_storage(N, T, [45], [263], [12])
        else:
            from golden import cts_cfg as cc
            from go2_rl_gym_b200.rl import algorithms as A, modules as Mo
            mcls, acls, pol, akw = {"cts": (Mo.ActorCriticCTS, A.CTS, cc.POLICY_CTS, cc.ALG_CTS), "moe_cts": (Mo.ActorCriticMoECTS, A.MoECTS, cc.POLICY, cc.ALG),
                                    "ac_moe_cts": (Mo.ActorCriticACMoECTS, A.ACMoECTS, cc.POLICY_AC, cc.ALG),
                                    "mcp_cts": (Mo.ActorCriticMCPCTS, A.MCPCTS, cc.POLICY_MCP, cc.ALG_CTS)}[variant]
            akw = dict(akw)
            if lb0 and "load_balance_coef" in akw:
                akw["load_balance_coef"] = 0.0       # (unused since round 2: the mean gate usage is summed over the ranks, dist_utils.SmallSum)
            model = mcls(45, 263, 12, N, 5, **pol)
            model.load_state_dict({k[4:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("sd0_")})
            alg = acls(model, N, 5, device=device, env_offset=env_offset, **akw)
            alg.init_storage(N, T, [45], [263], [12])
    return model, alg


def _env_major(variant, Z, NG):
    """Fixture storage arrays back in ENV order [T, NG, .] (the CTS fixtures are stored teacher-first)."""
    out = {}
    if variant == "ppo":
        perm = np.arange(NG)
    else:
        ids = np.arange(NG)
        perm = np.concatenate([ids[ids % 4 != 0], ids[ids % 4 == 0]])
    for k in STORAGE_KEYS:
        if "st_" + k in Z.files:
            a = np.empty_like(Z["st_" + k])
            a[:, perm] = Z["st_" + k]
            out[k] = a
    return out


def _fill(alg, data, sl, variant):
    st = alg.storage
    order = np.arange(sl.stop - sl.start) if variant == "ppo" else alg.perm.numpy()
    for k, a in data.items():
        getattr(st, k).copy_(torch.from_numpy(a[:, sl][:, order]))
    st.step = st.num_transitions_per_env


def _local_perms(variant, alg, rank, seed):
    g = torch.Generator().manual_seed(seed + rank)
    st = alg.storage
    T = st.num_transitions_per_env
    if variant == "ppo":
        return (torch.randperm(st.num_envs * T, generator=g),)
    return torch.randperm(alg.teacher_num_envs * T, generator=g), torch.randperm(alg.student_num_envs * T, generator=g)


def _last_args(variant, Z, sl):
    T = Z["st_rewards"].shape[0]
    t = lambda k: torch.from_numpy(Z[k][T][sl])
    if variant == "ppo":
        return (t("in_priv"),)
    return (t("in_obs"), t("in_priv"), t("in_hist")) if variant == "ac_moe_cts" else (t("in_priv"), t("in_hist"))


def _worker(rank, world, port, variant, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), GO2_GEMM="tc")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu_rl.install_process()
    Z = np.load(os.path.join(G, f"rl_{variant}.npz"))
    NG = Z["st_rewards"].shape[1]
    Nl = NG // world
    sl = slice(rank * Nl, (rank + 1) * Nl)
    model, alg = _build(variant, Z, "cpu", Nl, rank * Nl, lb0=False)
    assert alg.world_size == world
    _fill(alg, _env_major(variant, Z, NG), sl, variant)
    with torch.inference_mode():
        alg.compute_returns(*_last_args(variant, Z, sl))
    adv = alg.storage.advantages.clone()
    perms = _local_perms(variant, alg, rank, 11)
    losses = alg.update(perms[0]) if variant == "ppo" else alg.update(*perms)
    q.put((rank, {k: v.clone().numpy() for k, v in model.state_dict().items()}, np.array(losses), alg.learning_rate, adv.numpy(),
           None if variant == "ppo" else alg.perm.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("variant", ["ppo", "cts", "moe_cts", "ac_moe_cts", "mcp_cts"])
def test_two_rank_update_equals_the_single_process_update(variant, monkeypatch):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 21000 + (os.getpid() * 7 + len(variant) * 131) % 8000
    procs = [ctx.Process(target=_worker, args=(r, world, port, variant, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # (a) the ranks agree bit for bit: same all-reduced gradient, same learning-rate path
    for k in res[0][1]:
        assert np.array_equal(res[0][1][k], res[1][1][k]), k
    assert np.array_equal(res[0][2], res[1][2]) and res[0][3] == res[1][3]

    # (b) one process over all envs, mini-batch i = union of the ranks' mini-batches i
    emu_rl.install(monkeypatch)
    monkeypatch.setenv("GO2_GEMM", "tc")
    Z = np.load(os.path.join(G, f"rl_{variant}.npz"))
    T, NG = Z["st_rewards"].shape[:2]
    Nl = NG // world
    model, alg = _build(variant, Z, "cpu", NG, 0, lb0=False)
    _fill(alg, _env_major(variant, Z, NG), slice(0, NG), variant)
    with torch.inference_mode():
        alg.compute_returns(*_last_args(variant, Z, slice(0, NG)))
    # all-reduced advantage statistics == global normalisation
    adv_env = np.empty((T, NG, 1), dtype=np.float32)
    for r in range(world):
        order = np.arange(Nl) if variant == "ppo" else res[r][5]
        adv_env[:, r * Nl + order] = res[r][4]
    g_order = np.arange(NG) if variant == "ppo" else alg.perm.numpy()
    assert np.allclose(adv_env[:, g_order], alg.storage.advantages.numpy(), atol=2e-6)
    shards = [_build(variant, Z, "cpu", Nl, r * Nl, lb0=False)[1] for r in range(world)]
    lperms = [_local_perms(variant, shards[r], r, 11) for r in range(world)]
    if variant == "ppo":
        mbl = Nl * T // NB
        parts = []
        for i in range(NB):
            for r in range(world):
                l = lperms[r][0][i * mbl:(i + 1) * mbl]
                parts.append((l // Nl) * NG + r * Nl + l % Nl)           # local row t * Nl + n -> global row t * NG + r * Nl + n
        losses = alg.update(indices=torch.cat(parts))
    else:
        tn, sn = shards[0].teacher_num_envs * T, shards[0].student_num_envs * T
        tml, sml = tn // NB, sn // NB
        tp = torch.cat([lperms[r][0][i * tml:(i + 1) * tml] + r * tn for i in range(NB) for r in range(world)])
        sp = torch.cat([lperms[r][1][i * sml:(i + 1) * sml] + r * sn for i in range(NB) for r in range(world)])
        losses = alg.update(tp, sp)
    assert np.allclose(np.array(losses), res[0][2], rtol=2e-4, atol=2e-5), (losses, res[0][2])
    assert abs(alg.learning_rate - res[0][3]) < 1e-9
    num = den = 0.0
    for k, v in model.state_dict().items():
        o, d = torch.from_numpy(Z["sd0_" + k]), torch.from_numpy(res[0][1][k])
        num += float(((v - o) - (d - o)).pow(2).sum()); den += float((v - o).pow(2).sum())
    assert (num / den) ** 0.5 < 2e-3, (num / den) ** 0.5
