"""GPU parity tests of the trainer kernels (through the C ABI) against plain fp32 PyTorch / the RL oracle, and against the
fixture made by the REFERENCE's rsl_rl (tests/golden/rl_ppo.npz).  Tolerances are stated per test."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rl_ppo.npz"))


def _t(k, dev="cuda"):
    return torch.from_numpy(Z[k]).to(dev)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


# Multiply precision of the tensor-core GEMMs (include/go2_b200.h: go2_gemm_set_passes).  3 = 3xTF32 split, the DEFAULT and the path bench.py
# times: fp32-class products, bars below are fp32 bars against an fp64 reference.  1 = one tf32 pass (10-bit mantissas; round 1's kernel, A/B only).
REL_BAR = {3: 4e-6, 1: 2e-3}
# weight gradients contract over the 24576 batch rows (split-K partial sums in fp32, the tensor core adds into the accumulator with truncation): measured
# 4e-6 .. 1e-5 on the B200 (torch fp32 sgemm: 6e-7 on the same data; tools/bench_gemm_trainer.py), so the bar for the long contraction is 2e-5 — 40 x below one tf32 pass
REL_BAR_WGRAD = {3: 2e-5, 1: 2e-3}


@pytest.fixture(params=[3, 1], ids=["3xtf32", "tf32"])
def passes(request):
    from go2_rl_gym_b200.rl import _ops
    L = _ops.lib()
    assert L.go2_gemm_get_passes() == 3                      # the default
    assert L.go2_gemm_set_passes(request.param) == 0
    yield request.param
    assert L.go2_gemm_set_passes(3) == 0


def _tc_call(passes, persistable, name, *args):
    """3xTF32 lives in the persistent TMA-store kernel: operands that kernel cannot take (a row pitch that is not a multiple of 16 bytes) must be
    REFUSED under the default precision, never served by the single-pass kernel silently.  -> False when the call was (rightly) refused."""
    from go2_rl_gym_b200.rl import _ops
    if passes == 3 and not persistable:
        with pytest.raises(RuntimeError, match="alignment rules"):
            _ops.call(name, *args)
        return False
    _ops.call(name, *args)
    return True


@pytest.mark.parametrize("M,N,K,act", [(1000, 512, 45, 1), (4096, 256, 512, 1), (333, 12, 128, 0), (2048, 1, 128, 0), (24576, 512, 263, 1)])
def test_linear_forward(M, N, K, act):
    from go2_rl_gym_b200.rl import _ops
    g = torch.Generator(device="cpu").manual_seed(M + N)
    X, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / math.sqrt(K), torch.randn(N, generator=g)
    ref = torch.nn.functional.linear(X, W, b)
    ref = torch.nn.functional.elu(ref) if act else ref
    Xd, Wd, bd = X.cuda(), W.cuda(), b.cuda()
    Y, Yt = torch.empty(M, N, device="cuda"), torch.empty(N, M, device="cuda")
    _ops.call("go2_linear_forward_simt", Xd.data_ptr(), K, Wd.data_ptr(), K, bd.data_ptr(), Y.data_ptr(), N, Yt.data_ptr(), M, M, N, K, act)
    assert torch.allclose(Y.cpu(), ref, rtol=1e-4, atol=1e-4)      # fp32 FMA chain vs fp32 blocked sum
    assert torch.equal(Yt.t().contiguous(), Y)


@pytest.mark.parametrize("M,N,K,act", [(1000, 512, 48, 1), (4096, 256, 512, 1), (333, 12, 128, 0), (2048, 1, 128, 0), (24576, 512, 264, 1),
                                       (24576, 128, 256, 1), (130, 64, 32, 0), (6148, 2048, 256, 1), (49152, 296, 80, 0)])
def test_linear_forward_tensor_core(M, N, K, act, passes):
    """tcgen05 GEMM vs an fp64 reference.  3xTF32 (default): fp32-class, 4e-6 relative (norm-wise) and 2e-5 per element; one tf32 pass: operands
    rounded to 10 mantissa bits, 2e-3 / 5e-3."""
    from go2_rl_gym_b200.rl import _ops
    g = torch.Generator(device="cpu").manual_seed(M + N)
    X, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / math.sqrt(K), torch.randn(N, generator=g)
    ref = torch.nn.functional.linear(X.double(), W.double(), b.double())
    ref = torch.nn.functional.elu(ref) if act else ref
    Xd, Wd, bd = X.cuda(), W.cuda(), b.cuda()
    Y, Yt = torch.zeros(M, N, device="cuda"), torch.zeros(N, M, device="cuda")
    if not _tc_call(passes, N % 4 == 0 and M % 4 == 0, "go2_linear_forward_tc", Xd.data_ptr(), K, Wd.data_ptr(), K, bd.data_ptr(), Y.data_ptr(), N,
                    Yt.data_ptr(), M, M, N, K, act):
        return
    torch.cuda.synchronize()
    assert _rel(Y.cpu(), ref) < REL_BAR[passes], _rel(Y.cpu(), ref)
    tol = 2e-5 if passes == 3 else 5e-3
    assert torch.allclose(Y.cpu().double(), ref, rtol=tol, atol=tol)
    assert torch.equal(Yt.t().contiguous(), Y)
    # the trainer's own call shape: row-major output only
    Y2 = torch.zeros(M, N, device="cuda")
    if N % 4 == 0:
        _ops.call("go2_linear_forward_tc", Xd.data_ptr(), K, Wd.data_ptr(), K, bd.data_ptr(), Y2.data_ptr(), N, 0, 0, M, N, K, act)
        assert torch.equal(Y2, Y)


@pytest.mark.parametrize("M,N,K", [(8192, 256, 512), (24576, 128, 256), (777 * 4, 12, 128), (24576, 512, 45), (24576, 1, 128), (6144, 512, 264), (4100, 2048, 256)])
def test_linear_backward_tensor_core(M, N, K, passes):
    from go2_rl_gym_b200.rl import _ops
    g = torch.Generator(device="cpu").manual_seed(M + K)
    Xact = torch.nn.functional.elu(torch.randn(M, K, generator=g))
    W = torch.randn(N, K, generator=g) / math.sqrt(K)
    dY = torch.randn(M, N, generator=g)
    dX_ref = (dY.double() @ W.double()) * torch.where(Xact > 0, torch.ones_like(Xact), Xact + 1).double()
    dW_ref = dY.double().t() @ Xact.double()
    bar, bar_w = REL_BAR[passes], REL_BAR_WGRAD[passes]
    dYd, Xd = dY.cuda(), Xact.cuda()
    dYt = dYd.t().contiguous()
    Xt = torch.cat([Xd.t(), torch.ones(1, M, device="cuda")], 0).contiguous()      # [K+1, M]: last row of ones -> bias gradient
    dW, db = torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
    work = torch.empty(64 * ((N + 127) // 128 * 128 if N > 1 else 1) * ((K + 4) // 4 * 4), device="cuda")    # N = 1: small workspace -> legacy slice layout
    if _tc_call(passes, N > 1, "go2_linear_wgrad_tc", dYt.data_ptr(), M, Xt.data_ptr(), M, dW.data_ptr(), K, db.data_ptr(), M, N, K, work.data_ptr(), work.numel()):
        torch.cuda.synchronize()
        assert _rel(dW.cpu(), dW_ref) < bar_w, _rel(dW.cpu(), dW_ref)
        assert _rel(db.cpu(), dY.double().sum(0)) < bar_w, _rel(db.cpu(), dY.double().sum(0))
        dW2 = torch.zeros(N, K, device="cuda")
        _ops.call("go2_linear_wgrad_tc", dYt.data_ptr(), M, Xt.data_ptr(), M, dW2.data_ptr(), K, 0, M, N, K, work.data_ptr(), work.numel())
        torch.cuda.synchronize()
        assert _rel(dW2.cpu(), dW_ref) < bar_w, _rel(dW2.cpu(), dW_ref)
    if N % 4 == 0 and K % 4 == 0:
        Np = N
        Wt = W.t().contiguous().cuda()                        # [K, N]
        # the trainer's call: ELU' operand read ROW-MAJOR (TMA load into the staging buffer), row-major output
        dX = torch.zeros(M, K, device="cuda")
        _ops.call("go2_linear_dgrad_tc", dYd.data_ptr(), N, Wt.data_ptr(), Np, Xd.data_ptr(), K, 0, 0, dX.data_ptr(), K, 0, 0, M, N, K)
        torch.cuda.synchronize()
        assert _rel(dX.cpu(), dX_ref) < (bar if N <= 256 else bar_w), _rel(dX.cpu(), dX_ref)      # contraction over N (up to 2048 here)
        # round 1's form (ELU' operand given TRANSPOSED only, transposed output as well): served by the one-tile kernel, i.e. single-pass tf32 only —
        # under the 3xTF32 default it must be refused, not silently downgraded
        dX1, dXt = torch.zeros(M, K, device="cuda"), torch.zeros(K, M, device="cuda")
        if _tc_call(passes, False, "go2_linear_dgrad_tc", dYd.data_ptr(), N, Wt.data_ptr(), Np, 0, 0, Xt.data_ptr(), M, dX1.data_ptr(), K, dXt.data_ptr(), M, M, N, K):
            torch.cuda.synchronize()
            assert _rel(dX1.cpu(), dX_ref) < bar, _rel(dX1.cpu(), dX_ref)
            assert torch.equal(dXt.t().contiguous(), dX1)


@pytest.mark.parametrize("M,N,K", [(8192, 256, 512), (24576, 128, 256), (777 * 4, 12, 128), (24576, 512, 263), (6148, 2048, 256), (4099, 32, 256), (24576, 1, 128)])
def test_wgrad_from_row_major_operands(M, N, K, passes):
    """MN-major tf32 operands (TMA SWIZZLE_128B_ATOM_32B + UMMA 128B_BASE32B descriptors): dW = dZ^T X and db from the ones column,
    no transposed copies.  Same tolerance as the K-major weight gradient."""
    bar = REL_BAR_WGRAD[passes]
    from go2_rl_gym_b200.rl import _ops
    g = torch.Generator(device="cpu").manual_seed(M + K)
    X = torch.nn.functional.elu(torch.randn(M, K, generator=g))
    dY = torch.randn(M, N, generator=g)
    ldx = (K + 1 + 3) // 4 * 4
    Xp = torch.zeros(M, ldx, device="cuda"); Xp[:, :K] = X.cuda(); Xp[:, K] = 1.0
    ldy = (N + 3) // 4 * 4
    dYp = torch.zeros(M, ldy, device="cuda"); dYp[:, :N] = dY.cuda()
    dW, db = torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
    work = torch.empty(64 * ((N + 127) // 128 * 128) * ((K + 4) // 4 * 4), device="cuda")
    _ops.call("go2_linear_wgrad_tc_rm", dYp.data_ptr(), ldy, Xp.data_ptr(), ldx, dW.data_ptr(), K, db.data_ptr(), M, N, K, work.data_ptr(), work.numel())
    torch.cuda.synchronize()
    ref = dY.double().t() @ X.double()
    assert _rel(dW.cpu(), ref) < bar, _rel(dW.cpu(), ref)
    assert _rel(db.cpu(), dY.double().sum(0)) < bar, _rel(db.cpu(), dY.double().sum(0))
    dW2 = torch.zeros(N, K, device="cuda")
    _ops.call("go2_linear_wgrad_tc_rm", dYp.data_ptr(), ldy, Xp.data_ptr(), ldx, dW2.data_ptr(), K, 0, M, N, K, work.data_ptr(), work.numel())
    torch.cuda.synchronize()
    assert _rel(dW2.cpu(), ref) < bar


@pytest.mark.parametrize("M,N,K", [(24576, 12, 128), (4096, 1, 128), (1001, 16, 100), (5, 8, 32)])
def test_small_n_linear_kernels(M, N, K):
    """Narrow-output Linear layers (actor / critic heads): forward, dgrad (with ELU'), wgrad + bias gradient vs fp64 torch; deterministic."""
    from go2_rl_gym_b200.rl import _ops
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    ldx = K + 4
    Xp = torch.nn.functional.elu(torch.randn(M, ldx, generator=g))
    X = Xp[:, :K]
    W, b, dY = torch.randn(N, K, generator=g) / math.sqrt(K), torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    Xd, Wd, bd, dYd = Xp.cuda(), W.cuda(), b.cuda(), dY.cuda()
    Y = torch.zeros(M, N, device="cuda")
    _ops.call("go2_linear_forward_smalln", Xd.data_ptr(), ldx, Wd.data_ptr(), K, bd.data_ptr(), Y.data_ptr(), N, M, N, K)
    assert torch.allclose(Y.cpu(), (X.double() @ W.double().t() + b.double()).float(), rtol=1e-5, atol=1e-5)
    dX = torch.zeros(M, K, device="cuda")
    _ops.call("go2_linear_dgrad_smalln", dYd.data_ptr(), N, Wd.data_ptr(), K, Xd.data_ptr(), ldx, dX.data_ptr(), K, M, N, K)
    ref = (dY.double() @ W.double()) * torch.where(X > 0, torch.ones_like(X), X + 1).double()
    assert torch.allclose(dX.cpu(), ref.float(), rtol=1e-5, atol=1e-5)
    work = torch.empty(296 * (N * K + N), device="cuda")
    outs = []
    for _ in range(2):
        dW, db = torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
        _ops.call("go2_linear_wgrad_smalln", dYd.data_ptr(), N, Xd.data_ptr(), ldx, dW.data_ptr(), K, db.data_ptr(), M, N, K, work.data_ptr(), work.numel())
        outs.append((dW.clone(), db.clone()))
    scale = math.sqrt(M)
    assert torch.allclose(outs[0][0].cpu(), (dY.double().t() @ X.double()).float(), rtol=1e-4, atol=2e-5 * scale)
    assert torch.allclose(outs[0][1].cpu(), dY.double().sum(0).float(), rtol=1e-4, atol=2e-5 * scale)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("M,K,ldx", [(24576, 128, 132), (1000, 45, 45), (7, 512, 516), (49152, 128, 132)])
def test_wgrad_rank1(M, K, ldx):
    """Weight / bias gradient of the critic's 1-wide head (go2_linear_wgrad_rank1) vs fp64 torch; same bits on a second run."""
    from go2_rl_gym_b200.rl import _ops
    g = torch.Generator(device="cpu").manual_seed(M + K)
    X = torch.randn(M, ldx, generator=g)
    dY = torch.randn(M, 1, generator=g)
    Xd, dYd = X.cuda(), dY.cuda()
    work = torch.empty(300 * (K + 1), device="cuda")
    outs = []
    for _ in range(2):
        dW, db = torch.zeros(1, K, device="cuda"), torch.zeros(1, device="cuda")
        _ops.call("go2_linear_wgrad_rank1", dYd.data_ptr(), 1, Xd.data_ptr(), ldx, dW.data_ptr(), db.data_ptr(), M, K, work.data_ptr(), work.numel())
        outs.append((dW.clone(), db.clone()))
    ref_w = (dY.double().t() @ X[:, :K].double()).float()
    scale = math.sqrt(M)
    assert torch.allclose(outs[0][0].cpu(), ref_w, rtol=1e-4, atol=2e-5 * scale)
    assert torch.allclose(outs[0][1].cpu(), dY.double().sum().float().reshape(1), rtol=1e-4, atol=2e-5 * scale)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("M,N,K", [(1000, 512, 45), (8192, 256, 512), (24576, 128, 256), (777, 12, 128)])
def test_linear_backward(M, N, K):
    from go2_rl_gym_b200.rl import _ops
    g = torch.Generator(device="cpu").manual_seed(M + K)
    X = torch.randn(M, K, generator=g)
    Xact = torch.nn.functional.elu(X)                  # the layer input is a post-ELU activation
    W = torch.randn(N, K, generator=g) / math.sqrt(K)
    dY = torch.randn(M, N, generator=g)
    dX_ref = (dY @ W) * torch.where(Xact > 0, torch.ones_like(Xact), Xact + 1)
    dW_ref, db_ref = dY.t() @ Xact, dY.sum(0)
    dYd, Wd, Xd = dY.cuda(), W.cuda(), Xact.cuda()
    dX, dW, db = torch.empty(M, K, device="cuda"), torch.empty(N, K, device="cuda"), torch.empty(N, device="cuda")
    work = torch.empty(64 * N * K, device="cuda")
    _ops.call("go2_linear_dgrad_simt", dYd.data_ptr(), N, Wd.data_ptr(), K, Xd.data_ptr(), K, dX.data_ptr(), K, 0, 0, M, N, K)
    _ops.call("go2_linear_wgrad_simt", dYd.data_ptr(), N, Xd.data_ptr(), K, dW.data_ptr(), K, 0, M, N, K, work.data_ptr(), work.numel())
    _ops.call("go2_colsum", dYd.data_ptr(), N, db.data_ptr(), M, N, work.data_ptr())
    assert torch.allclose(dX.cpu(), dX_ref, rtol=1e-4, atol=1e-4)
    scale = math.sqrt(M)
    assert torch.allclose(dW.cpu(), dW_ref, rtol=1e-4, atol=2e-4 * scale)
    assert torch.allclose(db.cpu(), db_ref, rtol=1e-4, atol=2e-4 * scale)


def test_gae_matches_reference_fixture():
    from go2_rl_gym_b200.rl.storage import RolloutStorage
    T, N = Z["st_rewards"].shape[:2]
    st = RolloutStorage(N, T, [45], [263], [12], device="cuda")
    st.rewards.copy_(_t("st_rewards")); st.values.copy_(_t("st_values")); st.dones.copy_(_t("st_dones"))
    from oracle import rl_oracle as R
    sd = {k[4:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("sd0_")}
    last_v = R.mlp_forward(sd, "critic", torch.from_numpy(Z["in_priv"][-1])).cuda()
    st.compute_returns(last_v, 0.99, 0.95)
    assert torch.allclose(st.returns.cpu(), torch.from_numpy(Z["st_returns"]), atol=1e-5)
    assert torch.allclose(st.advantages.cpu(), torch.from_numpy(Z["st_advantages"]), atol=2e-5)


@pytest.mark.parametrize("gemm", ["simt", "tc", "tf32"])
def test_ppo_update_matches_reference_fixture(gemm, monkeypatch):
    """Whole PPO.update (5 epochs x 4 mini-batches, adaptive LR, clip + Adam) vs the reference's result on the same data.
    tc = the DEFAULT and benchmarked path (3xTF32 tcgen05 GEMMs) and simt = strict fp32 CUDA-core GEMMs are held to the SAME fp32 bars:
    parameters within 1e-3 rel / 2e-5 abs, relative error of the whole update < 1e-3, losses within 2e-4, Adam moments too.
    tf32 = one tf32 pass (10-bit mantissa operands; A/B only): Adam divides by sqrt(v), which amplifies gradient rounding on near-zero gradients,
    so its bar is on the UPDATE as a whole: 5 % of || ref_new - old ||, losses within 2e-3, identical learning-rate path."""
    from go2_rl_gym_b200.rl import _ops
    monkeypatch.setenv("GO2_GEMM", "simt" if gemm == "simt" else "tc")
    assert _ops.lib().go2_gemm_set_passes(1 if gemm == "tf32" else 3) == 0
    try:
        _ppo_update_case(gemm)
    finally:
        _ops.lib().go2_gemm_set_passes(3)


def _ppo_update_case(gemm):
    from golden.rl_cfg import CFG
    from go2_rl_gym_b200.rl.algorithms import PPO
    from go2_rl_gym_b200.rl.modules import ActorCritic
    T, N = Z["st_rewards"].shape[:2]
    ac = ActorCritic(45, 263, 12, actor_hidden_dims=[64, 32, 16], critic_hidden_dims=[64, 32, 16])
    ac.load_state_dict({k[4:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("sd0_")})
    alg = PPO(ac, device="cuda", **CFG)
    alg.init_storage(N, T, [45], [263], [12])
    st = alg.storage
    for k in ("observations", "privileged_observations", "actions", "rewards", "dones", "values", "returns", "advantages", "actions_log_prob", "mu", "sigma"):
        getattr(st, k).copy_(_t("st_" + k))
    mvl, msl = alg.update(indices=_t("perm"))
    strict = gemm != "tf32"
    tol_l, atol_p = (2e-4, 2e-5) if strict else (2e-3, None)
    assert abs(mvl - float(Z["mean_value_loss"])) < tol_l and abs(msl - float(Z["mean_surrogate_loss"])) < tol_l
    assert abs(alg.learning_rate - float(Z["lr"])) < 1e-9
    worst, num, den = 0.0, 0.0, 0.0
    for k, v in ac.state_dict().items():
        ref, old = torch.from_numpy(Z["sd1_" + k]), torch.from_numpy(Z["sd0_" + k])
        worst = max(worst, float((v.cpu() - ref).abs().max()))
        num += float(((v.cpu() - old) - (ref - old)).pow(2).sum()); den += float((ref - old).pow(2).sum())
        if atol_p is not None:
            assert torch.allclose(v.cpu(), ref, rtol=1e-3, atol=atol_p), (k, float((v.cpu() - ref).abs().max()))
    rel = (num / den) ** 0.5
    print(f"[{gemm}] after 20 optimiser steps: max |param - reference| = {worst:.2e}, relative error of the update = {rel:.3e}")
    assert rel < (1e-3 if strict else 5e-2)
    if strict:
        osd = alg.optimizer_state_dict()
        assert torch.allclose(osd["state"][1]["exp_avg"].cpu(), torch.from_numpy(Z["adam_exp_avg_1"]), rtol=1e-3, atol=1e-6)


def test_act_and_process_env_step():
    from golden.rl_cfg import CFG
    from go2_rl_gym_b200.rl.algorithms import PPO
    from go2_rl_gym_b200.rl.modules import ActorCritic
    N, T = 4096, 24
    torch.manual_seed(0)
    ac = ActorCritic(45, 263, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128])
    sd = {k: v.clone() for k, v in ac.state_dict().items()}
    alg = PPO(ac, device="cuda", **CFG)
    alg.init_storage(N, T, [45], [263], [12])
    obs, priv = torch.randn(N, 45, device="cuda"), torch.randn(N, 263, device="cuda")
    a = alg.act(obs, priv)
    from oracle import rl_oracle as R
    mu = R.mlp_forward(sd, "actor", obs.cpu()); v = R.mlp_forward(sd, "critic", priv.cpu())
    st = alg.storage
    assert torch.allclose(st.mu[0].cpu(), mu, atol=2e-5) and torch.allclose(st.values[0].cpu(), v, atol=2e-5)   # 3xTF32 forward: fp32-class
    z = (a.cpu() - mu) / sd["std"]
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02      # Philox Box-Muller normals
    lp = R.log_prob(st.mu[0].cpu(), sd["std"], a.cpu())
    assert torch.allclose(st.actions_log_prob[0].cpu().squeeze(-1), lp, atol=1e-4)
    rew = torch.randn(N, device="cuda"); dones = torch.rand(N, device="cuda") < 0.1; touts = dones & (torch.rand(N, device="cuda") < 0.5)
    alg.process_env_step(rew, dones, {"time_outs": touts})
    exp = rew + CFG["gamma"] * st.values[0].squeeze(-1) * touts.float()
    assert torch.allclose(st.rewards[0].squeeze(-1), exp, atol=1e-6) and torch.equal(st.dones[0].squeeze(-1).bool(), dones)


def test_runner_learns_two_iterations(tmp_path):
    from go2_rl_gym_b200.envs import task_registry
    from go2_rl_gym_b200.utils import get_args
    args = get_args(["--task", "go2", "--num_envs", "256", "--headless", "--max_iterations", "2"])
    env, env_cfg = task_registry.make_env("go2", args)
    runner, train_cfg = task_registry.make_alg_runner(env, "go2", args, log_root=str(tmp_path))
    runner.learn(2, init_at_random_ep_len=True)
    assert runner.current_learning_iteration == 2
    assert any(f.startswith("model_") for f in os.listdir(runner.log_dir))
    sd = torch.load(os.path.join(runner.log_dir, "model_2.pt"), weights_only=False)
    assert set(sd) == {"model_state_dict", "optimizer_state_dict", "iter", "infos"} and "actor.0.weight" in sd["model_state_dict"]


@pytest.mark.parametrize("task", ["go2", "go2_moe_cts"])
@pytest.mark.parametrize("log", [False, True])
def test_graph_rollout_equals_eager_rollout(task, log):
    """The rollout replayed as ONE CUDA graph over device-resident step parameters (runner.collect) must leave exactly the bits the
    per-step launches leave: env state, counters and every row of the rollout storage, over 3 rollouts (eager, capture + replay, replay) — and
    so must the host-buffer rollout (runner.collect_host, the path bench.py's `e2e` times)."""
    from go2_rl_gym_b200.envs import task_registry
    from go2_rl_gym_b200.utils import get_args
    outs = []
    for mode in ("graph", "eager") + (() if log else ("host",)):
        graphs = mode == "graph"
        args = get_args(["--task", task, "--num_envs", "512", "--headless"])
        env, _ = task_registry.make_env(task, args)
        runner, _ = task_registry.make_alg_runner(env, task, args, log_root=None)
        runner._rollout_graphs.enabled = graphs
        if task != "go2":
            runner._roll_history(env.get_observations(), None)
        env.episode_length_buf = torch.randint(1100, 1250, (512,), generator=torch.Generator().manual_seed(1)).cuda()     # time-outs inside the rollouts
        infos = []
        if mode == "host":
            # the same three rollouts through the HOST-buffer entry point (runner.collect_host -> env.step_host -> go2_env_step_host): actions up,
            # observations / rewards / resets down every step, the policy launches replayed as per-step graphs — and what came down is the device state
            h = [torch.empty(512, 12).pin_memory(), torch.empty(512, 45).pin_memory(), torch.empty(512, 263).pin_memory(), torch.empty(512).pin_memory(),
                 torch.empty(512, dtype=torch.uint8).pin_memory()]
            for _ in range(3):
                runner.collect_host(*h)
            torch.cuda.synchronize()
            assert torch.equal(h[1], env.obs_buf.cpu()) and torch.equal(h[2], env.privileged_obs_buf.cpu()) and torch.equal(h[3], env.rew_buf.cpu())
            assert ("act", 23) in runner._host_graphs._g and not runner._host_graphs._failed
        else:
            for _ in range(3):
                infos = runner.collect(log)
            torch.cuda.synchronize()
            assert len(infos) == (24 if log else 0)
        if graphs:
            assert ("rollout", log) in runner._rollout_graphs._g and not runner._rollout_graphs._failed      # the graph path really ran
        st = runner.alg.storage
        out = {k: getattr(st, k).clone() for k in ("observations", "privileged_observations", "actions", "rewards", "dones", "values", "mu", "actions_log_prob")}
        out.update(root=env.root_states.clone(), q=env.dof_pos.clone(), ep_len=env.episode_length_buf.clone(), levels=env.terrain_levels.clone(),
                   cmd=env.commands.clone(), counter=env.common_step_counter, act_step=runner.alg._act_step, st_step=st.step)
        if log:
            out.update(done_rew=torch.nan_to_num(runner._done_rew.clone(), nan=-1e9), rew_sum=runner._cur_reward_sum.clone())
        outs.append(out)
        assert int(st.dones.sum()) > 50
        del runner, env
    for o in outs[1:]:
        for k in outs[0]:
            a, b = outs[0][k], o[k]
            assert (torch.equal(a, b) if torch.is_tensor(a) else a == b), k


@pytest.mark.parametrize("task", ["go2", "go2_moe_cts"])
def test_play_loop_and_policy_export(task, tmp_path):
    """legged_gym/scripts/play.py (reference play.py:15-66): 7 x 7 non-curriculum terrain, noise / pushes / randomisation off, deterministic
    act_inference for a few steps, TorchScript export whose output matches the CUDA inference path on the same observations."""
    import importlib.util
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
    m = torch.jit.load(os.path.join(tmp_path, "policy.pt"))
    model = runner.alg.actor_critic if task == "go2" else runner.alg.model
    obs = torch.randn(1, 45)
    out = m(obs)
    out = out[0] if isinstance(out, tuple) else out
    if task == "go2":       # feed-forward actor: the exported module and the CUDA inference path see the same input
        ref = model.act_inference(obs.cuda().repeat(64, 1))[0:1].cpu()
        assert torch.allclose(out, ref, atol=5e-3)          # tf32 GEMMs on the CUDA side
    else:
        assert out.shape == (1, 12) and torch.isfinite(out).all()


def test_go2_learns_on_the_gpu():
    """The B200 path TRAINS (missing item 1; the reference logs the same statistics at on_policy_runner.py:203-207): 200 PPO iterations of
    --task=go2 at 4096 envs through task_registry (tools/train_gpu_curve.py).  The CPU counterpart — the oracle env under the UNMODIFIED reference PPO,
    1024 envs (profiles/r01n_cpu_oracle_learning_curve_relaxed.txt) — goes from episode length 13 / return -0.9 to ~1200 / +4 within 150 iterations."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from train_gpu_curve import train
    rows = train("go2", 4096, 200, report=lambda s: None)
    first, last = rows[0], rows[-1]
    print(f"first iteration: {first}\nlast iteration:  {last}")
    assert first["len"] < 100 and first["std"] > 0.95
    # robots stay up for most of the 1250-step episode, positive return (profiles/r02a_gpu_learning_curve_go2.txt: the 100-episode mean length swings
    # between 980 and 1240 from iteration 75 on, the return is +10 at iteration 200)
    assert last["len"] > 850 and last["ret"] > 2.0, last
    assert last["rew_step"] > first["rew_step"] + 0.03 and last["std"] < 0.7   # reward per step up from ~-0.055, action noise annealed
    assert all(r["lr"] >= 1e-5 - 1e-12 and r["lr"] <= 1e-2 + 1e-12 for r in rows)


@pytest.mark.parametrize("world,two_shot", [(1, False), (2, False), (4, False), (4, True), (8, True)])
def test_p2p_allreduce_kernel_protocol(world, two_shot):
    """csrc/dist_kernels.cu on ONE GPU: `world` buffers play the ranks' symmetric buffers (every 'peer' pointer is local), each rank's kernel runs on its
    own stream, all co-resident, so the ready / done flag protocol, the epoch counter and the rank-ordered sums are exercised for several exchanges in
    a row — including a slice exchange (CTS pass 2) and a rank that arrives late.  Every rank must hold the bit-identical sum.  two_shot: the
    result buffers are handed over as well, so each rank sums one slice and stores it into every rank's result (the >= 4 rank schedule)."""
    import ctypes as C
    from go2_rl_gym_b200.rl import _ops
    n, TAIL, FL = 4096 * 12 + 8, 32, 32
    g = torch.Generator(device="cuda").manual_seed(world)
    bufs = [torch.zeros(TAIL + n + FL, device="cuda") for _ in range(world)]
    outs = [torch.zeros(TAIL + n, device="cuda") for _ in range(world)]
    ctrs = [torch.zeros(2, dtype=torch.int32, device="cuda") for _ in range(world)]
    data = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
    flags = (C.c_void_p * world)(*[b.data_ptr() + 4 * (TAIL + n) for b in bufs])
    peer_out = (C.c_void_p * world)(*[o.data_ptr() for o in outs]) if two_shot else None
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for it, (off, cnt) in enumerate([(0, TAIL + n), (0, TAIL + n), (TAIL + 1024, n - 1024), (0, TAIL + 4096)]):
        for b in bufs:
            b[:TAIL + n].copy_(torch.randn(TAIL + n, device="cuda", generator=g))
        for o in outs:
            o.fill_(-7.0)
        torch.cuda.synchronize()
        for r in reversed(range(world)):          # the last rank launches first, rank 0 after a delay
            with torch.cuda.stream(streams[r]):
                if r == 0:
                    torch.cuda._sleep(2_000_000)
                _ops.call("go2_allreduce_p2p2", data, flags, peer_out, outs[r].data_ptr(), off, cnt, r, world, ctrs[r].data_ptr())
        torch.cuda.synchronize()
        ref = bufs[0][off:off + cnt].clone()
        for b in bufs[1:]:
            ref += b[off:off + cnt]               # same order as the kernel: rank 0, 1, ...
        for r in range(world):
            assert torch.equal(outs[r][off:off + cnt], ref), (it, r)
            assert (outs[r][:off] == -7.0).all() and (outs[r][off + cnt:] == -7.0).all()      # nothing outside the slice was touched
            assert ctrs[r].tolist() == [it + 1, 0]


@pytest.mark.parametrize("M,N,K,act", [(24576, 512, 264, 1), (24576, 256, 512, 1), (8192, 128, 256, 0), (3000, 268, 64, 0), (24576, 516, 48, 1)])
def test_pair_kernel_equals_one_cta_kernel(M, N, K, act):
    """The CTA-pair GEMM (tcgen05.mma.cta_group::2, 256-row tiles, half of B per SM) and the one-CTA persistent kernel run the same 3xTF32 arithmetic in
    the same K order per output element: bit-identical outputs, for every pair tile width (256 / 192 / 160 / 128), ragged M and N included."""
    from go2_rl_gym_b200.rl import _ops
    L = _ops.lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    X, W, b = torch.randn(M, K, device="cuda", generator=g), torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K), torch.randn(N, device="cuda", generator=g)
    outs = []
    try:
        for pair in (1, 0):
            assert L.go2_gemm_set_pair(pair) == 0
            Y = torch.zeros(M, N, device="cuda")
            _ops.call("go2_linear_forward_tc", X.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N, 0, 0, M, N, K, act)
            torch.cuda.synchronize()
            outs.append(Y)
    finally:
        L.go2_gemm_set_pair(1)
    assert torch.equal(outs[0], outs[1])
    ref = torch.nn.functional.linear(X.double(), W.double(), b.double())
    ref = torch.nn.functional.elu(ref) if act else ref
    assert _rel(outs[0].cpu(), ref.cpu()) < REL_BAR[3]


def test_side_streams_do_not_change_the_update(monkeypatch):
    """Actor / critic chains and the weight gradients run on side streams joined as graph edges (rl/_ops.py: SideStream).  Concurrency must not change a
    bit: PPO.update over the reference-made fixture with GO2_TWO_STREAMS=0 (everything on one stream) and =1, three times each (a race would show
    as run-to-run differences), must leave identical parameters and losses."""
    from golden.rl_cfg import CFG
    from go2_rl_gym_b200.rl.algorithms import PPO
    from go2_rl_gym_b200.rl.modules import ActorCritic
    T, N = Z["st_rewards"].shape[:2]
    results = []
    for ts in ("0", "1", "1", "1"):
        monkeypatch.setenv("GO2_TWO_STREAMS", ts)
        ac = ActorCritic(45, 263, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128])
        torch.manual_seed(3)
        for p in ac.parameters():
            p.data.copy_(torch.randn(p.shape) * 0.05)
        alg = PPO(ac, device="cuda", **CFG)
        alg.init_storage(N, T, [45], [263], [12])
        st = alg.storage
        for k in ("observations", "privileged_observations", "actions", "rewards", "dones", "values", "returns", "advantages", "actions_log_prob", "mu", "sigma"):
            getattr(st, k).copy_(_t("st_" + k))
        for _ in range(2):       # second call = the captured graph
            st.step = T
            losses = alg.update(indices=_t("perm"))
        assert alg._side.enabled == (ts == "1")
        results.append((ac.flat_params.clone(), losses))
    for r in results[1:]:
        assert torch.equal(r[0], results[0][0]) and r[1] == results[0][1]
