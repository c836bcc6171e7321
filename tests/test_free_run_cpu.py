"""Calibration of the free-running statistical comparison (tests/free_run.py) on the CPU: the oracle in fp32 against the SAME oracle in fp64 — two
roundings of one algorithm, never re-synchronised — must pass the bars the GPU test (tests/test_gpu_env.py::test_free_running_statistics_match_oracle)
applies to the CUDA kernel against the fp32 oracle."""
import torch

from go2_rl_gym_b200.envs.env_arrays import EnvArrays
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
from oracle.oracle import OracleEnv
import free_run


def test_free_run_fp32_oracle_vs_fp64_oracle():
    N, steps = 1024, 400
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 1
    Aa, Ab = EnvArrays(cfg, "cpu", seed=1), EnvArrays(cfg, "cpu", seed=1)
    ea, eb = OracleEnv(Aa), OracleEnv(Ab, double=True)
    for e, A in ((ea, Aa), (eb, Ab)):
        e.common_step_counter = 24 * 300
        e.reset_all()
        A.tensors["episode_length_buf"].copy_(torch.randint(0, 1250, (N,), generator=torch.Generator().manual_seed(2)).int())
    lines = []
    bad = free_run.run_pair(ea, eb, Aa.tensors, Ab.tensors, N, steps, report=lines.append)
    print("\n".join(lines))
    assert not bad, bad


def test_free_run_statistics_detect_a_different_process():
    """the bars are not vacuous: the same oracle with a different friction / solver setting fails them"""
    N, steps = 1024, 400
    cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"; cfg.seed = 1
    cfg2 = GO2Cfg(); cfg2.env.num_envs = N; cfg2.terrain.mesh_type = "heightfield"; cfg2.seed = 1
    cfg2.terrain.static_friction = 0.05; cfg2.domain_rand.friction_range = [0.0, 0.1]
    Aa, Ab = EnvArrays(cfg, "cpu", seed=1), EnvArrays(cfg2, "cpu", seed=1)
    ea, eb = OracleEnv(Aa), OracleEnv(Ab)
    for e, A in ((ea, Aa), (eb, Ab)):
        e.common_step_counter = 24 * 300
        e.reset_all()
        A.tensors["episode_length_buf"].copy_(torch.randint(0, 1250, (N,), generator=torch.Generator().manual_seed(2)).int())
    bad = free_run.run_pair(ea, eb, Aa.tensors, Ab.tensors, N, steps)
    assert bad
