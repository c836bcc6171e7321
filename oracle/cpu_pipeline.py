"""CPU restatement of one PPO iteration of the reference pipeline — TEST / BASELINE INFRASTRUCTURE ONLY.

24 x (policy forward + sample -> oracle env step -> bookkeeping) + GAE + PPO.update, i.e. OnPolicyRunner.learn's loop body
(rsl_rl/runners/on_policy_runner.py:132-163) with the reference's algorithm restated on the host:
  env      oracle/go2_oracle.cpp (OpenMP over envs for the physics)        [PhysX itself is unavailable: kind = "port"]
  RL       oracle/rl_oracle.py   (fp32 PyTorch autograd + torch.optim.Adam, all host threads)
bench.py times it as `cpu_baseline` and as the `--impl reference` arm."""
import time

import torch

from go2_rl_gym_b200.envs.env_arrays import EnvArrays
from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
from oracle import rl_oracle as R
from oracle.oracle import OracleEnv

CFG = dict(value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.01, num_learning_epochs=5, num_mini_batches=4,
           learning_rate=1e-3, schedule="adaptive", gamma=0.99, lam=0.95, desired_kl=0.01, max_grad_norm=1.0)


class CpuPipeline:
    def __init__(self, num_envs, mesh_type="heightfield", seed=1, threads=None):
        if threads:
            torch.set_num_threads(threads)
        cfg = GO2Cfg()
        cfg.env.num_envs = num_envs
        cfg.terrain.mesh_type = mesh_type
        cfg.seed = seed
        self.A = EnvArrays(cfg, "cpu", seed=seed)
        self.env = OracleEnv(self.A)
        self.env.reset_all()
        self.env.step(torch.zeros(num_envs, 12))
        torch.manual_seed(seed)
        import torch.nn as nn

        def mlp(dims):
            layers = []
            for i in range(len(dims) - 1):
                layers.append(nn.Linear(dims[i], dims[i + 1]))
                if i < len(dims) - 2:
                    layers.append(nn.ELU())
            return nn.Sequential(*layers)
        actor, critic = mlp([45, 512, 256, 128, 12]), mlp([263, 512, 256, 128, 1])
        self.sd = {"std": torch.ones(12)}
        self.sd.update({f"actor.{k}": v.detach().clone() for k, v in actor.state_dict().items()})
        self.sd.update({f"critic.{k}": v.detach().clone() for k, v in critic.state_dict().items()})
        self.N, self.T = num_envs, 24
        self.gen = torch.Generator().manual_seed(seed)

    @torch.no_grad()
    def _rollout(self):
        T, N, A = self.T, self.N, self.A.tensors
        st = {k: torch.zeros(T, N, d) for k, d in (("obs", 45), ("critic_obs", 263), ("actions", 12), ("old_mu", 12), ("old_sigma", 12),
                                                     ("rewards", 1), ("values", 1), ("old_logp", 1))}
        dones = torch.zeros(T, N, 1, dtype=torch.uint8)
        for t in range(T):
            obs, priv = A["obs_buf"].clone(), A["privileged_obs_buf"].clone()
            mu = R.mlp_forward(self.sd, "actor", obs)
            v = R.mlp_forward(self.sd, "critic", priv)
            std = self.sd["std"]
            act = mu + std * torch.randn(mu.shape, generator=self.gen)
            st["obs"][t], st["critic_obs"][t], st["actions"][t], st["old_mu"][t], st["old_sigma"][t] = obs, priv, act, mu, std.expand_as(mu)
            st["values"][t] = v
            st["old_logp"][t] = R.log_prob(mu, std, act).unsqueeze(-1)
            self.env.step(act)
            rew = A["rew_buf"].clone() + CFG["gamma"] * v.squeeze(-1) * A["time_out_buf"].float()
            st["rewards"][t] = rew.unsqueeze(-1)
            dones[t] = A["reset_buf"].unsqueeze(-1)
        last_v = R.mlp_forward(self.sd, "critic", A["privileged_obs_buf"].clone())
        ret, adv = R.gae(st["rewards"], st["values"], dones, last_v, CFG["gamma"], CFG["lam"])
        st["returns"], st["adv"] = ret, adv
        return st

    def iteration(self):
        """-> (env_steps, seconds_collect, seconds_learn)"""
        t0 = time.time()
        st = self._rollout()
        t1 = time.time()
        data = {k: v.flatten(0, 1) for k, v in st.items()}
        perm = torch.randperm(self.T * self.N, generator=self.gen)
        self.sd, _, _, _, _ = R.ppo_update(self.sd, data, perm, CFG)
        t2 = time.time()
        return self.N * self.T, t1 - t0, t2 - t1
