"""The runners' own Python (train.py's call sequence: learn -> log -> save -> load -> inference policy) for all seven registered tasks, on CPU:
a stand-in VecEnv with random dynamics replaces the Go2 env (the real one needs the GPU), the trainer's C ABI is emulated by tests/emu_rl.py, and
the registered train configs are used with narrower hidden layers.  Catches host-side errors (class maps, loss-tuple lengths, logging, checkpoint
keys, optimiser state round trips) before the GPU suite exercises the same code."""
import copy
import os
from types import SimpleNamespace

import pytest
import torch

import emu_rl

TASKS = ["go2", "go2_cts", "go2_moe_cts", "go2_moe_ng_cts", "go2_ac_moe_cts", "go2_dual_moe_cts", "go2_mcp_cts"]


class StubEnv:
    """VecEnv contract (rsl_rl/env/vec_env.py:36-59) with random observations / rewards / terminations on CPU."""

    def __init__(self, num_envs, seed=0):
        self.num_envs, self.num_obs, self.num_privileged_obs, self.num_actions = num_envs, 45, 263, 12
        self.max_episode_length = 1000
        self.device = "cpu"
        self.cfg = SimpleNamespace(env=SimpleNamespace(test=False))
        self.g = torch.Generator().manual_seed(seed)
        self.episode_length_buf = torch.zeros(num_envs, dtype=torch.long)
        self.extras = {}
        self._draw()

    def _draw(self):
        self.obs_buf = torch.randn(self.num_envs, 45, generator=self.g)
        self.privileged_obs_buf = torch.randn(self.num_envs, 263, generator=self.g)

    def reset(self):
        return self.obs_buf, self.privileged_obs_buf

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def step(self, actions):
        assert actions.shape == (self.num_envs, 12) and torch.isfinite(actions).all()
        self._draw()
        self.rew_buf = 0.01 * torch.randn(self.num_envs, generator=self.g)
        self.reset_buf = torch.rand(self.num_envs, generator=self.g) < 0.05
        self.extras = {"time_outs": self.reset_buf & (torch.rand(self.num_envs, generator=self.g) < 0.5),
                       "episode": {"rew_tracking_lin_vel": torch.tensor(0.1), "terrain_level_all": torch.tensor(2.0)}}
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras


def _narrow(d):
    """Registered train config with small hidden layers (the wirings, not the widths, are under test)."""
    d = copy.deepcopy(d)
    p = d["policy"]
    for k, v in (("actor_hidden_dims", [64, 32, 16]), ("critic_hidden_dims", [64, 32, 16]), ("teacher_encoder_hidden_dims", [64, 32])):
        if k in p:
            p[k] = v
    if "student_encoder_hidden_dims" in p:
        p["student_encoder_hidden_dims"] = [64, 32, 32][:len(p["student_encoder_hidden_dims"])]
    d["runner"]["save_interval"] = 1
    return d


@pytest.mark.parametrize("task", TASKS)
def test_runner_learn_save_load_on_a_stub_env(task, tmp_path, monkeypatch):
    emu_rl.install(monkeypatch)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    from go2_rl_gym_b200.envs import task_registry
    from go2_rl_gym_b200.rl import runners
    from go2_rl_gym_b200.utils.cfg_dict import class_to_dict
    _, train_cfg = task_registry.get_cfgs(task)
    cfg = _narrow(class_to_dict(train_cfg))
    N = 32
    make = lambda log: getattr(runners, train_cfg.runner_class_name)(StubEnv(N), cfg, log, device="cpu")
    runner = make(str(tmp_path))
    runner.learn(2, init_at_random_ep_len=True)
    assert runner.current_learning_iteration == 2
    path = os.path.join(str(tmp_path), "model_2.pt")
    sd = torch.load(path, weights_only=False)
    cts = task != "go2"
    assert set(sd) == ({"model_state_dict", "optimizer1_state_dict", "optimizer2_state_dict", "iter", "infos"} if cts else
                       {"model_state_dict", "optimizer_state_dict", "iter", "infos"})
    assert all(torch.isfinite(v).all() for v in sd["model_state_dict"].values())
    model = runner.alg.model if cts else runner.alg.actor_critic
    before = {k: v.clone() for k, v in model.state_dict().items()}
    runner2 = make(None)
    runner2.load(path)
    model2 = runner2.alg.model if cts else runner2.alg.actor_critic
    assert runner2.current_learning_iteration == 2
    for k, v in model2.state_dict().items():
        assert torch.equal(v, before[k]), k
    if cts:     # both optimiser states survive the round trip
        a, b = runner.alg.optimizer1_state_dict(), runner2.alg.optimizer1_state_dict()
        assert torch.equal(a["state"][0]["exp_avg"], b["state"][0]["exp_avg"]) and float(a["state"][0]["step"]) == float(b["state"][0]["step"]) == 40.0
    runner2.learn(1)          # resume training from the checkpoint
    policy = runner.get_inference_policy()
    act = policy(torch.randn(N, 45))
    assert act.shape == (N, 12) and torch.isfinite(act).all()
