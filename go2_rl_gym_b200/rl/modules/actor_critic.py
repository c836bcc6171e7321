"""ActorCritic (drop-in for rsl_rl/modules/actor_critic.py:38-136).

Same constructor, same `state_dict` keys (`std`, `actor.{0,2,4,6}.*`, `critic.{0,2,4,6}.*`), same methods.  The module
still owns ordinary `nn.Parameter`s (so checkpoints, `.parameters()` and `torch.save` behave as in the reference), but every
parameter is a view into ONE flat fp32 device vector, and forward / backward run on the library's GEMM kernels rather than
autograd: gradients land in a second flat vector that the fused clip+Adam kernel and the NCCL all-reduce consume whole."""
import torch
import torch.nn as nn

from .. import _ops, dist_utils


def _mlp(dims):
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append(nn.ELU())
    return nn.Sequential(*layers)


class ActorCritic(nn.Module):
    is_recurrent = False

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=[256, 256, 256],
                 critic_hidden_dims=[256, 256, 256], activation='elu', init_noise_std=1.0, **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str([key for key in kwargs.keys()]))
        super().__init__()
        if activation != 'elu':
            raise NotImplementedError("the fused epilogues implement ELU (every go2 task uses it: legged_robot_config.py:268)")
        self.actor_dims = [num_actor_obs] + list(actor_hidden_dims) + [num_actions]
        self.critic_dims = [num_critic_obs] + list(critic_hidden_dims) + [1]
        self.actor = _mlp(self.actor_dims)       # nn.Linear default init, built on the CPU like the reference (on_policy_runner.py:78-81)
        self.critic = _mlp(self.critic_dims)
        print(f"Actor MLP: {self.actor}")
        print(f"Critic MLP: {self.critic}")
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        self.num_actions = num_actions
        self._flat = None
        self._mean = None

    # ---- flat parameter storage ------------------------------------------------------------------------------
    def flatten_(self, device, max_rows, train_rows=0):
        """Move to `device`, re-home every parameter inside one flat vector (parameters() order: std, actor.*, critic.*)."""
        n = sum((p.numel() + 3) // 4 * 4 for p in self.parameters())     # every parameter starts 16-byte aligned (TMA operand rule)
        flat = torch.zeros(n, device=device, dtype=torch.float32)
        grad = dist_utils.new_flat_grad(n, device)        # a symmetric (peer-mapped) buffer when the envs are sharded over GPUs
        off = 0
        self._views, self._gviews, self._offsets = {}, {}, {}
        for name, p in self.named_parameters():
            k = p.numel()
            self._offsets[name] = off
            flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = flat[off:off + k].view(p.shape)
            self._views[name] = p.data
            self._gviews[name] = grad[off:off + k].view(p.shape)
            off += (k + 3) // 4 * 4
        self._flat, self._grad = flat, grad
        self.device = torch.device(device)

        def engine(seq_name, dims):
            idx = [i for i in range(0, 2 * (len(dims) - 1), 2)]
            W = [self._views[f"{seq_name}.{i}.weight"] for i in idx]
            b = [self._views[f"{seq_name}.{i}.bias"] for i in idx]
            gW = [self._gviews[f"{seq_name}.{i}.weight"] for i in idx]
            gb = [self._gviews[f"{seq_name}.{i}.bias"] for i in idx]
            return _ops.MlpEngine(dims, W, b, gW, gb, max_rows, device, train_rows=train_rows)

        self.actor_engine = engine("actor", self.actor_dims)
        self.critic_engine = engine("critic", self.critic_dims)
        self._mu_buf = torch.empty(max_rows, self.num_actions, device=device)
        self._val_buf = torch.empty(max_rows, 1, device=device)
        return self

    @property
    def flat_params(self):
        return self._flat

    @property
    def flat_grads(self):
        return self._grad

    def load_state_dict(self, state_dict, strict=True):
        """Checkpoints load IN PLACE so the flat vector stays the storage."""
        if self._flat is None:
            return super().load_state_dict(state_dict, strict)
        own = dict(self.named_parameters())
        missing = [k for k in own if k not in state_dict]
        unexpected = [k for k in state_dict if k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing}, unexpected {unexpected}")
        with torch.no_grad():
            for k, p in own.items():
                if k in state_dict:
                    p.data.copy_(state_dict[k].to(p.device))
        self.actor_engine.mark_dirty(); self.critic_engine.mark_dirty()
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    # ---- reference API ---------------------------------------------------------------------------------------
    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    @property
    def action_mean(self):
        return self._mean

    @property
    def action_std(self):
        return self.std.data.unsqueeze(0).expand_as(self._mean)

    @property
    def entropy(self):
        return (0.5 + 0.5 * torch.log(torch.tensor(2 * torch.pi, device=self.device)) + torch.log(self.std.data)).sum().expand(self._mean.shape[0])

    def _actor_forward(self, observations, out=None, save=False):
        obs = observations if observations.is_contiguous() else observations.contiguous()
        M = obs.shape[0]
        out = self._mu_buf[:M] if out is None else out
        self.actor_engine.forward(obs, obs.shape[1], M, out, self.num_actions)
        self._mean = out
        return out

    def update_distribution(self, observations):
        self._actor_forward(observations)

    def act_inference(self, observations):
        return self._actor_forward(observations).clone()

    def evaluate(self, critic_observations, out=None, **kwargs):
        obs = critic_observations if critic_observations.is_contiguous() else critic_observations.contiguous()
        M = obs.shape[0]
        out = self._val_buf[:M] if out is None else out
        self.critic_engine.forward(obs, obs.shape[1], M, out, 1)
        return out

    def act(self, observations, **kwargs):
        """actor_critic.py:123-125 for external callers (torch's generator); PPO.act samples in its own kernel, fused with the transition write."""
        self.update_distribution(observations)
        return torch.normal(self._mean, self.action_std)

    def get_actions_log_prob(self, actions):
        mu, std = self._mean, self.std.data
        return (-((actions - mu) ** 2) / (2 * std * std) - torch.log(std) - 0.9189385332046727).sum(-1)
