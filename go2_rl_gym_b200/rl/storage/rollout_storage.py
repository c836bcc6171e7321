"""RolloutStorage (drop-in for rsl_rl/storage/rollout_storage.py:36-183): the same [T, N, ·] buffers, filled row by row by
the sampling / process kernels, GAE by a one-thread-per-env reverse scan, and a shuffle-gather that materialises the permuted
data set once per iteration (the reference draws ONE randperm for all epochs, :150, so every epoch reuses the same batches)."""
import torch

from .. import _ops


class RolloutStorage:
    class Transition:
        def __init__(self):
            self.observations = None
            self.critic_observations = None
            self.actions = None
            self.rewards = None
            self.dones = None
            self.values = None
            self.actions_log_prob = None
            self.action_mean = None
            self.action_sigma = None
            self.hidden_states = None

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape, device='cpu'):
        self.device = device
        self.obs_shape, self.privileged_obs_shape, self.actions_shape = obs_shape, privileged_obs_shape, actions_shape
        T, N = num_transitions_per_env, num_envs
        z = lambda *s, **k: torch.zeros(*s, device=device, **k)
        self.observations = z(T, N, *obs_shape)
        self.privileged_observations = z(T, N, *privileged_obs_shape) if privileged_obs_shape[0] is not None else None
        self.rewards = z(T, N, 1)
        self.actions = z(T, N, *actions_shape)
        self.dones = z(T, N, 1, dtype=torch.uint8)
        self.actions_log_prob = z(T, N, 1)
        self.values = z(T, N, 1)
        self.returns = z(T, N, 1)
        self.advantages = z(T, N, 1)
        self.mu = z(T, N, *actions_shape)
        self.sigma = z(T, N, *actions_shape)
        self.num_transitions_per_env, self.num_envs = T, N
        self.step = 0
        self._stats = torch.zeros(2, device=device, dtype=torch.float64)

    def add_transitions(self, transition):
        raise NotImplementedError("transitions are written in place by PPO.act / PPO.process_env_step")

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam, reduce_stats=None):
        """rollout_storage.py:123-137.  reduce_stats(stats) -> global sample count lets a multi-GPU caller all-reduce the
        advantage sums so the normalisation is over ALL ranks' samples, as a single-GPU run of the same envs would do."""
        T, N = self.num_transitions_per_env, self.num_envs
        _ops.call("go2_gae", _ops.ptr(self.rewards), _ops.ptr(self.values), _ops.ptr(self.dones), _ops.ptr(last_values), _ops.ptr(self.returns),
                  _ops.ptr(self.advantages), T, N, gamma, lam, _ops.ptr(self._stats))
        count = float(T * N)
        if reduce_stats is not None:
            count = float(reduce_stats(self._stats))
        _ops.call("go2_adv_normalize", _ops.ptr(self.advantages), T * N, _ops.ptr(self._stats), count)

    def get_statistics(self):
        done = self.dones.clone()
        done[-1] = 1
        flat_dones = done.permute(1, 0, 2).reshape(-1, 1)
        done_indices = torch.cat((flat_dones.new_tensor([-1], dtype=torch.int64), flat_dones.nonzero(as_tuple=False)[:, 0]))
        return (done_indices[1:] - done_indices[:-1]).float().mean(), self.rewards.mean()

    def shuffled(self, indices, pads, transposed=()):
        """Gather every per-sample field in `indices` order (time-major flatten, rollout_storage.py:152-165).
        pads: field -> padded row length; transposed: fields that also get a [width, T*N] transposed copy (key + '_t').
        Returns dict of cached buffers."""
        n = indices.numel()
        if not hasattr(self, "_sh"):
            self._sh = {}
        out = {}
        fields = {"obs": self.observations, "critic_obs": self.privileged_observations if self.privileged_observations is not None else self.observations,
                  "actions": self.actions, "values": self.values, "returns": self.returns, "old_logp": self.actions_log_prob,
                  "adv": self.advantages, "old_mu": self.mu, "old_sigma": self.sigma}
        for k, src in fields.items():
            w = src.shape[-1]
            ld = pads.get(k, w)
            buf = self._sh.get(k)
            if buf is None or buf.shape != (n, ld):
                buf = self._sh[k] = torch.empty(n, ld, device=self.device)
            bt = None
            if k in transposed:
                bt = self._sh.get(k + "_t")
                if bt is None or bt.shape != (w + 1, n):
                    bt = self._sh[k + "_t"] = torch.ones(w + 1, n, device=self.device)   # last row stays 1: bias-gradient column of the wgrad
                out[k + "_t"] = bt
            _ops.call("go2_gather_rows", _ops.ptr(src), w, _ops.ptr(indices), _ops.ptr(buf), ld, _ops.ptr(bt), n)
            out[k] = buf
        return out

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        """Reference-shaped generator (rollout_storage.py:147-183) for external callers; PPO.update uses `shuffled`."""
        batch_size = self.num_envs * self.num_transitions_per_env
        mini_batch_size = batch_size // num_mini_batches
        indices = torch.randperm(num_mini_batches * mini_batch_size, requires_grad=False, device=self.device)
        sh = self.shuffled(indices, {})
        for epoch in range(num_epochs):
            for i in range(num_mini_batches):
                s = slice(i * mini_batch_size, (i + 1) * mini_batch_size)
                yield (sh["obs"][s], sh["critic_obs"][s], sh["actions"][s], sh["values"][s], sh["adv"][s], sh["returns"][s], sh["old_logp"][s],
                       sh["old_mu"][s], sh["old_sigma"][s], (None, None), None)
