"""RolloutStorageCTS (drop-in for rsl_rl/storage/rollout_storage_cts.py:36-216): the PPO buffers plus the flattened observation
history, rows kept TEACHER-FIRST (cts.py:126-141), mini-batches = a teacher slice followed by a student slice drawn from
separate permutations over the env-major flattening (rollout_storage_cts.py:153-211)."""
import torch

from .. import _ops
from .rollout_storage import RolloutStorage


class RolloutStorageCTS(RolloutStorage):
    class Transition(RolloutStorage.Transition):
        def __init__(self):
            super().__init__()
            self.history = None

    def __init__(self, num_envs, teacher_num_envs, history_length, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape,
                 device='cpu'):
        super().__init__(num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape, device)
        self.teacher_num_envs = teacher_num_envs
        self.student_num_envs = num_envs - teacher_num_envs
        self.history_length = history_length
        self.history = torch.zeros(num_transitions_per_env, num_envs, history_length * obs_shape[0], device=device)

    def batch_indices(self, num_mini_batches, teacher_perm=None, student_perm=None):
        """-> int64 [num_mini_batches * (tm + sm)] memory-row indices (t * N + n) so that mini-batch i is the contiguous slice i,
        teacher samples first.  Permutations index the env-major flattening f = n * T + t like the reference."""
        T, N = self.num_transitions_per_env, self.num_envs
        nt, ns = self.teacher_num_envs * T, self.student_num_envs * T
        tm, sm = nt // num_mini_batches, ns // num_mini_batches
        if teacher_perm is None:
            teacher_perm = torch.randperm(nt, device=self.device)
        if student_perm is None:
            student_perm = torch.randperm(ns, device=self.device)
        parts = []
        for i in range(num_mini_batches):
            parts.append(teacher_perm[i * tm:(i + 1) * tm])
            parts.append(nt + student_perm[i * sm:(i + 1) * sm])
        f = torch.cat(parts)
        return (f % T) * N + torch.div(f, T, rounding_mode="floor"), tm, sm

    def shuffled(self, indices, pads, transposed=()):
        out = super().shuffled(indices, pads, transposed)
        n = indices.numel()
        w = self.history.shape[-1]
        ld = pads.get("history", w)
        buf = self._sh.get("history")
        if buf is None or buf.shape != (n, ld):
            buf = self._sh["history"] = torch.empty(n, ld, device=self.device)
        bt = None
        if "history" in transposed:
            bt = self._sh.get("history_t")
            if bt is None or bt.shape != (w + 1, n):
                bt = self._sh["history_t"] = torch.ones(w + 1, n, device=self.device)
            out["history_t"] = bt
        _ops.call("go2_gather_rows", _ops.ptr(self.history), w, _ops.ptr(indices), _ops.ptr(buf), ld, _ops.ptr(bt), n)
        out["history"] = buf
        return out

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        """Reference-shaped generator (rollout_storage_cts.py:153-211) for external callers: every mini-batch is a teacher slice followed by a
        student slice, 12-tuples with the history in position 3; CTS.update uses `batch_indices` + `shuffled` (one gather per iteration)."""
        idx, tm, sm = self.batch_indices(num_mini_batches)
        sh = self.shuffled(idx, {})
        mb = tm + sm
        for epoch in range(num_epochs):
            for i in range(num_mini_batches):
                s = slice(i * mb, (i + 1) * mb)
                yield (sh["obs"][s], sh["critic_obs"][s], sh["actions"][s], sh["history"][s], sh["values"][s], sh["adv"][s], sh["returns"][s],
                       sh["old_logp"][s], sh["old_mu"][s], sh["old_sigma"][s], (None, None), None)
