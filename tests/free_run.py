"""Free-running (un-synchronised) comparison of two implementations of the env step: both start from the same reset, receive the same action stream
and are NEVER re-synchronised.  Contact dynamics are chaotic, so after a few hundred steps the two trajectories are different samples of the same
process; what must agree is the DISTRIBUTION of what training consumes: episode lengths, time-out / termination rates, terrain levels, rewards and
episode returns.  Used by tests/test_free_run_cpu.py (oracle fp32 vs oracle fp64: calibrates the bars) and tests/test_gpu_env.py (CUDA vs oracle at
4096 envs, 1000 steps)."""
import numpy as np
import torch

SCALES = (0.0, 0.25, 0.5, 1.0)      # action noise of env group g = e % 4: standing still ... N(0,1) (robots of the last groups fall within ~1 s)


def action_stream(N, steps, seed=0):
    g = torch.Generator().manual_seed(seed)
    scale = torch.tensor(SCALES)[torch.arange(N) % len(SCALES)].unsqueeze(1)
    for _ in range(steps):
        yield scale * torch.randn(N, 12, generator=g)


class Stats:
    """per-group episode statistics of one implementation, accumulated on the host from its public buffers after every step"""

    def __init__(self, N):
        self.N, self.G = N, len(SCALES)
        self.group = np.arange(N) % self.G
        self.ret = np.zeros(N); self.len = np.zeros(N, np.int64)
        self.ep_len = [[] for _ in range(self.G)]; self.ep_ret = [[] for _ in range(self.G)]
        self.n_tout = np.zeros(self.G, np.int64); self.n_term = np.zeros(self.G, np.int64)
        self.rew_sum = np.zeros(self.G); self.rew_n = np.zeros(self.G)
        self.nonfinite = 0

    def add(self, rew, reset, tout):
        rew, reset, tout = np.asarray(rew, np.float64), np.asarray(reset).astype(bool), np.asarray(tout).astype(bool)
        self.nonfinite += int((~np.isfinite(rew)).sum())
        self.ret += rew; self.len += 1
        for g in range(self.G):
            m = self.group == g
            self.rew_sum[g] += rew[m].sum(); self.rew_n[g] += m.sum()
            d = m & reset
            self.ep_len[g] += self.len[d].tolist(); self.ep_ret[g] += self.ret[d].tolist()
            self.n_tout[g] += int((d & tout).sum()); self.n_term[g] += int((d & ~tout).sum())
        self.ret[reset] = 0; self.len[reset] = 0

    def summary(self, levels):
        levels = np.asarray(levels)
        out = []
        for g in range(self.G):
            L, R = np.array(self.ep_len[g] or [0.0]), np.array(self.ep_ret[g] or [0.0])
            out.append({"episodes": len(self.ep_len[g]), "len_mean": float(L.mean()), "len_q": np.quantile(L, [0.1, 0.5, 0.9]).tolist(),
                        "ret_mean": float(R.mean()), "rew_mean": float(self.rew_sum[g] / max(self.rew_n[g], 1)),
                        "timeouts": int(self.n_tout[g]), "terminations": int(self.n_term[g]),
                        "level_mean": float(levels[self.group == g].mean())})
        return out


def ks(a, b):
    """two-sample Kolmogorov-Smirnov statistic"""
    a, b = np.sort(np.asarray(a, np.float64)), np.sort(np.asarray(b, np.float64))
    if len(a) == 0 or len(b) == 0:
        return 0.0 if len(a) == len(b) else 1.0
    x = np.concatenate([a, b])
    return float(np.abs(np.searchsorted(a, x, side="right") / len(a) - np.searchsorted(b, x, side="right") / len(b)).max())


def compare(sa, sb, levels_a, levels_b, N, steps):
    """-> list of violated bars (empty = the two free runs are statistically the same process)"""
    A, B = sa.summary(levels_a), sb.summary(levels_b)
    bad = []
    if sa.nonfinite or sb.nonfinite:
        bad.append(("nonfinite rewards", sa.nonfinite, sb.nonfinite))
    for g, (x, y) in enumerate(zip(A, B)):
        n = min(x["episodes"], y["episodes"])
        # sampling noise of an episode statistic ~ 1 / sqrt(episodes); the bars are 4 sigma-ish plus a small systematic allowance
        slack = 4.0 / np.sqrt(max(n, 1))
        if abs(x["episodes"] - y["episodes"]) > 0.03 * max(n, 1) + 4 * np.sqrt(max(n, 1)):
            bad.append((g, "episodes", x["episodes"], y["episodes"]))
        if n >= 50:
            d = ks(sa.ep_len[g], sb.ep_len[g])
            if d > 0.03 + 1.5 * slack / 4:
                bad.append((g, "KS episode length", d))
            d = ks(sa.ep_ret[g], sb.ep_ret[g])
            if d > 0.03 + 1.5 * slack / 4:
                bad.append((g, "KS episode return", d))
            if abs(x["len_mean"] - y["len_mean"]) > (0.03 + slack) * max(x["len_mean"], y["len_mean"]):
                bad.append((g, "len_mean", x["len_mean"], y["len_mean"]))
        if abs(x["rew_mean"] - y["rew_mean"]) > 0.03 * max(abs(x["rew_mean"]), abs(y["rew_mean"])) + 2e-3:
            bad.append((g, "rew_mean", x["rew_mean"], y["rew_mean"]))
        if abs(x["level_mean"] - y["level_mean"]) > 0.15:
            bad.append((g, "level_mean", x["level_mean"], y["level_mean"]))
        if abs(x["timeouts"] - y["timeouts"]) > 0.05 * max(x["timeouts"], y["timeouts"]) + 4 * np.sqrt(max(x["timeouts"], y["timeouts"], 1)):
            bad.append((g, "timeouts", x["timeouts"], y["timeouts"]))
    return bad, A, B


def run_pair(env_a, env_b, Ta, Tb, N, steps, seed=0, report=None):
    """advance both envs over the same action stream without ever re-synchronising them"""
    sa, sb = Stats(N), Stats(N)
    first_div = None
    for k, a in enumerate(action_stream(N, steps, seed)):
        env_a.step(a); env_b.step(a)
        ra, rb = Ta["reset_buf"].cpu().numpy(), Tb["reset_buf"].cpu().numpy()
        sa.add(Ta["rew_buf"].cpu().numpy(), ra, Ta["time_out_buf"].cpu().numpy())
        sb.add(Tb["rew_buf"].cpu().numpy(), rb, Tb["time_out_buf"].cpu().numpy())
        if first_div is None and not np.array_equal(ra.astype(bool), rb.astype(bool)):
            first_div = k
    bad, A, B = compare(sa, sb, Ta["terrain_levels"].cpu().numpy(), Tb["terrain_levels"].cpu().numpy(), N, steps)
    if report is not None:
        report(f"free run: {N} envs x {steps} steps, first step with different reset flags: {first_div}")
        for g, (x, y) in enumerate(zip(A, B)):
            report(f"  group {g} (action sigma {SCALES[g]}):")
            report(f"    a: {x}")
            report(f"    b: {y}")
    return bad
