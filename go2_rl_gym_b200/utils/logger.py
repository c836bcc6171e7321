"""Evaluation-side bookkeeping with the interface of the reference's `legged_gym.utils.Logger` (legged_gym/utils/logger.py:5-40), which
play-style scripts use: per-step state traces, and the per-episode reward means of `extras["episode"]` turned into "average rewards per second".

Written for this package (the reference's plotting process is not part of it: there is no viewer / display on the training boxes)."""
import numpy as np


class Logger:
    """state_log[key] = list of logged values; rew_log[key] = list of (episode mean x episodes in that report); num_episodes = episodes seen."""

    def __init__(self, dt):
        self.dt = dt
        self.state_log, self.rew_log = {}, {}
        self.num_episodes = 0
        self.plot_process = None          # attribute kept for scripts that test it

    # ---- state traces
    def log_state(self, key, value):
        self.state_log.setdefault(key, []).append(value)

    def log_states(self, dict):
        for key in dict:
            self.log_state(key, dict[key])

    # ---- episode rewards: `dict` is extras["episode"]; only its reward entries ('rew_*') count, weighted by the episodes they average over
    def log_rewards(self, dict, num_episodes):
        for key in dict:
            if 'rew' in key:
                self.rew_log.setdefault(key, []).append(dict[key].item() * num_episodes)
        self.num_episodes += num_episodes

    def mean_rewards(self):
        """{reward name: mean over all reported episodes} (what print_rewards prints)."""
        return {key: float(np.sum(np.asarray(vals))) / self.num_episodes for key, vals in self.rew_log.items()}

    def print_rewards(self):
        print("Average rewards per second:")
        for key, vals in self.rew_log.items():
            print(f" - {key}: {np.sum(np.array(vals)) / self.num_episodes}")
        print(f"Total number of episodes: {self.num_episodes}")

    def reset(self):
        self.state_log.clear()
        self.rew_log.clear()
