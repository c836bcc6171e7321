"""Logger (drop-in for legged_gym/utils/logger.py:5-40): the evaluation-side log of play.py — per-step states and the per-episode reward means that
`extras["episode"]` reports, printed as "average rewards per second"."""
from collections import defaultdict

import numpy as np


class Logger:
    def __init__(self, dt):
        self.state_log = defaultdict(list)
        self.rew_log = defaultdict(list)
        self.dt = dt
        self.num_episodes = 0
        self.plot_process = None

    def log_state(self, key, value):
        self.state_log[key].append(value)

    def log_states(self, dict):
        for key, value in dict.items():
            self.log_state(key, value)

    def log_rewards(self, dict, num_episodes):
        for key, value in dict.items():
            if 'rew' in key:
                self.rew_log[key].append(value.item() * num_episodes)
        self.num_episodes += num_episodes

    def reset(self):
        self.state_log.clear()
        self.rew_log.clear()

    def print_rewards(self):
        print("Average rewards per second:")
        for key, values in self.rew_log.items():
            print(f" - {key}: {np.sum(np.array(values)) / self.num_episodes}")
        print(f"Total number of episodes: {self.num_episodes}")
