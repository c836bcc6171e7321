"""Import shim: `legged_gym.*` resolves to the B200-native package so the reference's entry points run unchanged
(legged_gym/scripts/train.py:1-19 imports `legged_gym.envs` and `legged_gym.utils`)."""
import os

LEGGED_GYM_ROOT_DIR = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
LEGGED_GYM_ENVS_DIR = os.path.join(LEGGED_GYM_ROOT_DIR, 'legged_gym', 'envs')
