"""isaacgym.terrain_utils stand-in: delegates to the build's own restated generators with the GLOBAL numpy RNG
(the original draws from np.random), so the reference's Terrain class can be executed for layout parity."""
import numpy as np
from go2_rl_gym_b200.utils import terrain as _t

SubTerrain = _t.SubTerrain
wave_terrain = _t.wave_terrain
pyramid_sloped_terrain = _t.pyramid_sloped_terrain
pyramid_stairs_terrain = _t.pyramid_stairs_terrain


def random_uniform_terrain(terrain, min_height, max_height, step=1, downsampled_scale=None):
    return _t.random_uniform_terrain(terrain, min_height, max_height, step, downsampled_scale, rng=np.random)


def discrete_obstacles_terrain(terrain, max_height, min_size, max_size, num_rects, platform_size=1.):
    return _t.discrete_obstacles_terrain(terrain, max_height, min_size, max_size, num_rects, platform_size, rng=np.random)


def stepping_stones_terrain(terrain, stone_size, stone_distance, max_height, platform_size=1., depth=-10):
    return _t.stepping_stones_terrain(terrain, stone_size, stone_distance, max_height, platform_size, depth, rng=np.random)


def convert_heightfield_to_trimesh(*a, **k):
    raise NotImplementedError
