"""Init-time heightfield for the terrain curriculum (host side, numpy).

Restates `legged_gym/utils/terrain.py:10-197` (grid of num_rows x num_cols sub-terrains, curriculum layout,
env origins, name2cols / cols2id maps) and the `isaacgym.terrain_utils` generators it calls
(terrain.py:114-145).  isaacgym is not in the reference tree: the generators are restated from their public
behaviour and are UNPINNED for the two that draw random numbers; `random_uniform` replaces
the removed `scipy.interpolate.interp2d` by an explicit bilinear upsample (same function, linear kind).
All randomness comes from one `numpy.random.RandomState(seed)` owned by the Terrain object, so a given
(cfg, seed) always produces the same int16 field on every rank.
"""
from collections import defaultdict

import numpy as np

TERRAIN_NAMES = ["wave", "slope", "rough_slope", "stairs_up", "stairs_down", "obstacles", "stepping_stones", "gap", "flat"]


class SubTerrain:
    """80x80 tile of integer heights (units of vertical_scale), axis 0 = x (level direction), axis 1 = y."""

    def __init__(self, terrain_name="terrain", width=256, length=256, vertical_scale=1.0, horizontal_scale=1.0):
        self.terrain_name = terrain_name
        self.terrain_id = -1
        self.vertical_scale = vertical_scale
        self.horizontal_scale = horizontal_scale
        self.width = width
        self.length = length
        self.height_field_raw = np.zeros((width, length), dtype=np.int16)


# ---- generators (isaacgym.terrain_utils restated) -----------------------------------------------------
def wave_terrain(t, num_waves=1, amplitude=1.0):
    amp = int(0.5 * amplitude / t.vertical_scale)
    if num_waves > 0:
        div = t.length / (num_waves * np.pi * 2)
        xs = np.arange(t.width).reshape(-1, 1)
        ys = np.arange(t.length).reshape(1, -1)
        t.height_field_raw += (amp * np.cos(ys / div) + amp * np.sin(xs / div)).astype(np.int16)
    return t


def _bilinear_resize(coarse, n0, n1):
    """Sample `coarse` (defined on linspace(0,1,c0) x linspace(0,1,c1)) at linspace(0,1,n0) x linspace(0,1,n1)."""
    c0, c1 = coarse.shape
    u = np.linspace(0, c0 - 1, n0)
    v = np.linspace(0, c1 - 1, n1)
    i0 = np.clip(np.floor(u).astype(int), 0, c0 - 2)
    j0 = np.clip(np.floor(v).astype(int), 0, c1 - 2)
    fu = (u - i0).reshape(-1, 1)
    fv = (v - j0).reshape(1, -1)
    a = coarse[np.ix_(i0, j0)]
    b = coarse[np.ix_(i0 + 1, j0)]
    c = coarse[np.ix_(i0, j0 + 1)]
    d = coarse[np.ix_(i0 + 1, j0 + 1)]
    return (1 - fu) * (1 - fv) * a + fu * (1 - fv) * b + (1 - fu) * fv * c + fu * fv * d


def random_uniform_terrain(t, min_height, max_height, step=1, downsampled_scale=None, rng=np.random):
    if downsampled_scale is None:
        downsampled_scale = t.horizontal_scale
    lo, hi, st = int(min_height / t.vertical_scale), int(max_height / t.vertical_scale), int(step / t.vertical_scale)
    levels = np.arange(lo, hi + st, st)
    coarse = rng.choice(levels, (int(t.width * t.horizontal_scale / downsampled_scale),
                                 int(t.length * t.horizontal_scale / downsampled_scale))).astype(np.float64)
    t.height_field_raw += np.rint(_bilinear_resize(coarse, t.width, t.length)).astype(np.int16)
    return t


def pyramid_sloped_terrain(t, slope=1, platform_size=1.0):
    cx, cy = int(t.width / 2), int(t.length / 2)
    xs = ((cx - np.abs(cx - np.arange(t.width))) / cx).reshape(-1, 1)
    ys = ((cy - np.abs(cy - np.arange(t.length))) / cy).reshape(1, -1)
    peak = int(slope * (t.horizontal_scale / t.vertical_scale) * (t.width / 2))
    t.height_field_raw += (peak * xs * ys).astype(np.int16)
    half = int(platform_size / t.horizontal_scale / 2)
    x1, y1 = t.width // 2 - half, t.length // 2 - half
    ref = t.height_field_raw[x1, y1]
    t.height_field_raw = np.clip(t.height_field_raw, min(ref, 0), max(ref, 0))
    return t


def pyramid_stairs_terrain(t, step_width, step_height, platform_size=1.0):
    sw, sh = int(step_width / t.horizontal_scale), int(step_height / t.vertical_scale)
    plat = int(platform_size / t.horizontal_scale)
    h, x0, x1, y0, y1 = 0, 0, t.width, 0, t.length
    while (x1 - x0) > plat and (y1 - y0) > plat:
        x0 += sw; x1 -= sw; y0 += sw; y1 -= sw
        h += sh
        t.height_field_raw[x0:x1, y0:y1] = h
    return t


def discrete_obstacles_terrain(t, max_height, min_size, max_size, num_rects, platform_size=1.0, rng=np.random):
    mh = int(max_height / t.vertical_scale)
    smin, smax = int(min_size / t.horizontal_scale), int(max_size / t.horizontal_scale)
    plat = int(platform_size / t.horizontal_scale)
    ni, nj = t.height_field_raw.shape
    heights = [-mh, -mh // 2, mh // 2, mh]
    sizes = list(range(smin, smax, 4))
    for _ in range(num_rects):
        w, l = rng.choice(sizes), rng.choice(sizes)
        si, sj = rng.choice(range(0, ni - w, 4)), rng.choice(range(0, nj - l, 4))
        t.height_field_raw[si:si + w, sj:sj + l] = rng.choice(heights)
    x1, x2 = (t.width - plat) // 2, (t.width + plat) // 2
    y1, y2 = (t.length - plat) // 2, (t.length + plat) // 2
    t.height_field_raw[x1:x2, y1:y2] = 0
    return t


def stepping_stones_terrain(t, stone_size, stone_distance, max_height, platform_size=1.0, depth=-10, rng=np.random):
    ss, sd = int(stone_size / t.horizontal_scale), int(stone_distance / t.horizontal_scale)
    mh, plat = int(max_height / t.vertical_scale), int(platform_size / t.horizontal_scale)
    heights = np.arange(-mh - 1, mh, step=1)
    t.height_field_raw[:, :] = int(depth / t.vertical_scale)
    sx = 0
    if t.length >= t.width:
        sy = 0
        while sy < t.length:
            ey = min(t.length, sy + ss)
            sx = rng.randint(0, ss)
            t.height_field_raw[0:max(0, sx - sd), sy:ey] = rng.choice(heights)
            while sx < t.width:
                ex = min(t.width, sx + ss)
                t.height_field_raw[sx:ex, sy:ey] = rng.choice(heights)
                sx += ss + sd
            sy += ss + sd
    x1, x2 = (t.width - plat) // 2, (t.width + plat) // 2
    y1, y2 = (t.length - plat) // 2, (t.length + plat) // 2
    t.height_field_raw[x1:x2, y1:y2] = 0
    return t


def gap_terrain(t, gap_size, platform_size=1.0):  # terrain.py:176-188
    g, plat = int(gap_size / t.horizontal_scale), int(platform_size / t.horizontal_scale)
    cx, cy = t.length // 2, t.width // 2
    x1 = (t.length - plat) // 2
    y1 = (t.width - plat) // 2
    x2, y2 = x1 + g, y1 + g
    t.height_field_raw[cx - x2:cx + x2, cy - y2:cy + y2] = -1000
    t.height_field_raw[cx - x1:cx + x1, cy - y1:cy + y1] = 0


def pit_terrain(t, depth, platform_size=1.0):  # terrain.py:190-197
    d, half = int(depth / t.vertical_scale), int(platform_size / t.horizontal_scale / 2)
    x1, x2 = t.length // 2 - half, t.length // 2 + half
    y1, y2 = t.width // 2 - half, t.width // 2 + half
    t.height_field_raw[x1:x2, y1:y2] = -d


def convert_heightfield_to_trimesh(height_field_raw, horizontal_scale, vertical_scale, slope_threshold=None):
    """isaacgym.terrain_utils.convert_heightfield_to_trimesh restated (Isaac Gym is absent: UNPINNED, written from its published algorithm): one vertex per
    heightfield sample, two triangles per cell; where the slope between neighbouring samples exceeds slope_threshold the LOWER vertex is moved onto its
    higher neighbour's xy, so that steep ramps become vertical walls (legged_gym/utils/terrain.py:45-49, `slope_treshold`).
    Returns (vertices [rows * cols, 3] float32, triangles [2 (rows - 1) (cols - 1), 3] uint32).  The physics of this package collides against the sampled
    heightfield; the mesh is provided for exporters / viewers, as `Terrain.vertices / .triangles` like the reference."""
    hf = height_field_raw
    num_rows, num_cols = hf.shape
    y = np.linspace(0, (num_cols - 1) * horizontal_scale, num_cols)
    x = np.linspace(0, (num_rows - 1) * horizontal_scale, num_rows)
    yy, xx = np.meshgrid(y, x)
    if slope_threshold is not None:
        thr = slope_threshold * horizontal_scale / vertical_scale
        move_x, move_y, move_c = (np.zeros((num_rows, num_cols)) for _ in range(3))
        move_x[:num_rows - 1, :] += (hf[1:num_rows, :] - hf[:num_rows - 1, :] > thr)
        move_x[1:num_rows, :] -= (hf[:num_rows - 1, :] - hf[1:num_rows, :] > thr)
        move_y[:, :num_cols - 1] += (hf[:, 1:num_cols] - hf[:, :num_cols - 1] > thr)
        move_y[:, 1:num_cols] -= (hf[:, :num_cols - 1] - hf[:, 1:num_cols] > thr)
        move_c[:num_rows - 1, :num_cols - 1] += (hf[1:num_rows, 1:num_cols] - hf[:num_rows - 1, :num_cols - 1] > thr)
        move_c[1:num_rows, 1:num_cols] -= (hf[:num_rows - 1, :num_cols - 1] - hf[1:num_rows, 1:num_cols] > thr)
        xx = xx + (move_x + move_c * (move_x == 0)) * horizontal_scale
        yy = yy + (move_y + move_c * (move_y == 0)) * horizontal_scale
    vertices = np.zeros((num_rows * num_cols, 3), dtype=np.float32)
    vertices[:, 0], vertices[:, 1], vertices[:, 2] = xx.flatten(), yy.flatten(), hf.flatten() * vertical_scale
    triangles = -np.ones((2 * (num_rows - 1) * (num_cols - 1), 3), dtype=np.uint32)
    for i in range(num_rows - 1):
        ind0 = np.arange(0, num_cols - 1) + i * num_cols
        ind1, ind2, ind3 = ind0 + 1, ind0 + num_cols, ind0 + num_cols + 1
        start, stop = 2 * i * (num_cols - 1), 2 * i * (num_cols - 1) + 2 * (num_cols - 1)
        triangles[start:stop:2, 0], triangles[start:stop:2, 1], triangles[start:stop:2, 2] = ind0, ind3, ind1
        triangles[start + 1:stop:2, 0], triangles[start + 1:stop:2, 1], triangles[start + 1:stop:2, 2] = ind0, ind2, ind3
    return vertices, triangles


# ---- the grid -------------------------------------------------------------------------------------------
class Terrain:
    def __init__(self, cfg, num_robots, seed=0):
        self.cfg = cfg
        self.num_robots = num_robots
        self.type = cfg.mesh_type
        if self.type in ["none", "plane"]:
            return
        self.rng = np.random.RandomState(seed)
        self.env_length, self.env_width = cfg.terrain_length, cfg.terrain_width
        self.proportions = [float(np.sum(cfg.terrain_proportions[:i + 1])) for i in range(len(cfg.terrain_proportions))]
        cfg.num_sub_terrains = cfg.num_rows * cfg.num_cols
        self.env_origins = np.zeros((cfg.num_rows, cfg.num_cols, 3))
        self.width_per_env_pixels = int(self.env_width / cfg.horizontal_scale)
        self.length_per_env_pixels = int(self.env_length / cfg.horizontal_scale)
        self.spacing = cfg.terrain_spacing
        self.spacing_pixels = int(self.spacing / cfg.horizontal_scale)
        self.border = int(cfg.border_size / cfg.horizontal_scale)
        self.tot_cols = int(cfg.num_cols * self.width_per_env_pixels + max(0, cfg.num_cols - 1) * self.spacing_pixels) + 2 * self.border
        self.tot_rows = int(cfg.num_rows * self.length_per_env_pixels + max(0, cfg.num_rows - 1) * self.spacing_pixels) + 2 * self.border
        self.name2cols = defaultdict(set)
        self.cols2id = []
        self.height_field_raw = np.zeros((self.tot_rows, self.tot_cols), dtype=np.int16)
        if cfg.curriculum:
            self.curiculum()
        elif cfg.selected:
            raise NotImplementedError("terrain.selected is outside the hot-path scope (SURVEY 8f)")
        else:
            self.randomized_terrain()
        self.heightsamples = self.height_field_raw
        self._mesh = None

    def _trimesh(self):      # terrain.py:45-49; built on first use (3 M vertices at the GO2 grid): the simulation itself never reads it
        if self._mesh is None:
            self._mesh = convert_heightfield_to_trimesh(self.height_field_raw, self.cfg.horizontal_scale, self.cfg.vertical_scale, self.cfg.slope_treshold)
        return self._mesh

    @property
    def vertices(self):
        if self.type != "trimesh":
            raise AttributeError("Terrain.vertices exists for mesh_type == 'trimesh' only (terrain.py:44-49)")
        return self._trimesh()[0]

    @property
    def triangles(self):
        if self.type != "trimesh":
            raise AttributeError("Terrain.triangles exists for mesh_type == 'trimesh' only (terrain.py:44-49)")
        return self._trimesh()[1]

    def randomized_terrain(self):  # terrain.py:51-59
        for k in range(self.cfg.num_sub_terrains):
            i, j = np.unravel_index(k, (self.cfg.num_rows, self.cfg.num_cols))
            choice = self.rng.uniform(0, 1)
            difficulty = self.rng.choice([0.5, 0.75, 0.9])
            self.add_terrain_to_map(self.make_terrain(choice, difficulty), i, j)

    def curiculum(self):  # terrain.py:61-70 (sic)
        for j in range(self.cfg.num_cols):
            tile = None
            for i in range(self.cfg.num_rows):
                tile = self.make_terrain(j / self.cfg.num_cols + 0.001, i / self.cfg.num_rows)
                self.add_terrain_to_map(tile, i, j)
            self.name2cols[tile.terrain_name].add(j)
            self.cols2id.append(tile.terrain_id)

    def make_terrain(self, choice, difficulty):  # terrain.py:87-155 (IS_HARD branch)
        t = SubTerrain("terrain", width=self.width_per_env_pixels, length=self.width_per_env_pixels,
                       vertical_scale=self.cfg.vertical_scale, horizontal_scale=self.cfg.horizontal_scale)
        slope = 0.1 + difficulty * 0.52
        step_height = 0.05 + 0.23 * difficulty
        obstacle_height = 0.05 + difficulty * 0.25
        stone_size = 1.5 * (1.05 - difficulty)
        stone_distance = 0.05 if difficulty == 0 else 0.1
        gap_size = 1.0 * difficulty
        amplitude = 0.1 + 0.2 * difficulty
        p = self.proportions
        rough = dict(min_height=-0.05, max_height=0.05, step=0.005, downsampled_scale=0.2, rng=self.rng)
        if choice < p[0]:
            tid = 0
            wave_terrain(t, num_waves=5, amplitude=amplitude)
            random_uniform_terrain(t, **rough)
        elif choice < p[1]:
            tid = 1
            if choice < (p[0] + p[1]) / 2:
                slope *= -1
            pyramid_sloped_terrain(t, slope=slope, platform_size=3.0)
        elif choice < p[2]:
            tid = 2
            pyramid_sloped_terrain(t, slope=slope, platform_size=3.0)
            random_uniform_terrain(t, **rough)
        elif choice < p[4]:
            tid = 4
            if choice < p[3]:
                tid = 3
                step_height *= -1
            pyramid_stairs_terrain(t, step_width=0.31, step_height=step_height, platform_size=3.0)
        elif choice < p[5]:
            tid = 5
            discrete_obstacles_terrain(t, obstacle_height, 1.0, 2.0, 20, platform_size=3.0, rng=self.rng)
        elif choice < p[6]:
            tid = 6
            stepping_stones_terrain(t, stone_size=stone_size, stone_distance=stone_distance, max_height=0.0,
                                    platform_size=4.0, rng=self.rng)
        elif choice < p[7]:
            tid = 7
            gap_terrain(t, gap_size=gap_size, platform_size=3.0)
        else:
            tid = 8
            pit_terrain(t, depth=0.0, platform_size=4.0)
        t.terrain_id, t.terrain_name = tid, TERRAIN_NAMES[tid]
        return t

    def add_terrain_to_map(self, t, row, col):  # terrain.py:157-174
        sx = self.border + row * (self.length_per_env_pixels + self.spacing_pixels)
        sy = self.border + col * (self.width_per_env_pixels + self.spacing_pixels)
        self.height_field_raw[sx:sx + self.length_per_env_pixels, sy:sy + self.width_per_env_pixels] = t.height_field_raw
        ox = (row + 0.5) * self.env_length + row * self.spacing
        oy = (col + 0.5) * self.env_width + col * self.spacing
        x1, x2 = int((self.env_length / 2.0 - 1) / t.horizontal_scale), int((self.env_length / 2.0 + 1) / t.horizontal_scale)
        y1, y2 = int((self.env_width / 2.0 - 1) / t.horizontal_scale), int((self.env_width / 2.0 + 1) / t.horizontal_scale)
        oz = np.max(t.height_field_raw[x1:x2, y1:y2]) * t.vertical_scale
        self.env_origins[row, col] = [ox, oy, oz]
