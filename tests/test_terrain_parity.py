"""Terrain layout parity: the REFERENCE's `legged_gym.utils.terrain.Terrain` (imported from /root/reference; its
isaacgym.terrain_utils generators are absent from the tree and are served by the stand-in in tests/ref_stub, i.e. by this package's restated
generators) and this package's Terrain must build the same 10 x 20 curriculum heightfield, env origins and column -> terrain-id map from
the same numpy seed.  That pins everything terrain.py itself does (difficulty / proportion logic, tile placement, borders, origin heights);
the generators' own formulas stay 'as documented for Isaac Gym' (unpinned).  Build container only, in a subprocess."""
import json
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, json, io, contextlib
import numpy as np
sys.path[:0] = [ROOT + "/tests/ref_stub", REF, REF + "/rsl_rl", ROOT]
with contextlib.redirect_stdout(io.StringIO()):
    from legged_gym.envs.go2.go2_config import GO2Cfg as RefCfg          # (envs first: the reference's utils <-> envs import cycle)
    from legged_gym.utils.terrain import Terrain as RefTerrain
    import importlib
    mine_t = importlib.import_module("go2_rl_gym_b200.utils.terrain")
    MyCfg = importlib.import_module("go2_rl_gym_b200.envs.go2.go2_config").GO2Cfg
out = {}
for mode in ("curriculum", "random"):
    rc, mc = RefCfg().terrain, MyCfg().terrain
    for c in (rc, mc):
        c.mesh_type = "heightfield"
        c.curriculum = mode == "curriculum"
        if mode == "random":
            c.num_rows, c.num_cols = 4, 5
    np.random.seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        r = RefTerrain(rc, 64)
    m = mine_t.Terrain(mc, 64, seed=7)
    out[mode] = {"shape": [list(r.height_field_raw.shape), list(m.height_field_raw.shape)],
                 "hf_equal": bool(np.array_equal(r.height_field_raw, m.height_field_raw)),
                 "hf_maxdiff": int(np.abs(r.height_field_raw.astype(np.int64) - m.height_field_raw.astype(np.int64)).max()),
                 "origins_equal": bool(np.array_equal(np.asarray(r.env_origins), np.asarray(m.env_origins))),
                 "nonflat": int((m.height_field_raw != 0).sum())}
    if mode == "curriculum":
        out[mode]["cols2id"] = [list(map(int, getattr(r, "cols2id", []))), list(map(int, m.cols2id))]
        out[mode]["name2cols"] = [{k: sorted(map(int, v)) for k, v in getattr(r, "name2cols", {}).items()}, {k: sorted(map(int, v)) for k, v in m.name2cols.items()}]
print("RESULT" + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "legged_gym")), reason="needs the reference tree (runs in the build container)")
def test_terrain_grid_equals_the_reference():
    code = SCRIPT.replace("ROOT", repr(ROOT)).replace("REF", repr(REF))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=600)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    assert line, out.stderr[-3000:]
    res = json.loads(line[0][len("RESULT"):])
    for mode, r in res.items():
        assert r["shape"][0] == r["shape"][1], (mode, r["shape"])
        assert r["hf_equal"] and r["origins_equal"], (mode, r)
        assert r["nonflat"] > 10000, (mode, r["nonflat"])
    c = res["curriculum"]
    assert c["cols2id"][0] == c["cols2id"][1] and c["name2cols"][0] == c["name2cols"][1], c
    assert c["shape"][1] == [1345, 2195]          # 10 x 20 tiles of 8 m at 0.1 m, 0.5 m spacing, 25 m border


@pytest.mark.parametrize("N", [32, 48, 64, 4096, 8192, 16384, 65536, 100])
def test_round_robin_terrain_assignment_uses_torch_float32_semantics(N):
    """legged_robot.py:1071-1072: levels = i % 6, types = floor(i / (N / 20)) evaluated by torch in FLOAT32 — at i = k N / 4 the quotient lands just
    below the integer (env 1024 of 4096 is on column 4, not 5).  EnvArrays must reproduce that, for whole runs and for shards (global env ids)."""
    import torch
    from go2_rl_gym_b200.envs.env_arrays import EnvArrays
    from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
    ref_types = torch.div(torch.arange(N), (N / 20), rounding_mode="floor").to(torch.long)
    ref_levels = torch.fmod(torch.arange(N), 5 + 1)
    if N <= 8192:
        cfg = GO2Cfg(); cfg.env.num_envs = N; cfg.terrain.mesh_type = "heightfield"
        A = EnvArrays(cfg, "cpu", seed=1)
        assert torch.equal(A.tensors["terrain_types"].long(), ref_types) and torch.equal(A.tensors["terrain_levels"].long(), ref_levels)
    # a shard keeps the global assignment
    Ns = min(N // 4, 2048) if N % 4 == 0 else None
    if Ns:
        cfg = GO2Cfg(); cfg.env.num_envs = Ns; cfg.terrain.mesh_type = "heightfield"
        off = N // 4
        A = EnvArrays(cfg, "cpu", num_envs=Ns, env_offset=off, num_envs_global=N, seed=1)
        assert torch.equal(A.tensors["terrain_types"].long(), ref_types[off:off + Ns])
        assert int(A.tensors["terrain_types"][0]) == int(ref_types[off]) == (5 if N == 100 else 4)   # the float32 boundary case itself (100 / 20 is exact)


def test_trimesh_conversion_makes_steep_edges_vertical():
    """convert_heightfield_to_trimesh (restated from isaacgym.terrain_utils, unpinned): vertex / triangle counts, heights untouched, and the wall
    correction of terrain.py:45-49 — across every cell edge steeper than slope_treshold the lower vertex sits on its higher neighbour's xy."""
    import numpy as np
    from go2_rl_gym_b200.utils.terrain import convert_heightfield_to_trimesh
    hf = np.zeros((12, 9), dtype=np.int16)
    hf[5:, :] = 40                      # a 0.2 m step along x at vertical_scale 0.005: slope 2 > 0.75
    hf[:, 6:] += 6                      # a 0.03 m step along y: slope 0.3, left alone
    v, t = convert_heightfield_to_trimesh(hf, 0.1, 0.005, 0.75)
    assert v.shape == (12 * 9, 3) and t.shape == (2 * 11 * 8, 3) and t.max() == 12 * 9 - 1 and v.dtype == np.float32
    V = v.reshape(12, 9, 3)
    assert np.allclose(V[..., 2], hf * 0.005)
    assert np.allclose(V[4, :, 0], V[5, :, 0]) and np.allclose(V[5, :, 0], 0.5)        # row 4 (low side) moved onto row 5's x
    assert np.allclose(V[3, :, 0], 0.3) and np.allclose(V[6, :, 0], 0.6)                # the other rows stay on the grid
    keep = [r for r in range(12) if r != 4]
    assert np.allclose(V[keep, :, 1], np.arange(9) * 0.1)                               # the shallow step along y moves nothing (row 4 also takes the corner rule)
    v0, _ = convert_heightfield_to_trimesh(hf, 0.1, 0.005, None)
    assert np.allclose(v0.reshape(12, 9, 3)[..., 0], (np.arange(12) * 0.1)[:, None])
    # Terrain exposes the mesh for mesh_type == 'trimesh' (terrain.py:44-49) and only then
    from go2_rl_gym_b200.envs.go2.go2_config import GO2Cfg
    from go2_rl_gym_b200.utils.terrain import Terrain
    cfg = GO2Cfg().terrain
    cfg.num_rows, cfg.num_cols, cfg.mesh_type = 2, 2, "trimesh"
    T = Terrain(cfg, 16, seed=1)
    assert T.vertices.shape[0] == T.tot_rows * T.tot_cols and T.triangles.shape[0] == 2 * (T.tot_rows - 1) * (T.tot_cols - 1)
    cfg.mesh_type = "heightfield"
    with pytest.raises(AttributeError):
        Terrain(cfg, 16, seed=1).vertices
