"""Go2 rigid-body model on the host: constants for the kernels and per-env randomised inertials.

Model facts come from go2.urdf through tools/gen_go2_model.py -> assets/go2_model.json.
Per-env mass / COM randomisation restates LeggedRobot._process_rigid_body_props
(legged_gym/envs/base/legged_robot.py:379-402): base mass += U[added_mass_range], every other REPORTED body
mass *= U[multiplied_link_mass_range], base COM += U[added_base_com_range]^3; with recomputeInertia=True
(legged_robot.py:1034) we scale each body's inertia tensor with its mass ratio, then merge the reported bodies
into the 13 dynamic bodies (parallel-axis theorem).
"""
import json
import os

import numpy as np

from .. import _abi

_ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "go2_model.json")


def load_model_json(path=None):
    with open(path or _ASSET) as f:
        return json.load(f)


def build_model_struct(m=None):
    m = m or load_model_json()
    s = _abi.Go2Model()
    for j, jn in enumerate(m["joints"]):
        for k in range(3):
            s.joint_origin[j][k] = jn["origin"][k]
        s.joint_axis[j] = jn["axis"]
        s.q_lower[j], s.q_upper[j] = jn["lower"], jn["upper"]
        s.effort[j], s.vel_limit[j] = jn["effort"], jn["velocity"]
    assert len(m["colliders"]) == _abi.NUM_COL
    for c, col in enumerate(m["colliders"]):
        for k in range(3):
            s.col_pos[c][k] = col["pos"][k]
        s.col_radius[c] = col["radius"]
        s.col_dyn[c] = col["dyn"]
        s.col_report[c] = col["report"]
    feet = [r for r in m["report_inertials"] if r["name"].endswith("_foot")]
    for l, r in enumerate(feet):
        for k in range(3):
            s.foot_offset[l][k] = r["offset"][k]
    return s


def _inertia_mat(v6):
    ixx, iyy, izz, ixy, ixz, iyz = v6
    return np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]], dtype=np.float64)


def composite_inertials(m, num_envs, base_added_mass=None, link_mass_ratio=None, base_added_com=None):
    """-> float32 [num_envs, 13, 10]: mass, com xyz, Ixx Iyy Izz Ixy Ixz Iyz (about the composite COM, body frame).

    base_added_mass [N], link_mass_ratio [N,18] (reported bodies 1..18), base_added_com [N,3]; None = nominal."""
    rep = m["report_inertials"]
    N = num_envs
    out = np.zeros((N, _abi.NUM_DYN, _abi.INERTIA_STRIDE), dtype=np.float64)
    mass = np.zeros((N, len(rep)))
    com = np.zeros((N, len(rep), 3))
    iner = np.zeros((N, len(rep), 3, 3))
    for r, body in enumerate(rep):
        m0 = body["mass"]
        ratio = np.ones(N)
        if r == 0:
            mr = m0 + (base_added_mass if base_added_mass is not None else 0.0)
            ratio = mr / m0
            mass[:, r] = mr
            com[:, r] = np.asarray(body["com"]) + (base_added_com if base_added_com is not None else 0.0)
        else:
            if link_mass_ratio is not None:
                ratio = link_mass_ratio[:, r - 1]
            mass[:, r] = m0 * ratio
            com[:, r] = np.asarray(body["com"])
        iner[:, r] = _inertia_mat(body["inertia"])[None] * np.reshape(ratio, (-1, 1, 1))
    for d in range(_abi.NUM_DYN):
        members = [r for r, body in enumerate(rep) if body["dyn"] == d]
        mt = mass[:, members].sum(1)
        ct = (mass[:, members, None] * com[:, members]).sum(1) / mt[:, None]
        It = np.zeros((N, 3, 3))
        for r in members:
            dvec = com[:, r] - ct
            d2 = (dvec * dvec).sum(1)
            It += iner[:, r] + mass[:, r, None, None] * (d2[:, None, None] * np.eye(3)[None] - dvec[:, :, None] * dvec[:, None, :])
        out[:, d, 0] = mt
        out[:, d, 1:4] = ct
        out[:, d, 4], out[:, d, 5], out[:, d, 6] = It[:, 0, 0], It[:, 1, 1], It[:, 2, 2]
        out[:, d, 7], out[:, d, 8], out[:, d, 9] = It[:, 0, 1], It[:, 0, 2], It[:, 1, 2]
    return out.astype(np.float32)
