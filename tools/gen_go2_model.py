#!/usr/bin/env python3
"""Extract the Go2 rigid-body model the physics kernel needs from a URDF.

Reads   <reference>/resources/robots/go2/urdf/go2.urdf   (facts: masses, inertias, joint frames,
limits, primitive collision shapes) and writes go2_rl_gym_b200/assets/go2_model.json.

What Isaac Gym does to this asset under the reference's options
(legged_gym/envs/base/legged_robot_config.py:114-134, go2_config.py:148-154) and what we restate:
  * collapse_fixed_joints=True  -> fixed children are merged into their parent for the DYNAMICS
    (13 dynamic bodies: base + 4x{hip,thigh,calf}), but links tagged dont_collapse="true"
    (Head_upper, Head_lower, *_foot; go2.urdf:72,110,369,628,887,1146) are still REPORTED as
    separate rigid bodies (19 reported bodies).  We keep a per-reported-body inertial record so the
    per-env mass randomisation of legged_robot.py:379-402 can be applied body by body and the
    composite is rebuilt afterwards (go2_rl_gym_b200/utils/robot_model.py).
  * replace_cylinder_with_capsule=True -> cylinders are capsules.  Every collider is then sampled
    by spheres (sphere-vs-heightfield is the only narrow phase the kernel has): sphere -> itself,
    capsule -> its two end spheres (one centre sphere when the segment is shorter than the radius),
    box -> inset corner spheres (long thin boxes -> capsule along the long axis).  The sample set is
    pruned to 32 points (one per warp lane); the pruning is stated in COLLIDER_POLICY below.

Usage: python tools/gen_go2_model.py [/root/reference/resources/robots/go2/urdf/go2.urdf]
"""
import json
import math
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

LEGS = ["FL", "FR", "RL", "RR"]
# reported body order = URDF depth-first order (Isaac Gym convention), 19 bodies
REPORT_BODIES = ["base", "Head_upper", "Head_lower"] + [f"{l}_{p}" for l in LEGS for p in ("hip", "thigh", "calf", "foot")]
DYN_BODIES = ["base"] + [f"{l}_{p}" for l in LEGS for p in ("hip", "thigh", "calf")]
DOF_NAMES = [f"{l}_{p}_joint" for l in LEGS for p in ("hip", "thigh", "calf")]


def rpy_to_mat(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def fl(s):
    return [float(x) for x in s.split()]


def parse(path):
    root = ET.parse(path).getroot()
    links, joints = {}, {}
    for L in root.findall("link"):
        d = {"name": L.get("name"), "mass": 0.0, "com": [0, 0, 0], "inertia": [0] * 6, "colliders": []}
        I = L.find("inertial")
        if I is not None:
            o = I.find("origin")
            assert fl(o.get("rpy", "0 0 0")) == [0, 0, 0], "rotated inertial frame not handled"
            d["com"] = fl(o.get("xyz", "0 0 0"))
            d["mass"] = float(I.find("mass").get("value"))
            t = I.find("inertia")
            d["inertia"] = [float(t.get(k)) for k in ("ixx", "iyy", "izz", "ixy", "ixz", "iyz")]
        for C in L.findall("collision"):
            o = C.find("origin")
            xyz = fl(o.get("xyz", "0 0 0")) if o is not None else [0, 0, 0]
            rpy = fl(o.get("rpy", "0 0 0")) if o is not None else [0, 0, 0]
            g = C.find("geometry")[0]
            shape = {"type": g.tag, "xyz": xyz, "rpy": rpy}
            if g.tag == "box":
                shape["size"] = fl(g.get("size"))
            elif g.tag == "cylinder":
                shape["length"], shape["radius"] = float(g.get("length")), float(g.get("radius"))
            elif g.tag == "sphere":
                shape["radius"] = float(g.get("radius"))
            else:
                continue  # meshes are not collision shapes in this asset
            d["colliders"].append(shape)
        links[d["name"]] = d
    for J in root.findall("joint"):
        o = J.find("origin")
        d = {"name": J.get("name"), "type": J.get("type"), "dont_collapse": J.get("dont_collapse") == "true",
             "parent": J.find("parent").get("link"), "child": J.find("child").get("link"),
             "xyz": fl(o.get("xyz", "0 0 0")), "rpy": fl(o.get("rpy", "0 0 0"))}
        if d["type"] == "revolute":
            d["axis"] = fl(J.find("axis").get("xyz"))
            lim = J.find("limit")
            d.update(lower=float(lim.get("lower")), upper=float(lim.get("upper")),
                     effort=float(lim.get("effort")), velocity=float(lim.get("velocity")))
            dyn = J.find("dynamics")
            d["damping"] = float(dyn.get("damping", 0)) if dyn is not None else 0.0
            d["friction"] = float(dyn.get("friction", 0)) if dyn is not None else 0.0
        joints[d["child"]] = d
    return links, joints


def sample_spheres(shape):
    """Sphere samples (centre in shape-parent link frame, radius) of one primitive."""
    R = rpy_to_mat(*shape["rpy"])
    c = np.array(shape["xyz"])
    if shape["type"] == "sphere":
        return [(c, shape["radius"])]
    if shape["type"] == "cylinder":  # capsule, axis = local z
        h, r = shape["length"] / 2, shape["radius"]
        if h < r:
            return [(c, r)]
        ax = R @ np.array([0, 0, 1.0])
        return [(c + h * ax, r), (c - h * ax, r)]
    if shape["type"] == "box":
        he = np.array(shape["size"]) / 2
        order = np.argsort(he)
        if he[order[1]] < 0.02 and he[order[2]] >= 3 * he[order[1]]:  # long thin bar -> capsule along its long axis
            r = float((he[order[0]] + he[order[1]]) / 2)
            ax = np.zeros(3)
            ax[order[2]] = 1.0
            ax = R @ ax
            h = he[order[2]] - r
            return [(c + h * ax, r), (c - h * ax, r)]
        r = float(min(0.02, he.min()))
        out = []
        for sx in (-1, 1):
            for sy in (-1, 1):
                for sz in (-1, 1):
                    out.append((c + R @ (np.array([sx, sy, sz]) * (he - r)), r))
        return out
    raise ValueError(shape["type"])


COLLIDER_POLICY = """32 sphere samples, one per warp lane:
 base box -> 4 bottom inset corners + 2 top centre-line points (front/rear);  Head_upper capsule -> top sphere;
 Head_lower sphere; per leg: hip capsule (1, segment shorter than radius), thigh bar (1: lower sphere of the capsule
 along its long axis), calf capsule knee end + lower end + calflower1 ankle capsule (3, reported as calf), foot sphere (1).
 Dropped: thigh upper sphere (6 cm from the hip sphere, which is 3x larger), calflower (between calf lower end and
 ankle), base top corners (replaced by 2 centre-line points)."""


def build(path):
    links, joints = parse(path)

    # transform of every link relative to its dynamic ancestor (pure fixed-joint chains)
    def to_dyn(name):
        """-> (dyn_body_name, R, p) with x_dyn = R x_link + p"""
        R, p = np.eye(3), np.zeros(3)
        while name not in DYN_BODIES:
            j = joints[name]
            assert j["type"] == "fixed"
            Rj, pj = rpy_to_mat(*j["rpy"]), np.array(j["xyz"])
            R, p = Rj @ R, Rj @ p + pj
            name = j["parent"]
        return name, R, p

    def to_report(name):
        R, p = np.eye(3), np.zeros(3)
        while name not in REPORT_BODIES:
            j = joints[name]
            Rj, pj = rpy_to_mat(*j["rpy"]), np.array(j["xyz"])
            R, p = Rj @ R, Rj @ p + pj
            name = j["parent"]
        return name, R, p

    model = {"source": "resources/robots/go2/urdf/go2.urdf", "dof_names": DOF_NAMES,
             "report_bodies": REPORT_BODIES, "dyn_bodies": DYN_BODIES, "collider_policy": COLLIDER_POLICY}

    # joints
    jl = []
    for n in DOF_NAMES:
        child = n[:-len("_joint")]
        j = joints[child]
        assert j["rpy"] == [0, 0, 0] and j["damping"] == 0 and j["friction"] == 0
        ax = j["axis"]
        assert sorted(ax) == [0, 0, 1]
        jl.append({"name": n, "parent": DYN_BODIES.index(j["parent"]), "origin": j["xyz"], "axis": ax.index(1),
                   "lower": j["lower"], "upper": j["upper"], "effort": j["effort"], "velocity": j["velocity"]})
    model["joints"] = jl

    # reported-body inertial records, expressed in their dynamic ancestor's frame
    rb = []
    for n in REPORT_BODIES:
        L = links[n]
        dyn, R, p = to_dyn(n)
        assert np.allclose(R, np.eye(3)), "rotated dont_collapse link not handled"
        rb.append({"name": n, "dyn": DYN_BODIES.index(dyn), "mass": L["mass"],
                   "com": (np.array(L["com"]) + p).tolist(), "inertia": L["inertia"], "offset": p.tolist()})
    # massless links merged for good must really be massless
    for n, L in links.items():
        if n not in REPORT_BODIES:
            assert L["mass"] == 0.0, n
    model["report_inertials"] = rb

    # colliders
    def pts_of(link):
        out = []
        dyn, R, p = to_dyn(link)
        rep, _, _ = to_report(link)
        for s in links[link]["colliders"]:
            for c, r in sample_spheres(s):
                out.append({"dyn": DYN_BODIES.index(dyn), "report": REPORT_BODIES.index(rep),
                            "pos": (R @ c + p).tolist(), "radius": r, "link": link})
        return out

    cols = []
    base = pts_of("base")
    he = np.array(links["base"]["colliders"][0]["size"]) / 2
    bottom = [c for c in base if c["pos"][2] < 0]
    assert len(bottom) == 4
    cols += bottom
    r = bottom[0]["radius"]
    for sx in (-1, 1):
        cols.append({"dyn": 0, "report": 0, "pos": [sx * (he[0] - r), 0.0, he[2] - r], "radius": r, "link": "base"})
    hu = pts_of("Head_upper")
    assert len(hu) == 1
    cols += hu  # segment shorter than radius -> 1 centre sphere
    cols += pts_of("Head_lower")
    for l in LEGS:
        cols += pts_of(f"{l}_hip")
        thigh = pts_of(f"{l}_thigh")
        cols.append(min(thigh, key=lambda c: c["pos"][2]))     # lower sphere of the thigh bar
        calf = pts_of(f"{l}_calf")
        assert len(calf) == 2
        cols += sorted(calf, key=lambda c: -c["pos"][2])       # knee end, then lower end of the calf capsule
        cols += pts_of(f"{l}_calflower1")                        # ankle
        cols += pts_of(f"{l}_foot")
    assert len(cols) == 32, len(cols)
    model["colliders"] = cols
    return model


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/resources/robots/go2/urdf/go2.urdf"
    m = build(src)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "go2_rl_gym_b200", "assets", "go2_model.json")
    with open(out, "w") as f:
        json.dump(m, f, indent=1)
    print("wrote", os.path.normpath(out), "colliders:", len(m["colliders"]))
    for c in m["colliders"]:
        print(f"  {c['link']:14s} dyn={c['dyn']:2d} rep={c['report']:2d} r={c['radius']:.4f} pos={np.round(c['pos'], 4).tolist()}")
