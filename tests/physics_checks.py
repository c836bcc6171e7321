"""Independent numpy forward kinematics / energy bookkeeping used to sanity-check the physics oracle."""
import numpy as np


def rot_axis(axis, q):
    c, s = np.cos(q), np.sin(q)
    if axis == 0:
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    if axis == 1:
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def quat_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def mechanics(model_json, inertia, root, q, qd, g=9.81):
    """-> (total mass, COM world [3], kinetic energy, potential energy, linear momentum [3], angular momentum about origin [3])
    computed with world-frame classical mechanics (no spatial algebra)."""
    joints = model_json["joints"]
    R = [None] * 13; p = [None] * 13; w = [None] * 13; v = [None] * 13
    R[0] = quat_to_mat(root[3:7]); p[0] = root[0:3].astype(np.float64)
    v[0] = root[7:10].astype(np.float64); w[0] = root[10:13].astype(np.float64)
    for b in range(1, 13):
        j = joints[b - 1]; par = j["parent"]
        r = R[par] @ np.array(j["origin"])
        p[b] = p[par] + r
        R[b] = R[par] @ rot_axis(j["axis"], q[b - 1])
        axis_w = R[b][:, j["axis"]]
        w[b] = w[par] + axis_w * qd[b - 1]
        v[b] = v[par] + np.cross(w[par], r)
    M = 0.0; com = np.zeros(3); KE = 0.0; PE = 0.0; P = np.zeros(3); L = np.zeros(3)
    for b in range(13):
        m = float(inertia[b, 0]); c = R[b] @ inertia[b, 1:4].astype(np.float64)
        ixx, iyy, izz, ixy, ixz, iyz = inertia[b, 4:10].astype(np.float64)
        Ib = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
        Iw = R[b] @ Ib @ R[b].T
        pc = p[b] + c; vc = v[b] + np.cross(w[b], c)
        M += m; com += m * pc
        KE += 0.5 * m * vc @ vc + 0.5 * w[b] @ Iw @ w[b]
        PE += m * g * pc[2]
        P += m * vc; L += np.cross(pc, m * vc) + Iw @ w[b]
    return M, com / M, KE, PE, P, L
