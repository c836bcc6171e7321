"""Host-side torch helpers of legged_gym/utils/math.py:7-27 (the env's own use of them — yaw-only rotation of the height-scan grid, the heading
error — lives inside the fused step kernel; these are for user code written against the reference)."""
import numpy as np
import torch


def _normalize(x, eps=1e-9):
    return x / x.norm(p=2, dim=-1).clamp(min=eps, max=None).unsqueeze(-1)


def _quat_apply(a, b):       # isaacgym.torch_utils.quat_apply, xyzw
    shape = b.shape
    a, b = a.reshape(-1, 4), b.reshape(-1, 3)
    xyz = a[:, :3]
    t = xyz.cross(b, dim=-1) * 2
    return (b + a[:, 3:] * t + xyz.cross(t, dim=-1)).view(shape)


def quat_apply_yaw(quat, vec):
    """Rotate vec by the yaw component of quat only (math.py:8-12)."""
    quat_yaw = quat.clone().view(-1, 4)
    quat_yaw[:, :2] = 0.
    return _quat_apply(_normalize(quat_yaw), vec)


def wrap_to_pi(angles):
    """In place, like the reference (math.py:15-18)."""
    angles %= 2 * np.pi
    angles -= 2 * np.pi * (angles > np.pi)
    return angles


def torch_rand_sqrt_float(lower, upper, shape, device):
    """math.py:21-27: square-root-shaped density around the middle of [lower, upper]."""
    r = 2 * torch.rand(*shape, device=device) - 1
    r = torch.where(r < 0., -torch.sqrt(-r), torch.sqrt(r))
    r = (r + 1.) / 2.
    return (upper - lower) * r + lower
