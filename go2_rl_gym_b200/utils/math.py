"""Host-side torch helpers with the names and semantics of legged_gym/utils/math.py:7-27, for user code written against the reference.  The env's
own use of the same arithmetic (yaw-only rotation of the height-scan grid, the heading error) lives inside the fused step kernel."""
import math as _math

import torch

_TWO_PI = 2.0 * _math.pi


def quat_apply_yaw(quat, vec):
    """Rotate `vec` [..., 3] by the YAW part of `quat` [..., 4] (xyzw) only: drop x / y, renormalise (z, w), rotate about the vertical axis.
    Closed form of that rotation: cos(yaw) = w^2 - z^2, sin(yaw) = 2 w z on the unit (z, w)."""
    q = quat.reshape(-1, 4)
    v = vec.reshape(-1, 3)
    z, w = q[:, 2], q[:, 3]
    n = torch.sqrt(z * z + w * w).clamp(min=1e-9)
    z, w = z / n, w / n
    c, s = w * w - z * z, 2.0 * w * z
    out = torch.stack((c * v[:, 0] - s * v[:, 1], s * v[:, 0] + c * v[:, 1], v[:, 2]), dim=-1)
    return out.view(vec.shape)


def wrap_to_pi(angles):
    """Map angles to (-pi, pi], IN PLACE like the reference (its callers rely on that), and return the same tensor."""
    angles.copy_(torch.remainder(angles, _TWO_PI))
    angles.sub_(_TWO_PI * (angles > _math.pi))
    return angles


def torch_rand_sqrt_float(lower, upper, shape, device):
    """Uniform draw pushed through a signed square root: samples in [lower, upper] with a density that thins out around the middle."""
    u = 2.0 * torch.rand(*shape, device=device) - 1.0
    r = torch.sign(u) * torch.sqrt(u.abs())
    return (upper - lower) * ((r + 1.0) / 2.0) + lower
