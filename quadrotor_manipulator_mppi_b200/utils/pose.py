"""Target-pose holder with the attribute surface of the reference's `utils/pose.py:4-112`
(`.pose`, `.orientation` xyzw, `.clone()`, `.tf_matrix()`, `.rpy`, `.x/.y/.z`), kept so callers
can keep passing / mutating `mppi.target_pose`.  Host-side only: the solver reads the seven
numbers and uploads them with `mppi_set_target`.
"""
from __future__ import annotations

import math

import torch


def _quat_to_matrix_xyzw(q: torch.Tensor) -> torch.Tensor:
    # quaternion_to_matrix patched to xyzw (utils/rotation_conversions.py:45-75)
    i, j, k, r = (float(v) for v in q)
    s = 2.0 / (i * i + j * j + k * k + r * r)
    return torch.tensor([[1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r)],
                         [s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r)],
                         [s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j)]], dtype=torch.float32)


class Pose:
    def __init__(self):
        self._pose = torch.zeros(3)
        self._orientation = torch.tensor([0.0, 0.0, 0.0, 1.0])

    @property
    def pose(self) -> torch.Tensor:
        return self._pose

    @pose.setter
    def pose(self, pos):
        self._pose = pos if isinstance(pos, torch.Tensor) else torch.tensor([pos.x, pos.y, pos.z])

    @property
    def orientation(self) -> torch.Tensor:
        return self._orientation

    @orientation.setter
    def orientation(self, ori):
        self._orientation = ori if isinstance(ori, torch.Tensor) else torch.tensor([ori.x, ori.y, ori.z, ori.w])

    @property
    def tf_return(self) -> torch.Tensor:
        return _quat_to_matrix_xyzw(self._orientation)

    @property
    def rpy(self) -> torch.Tensor:
        m = self.tf_return
        # matrix_to_euler_angles(M, "ZYX") (utils/rotation_conversions.py:277-319)
        return torch.tensor([math.atan2(m[1, 0], m[0, 0]), math.asin(max(-1.0, min(1.0, -float(m[2, 0])))),
                             math.atan2(m[2, 1], m[2, 2])])

    @property
    def np_pose(self):
        return self._pose.detach().cpu().numpy()

    @property
    def np_orientation(self):
        return self._orientation.detach().cpu().numpy()

    @property
    def x(self):
        return self._pose[0].item()

    @x.setter
    def x(self, v):
        self._pose[0] = v

    @property
    def y(self):
        return self._pose[1].item()

    @y.setter
    def y(self, v):
        self._pose[1] = v

    @property
    def z(self):
        return self._pose[2].item()

    @z.setter
    def z(self, v):
        self._pose[2] = v

    def tf_matrix(self, device=None) -> torch.Tensor:
        m = torch.eye(4, device=device)
        m[:3, :3] = self.tf_return.to(m.device)
        m[:3, 3] = self._pose.to(m.device)
        return m

    def clone(self) -> "Pose":
        p = Pose()
        p.pose = self._pose.clone()
        p.orientation = self._orientation.clone()
        return p

    def version_key(self):
        """(tensor, in-place version) pairs; compare with `TensorWatch`, which holds the tensors it last saw so that
        CPython cannot hand a NEW tensor the address (and version 0) of the one that was uploaded."""
        return (self._pose, self._pose._version, self._orientation, self._orientation._version)

    def as_floats(self):
        """(x, y, z, qx, qy, qz, qw) as Python floats -- what the solver uploads."""
        return tuple(float(v) for v in self._pose.detach().cpu().reshape(-1)) + \
            tuple(float(v) for v in self._orientation.detach().cpu().reshape(-1))


class TensorWatch:
    """Change detector for attributes the caller may re-assign or edit in place between steps (`target_pose.pose`,
    `drone.target`): identity + torch's in-place version counter, with a STRONG reference to the tensor last seen --
    an `id()`-based key can repeat when a freed tensor's address is reused, and the solver would then skip an upload
    the reference (which re-reads its target every step, mppi.py:137, drone_mppi.py:141) never misses."""

    def __init__(self):
        self._seen = None

    def changed(self, *values) -> bool:
        cur = []
        for v in values:
            cur.append((v, v._version) if isinstance(v, torch.Tensor) else (None, tuple(float(x) for x in v)))
        prev, self._seen = self._seen, cur
        if prev is None or len(prev) != len(cur):
            return True
        for (a, av), (b, bv) in zip(prev, cur):
            if a is not b or av != bv:
                return True
        return False
