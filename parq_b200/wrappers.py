"""Host-side input types of the decoder boundary: ``Pose`` and ``Camera``.

These mirror the tensor-wrapper interface of the reference
(/root/reference/utils/wrappers.py:114-294 ``TensorWrapper``/``Pose`` and
:441-553 ``Camera``) closely enough that the decoder accepts either the
reference's own objects or these (duck typing on ``._data``):

* ``Pose._data``  (..., 12) = rotation row-major (9) followed by translation (3)
* ``Camera._data`` (..., 6) = [w, h, fx, fy, cx, cy]   (pinhole only; the
  reference asserts exactly 6 parameters, wrappers.py:444-446 -- there is no
  distortion model to reproduce)
* ``.shape`` hides the trailing parameter dimension (wrappers.py:121-123).

They are plumbing: the arithmetic that matters for parity lives in the CUDA
kernels (``parq_pose_chain`` / ``parq_project_sample``), not here.
"""
from typing import Tuple

import torch


class _Wrapped:
    _width = None

    def __init__(self, data: torch.Tensor):
        if not isinstance(data, torch.Tensor):
            data = torch.as_tensor(data)
        if data.shape[-1] != self._width:
            raise ValueError("%s expects last dim %d, got %s" % (type(self).__name__, self._width, tuple(data.shape)))
        self._data = data

    @property
    def shape(self):
        return self._data.shape[:-1]

    @property
    def device(self):
        return self._data.device

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def data(self):
        return self._data

    def __getitem__(self, index):
        return self.__class__(self._data[index])

    def to(self, *args, **kwargs):
        return self.__class__(self._data.to(*args, **kwargs))

    def cuda(self):
        return self.__class__(self._data.cuda())

    def cpu(self):
        return self.__class__(self._data.cpu())

    def float(self):
        return self.__class__(self._data.float())

    def pin_memory(self):
        return self.__class__(self._data.pin_memory())

    def clone(self):
        return self.__class__(self._data.clone())

    def unsqueeze(self, dim):
        assert dim != -1 and dim != self._data.dim()
        return self.__class__(self._data.unsqueeze(dim))

    def __repr__(self):
        return "%s %s %s %s" % (type(self).__name__, tuple(self.shape), self.dtype, self.device)


class Pose(_Wrapped):
    """SE(3) pose, (..., 12) = R row-major | t  (reference wrappers.py:194-294)."""
    _width = 12

    @classmethod
    def from_Rt(cls, R: torch.Tensor, t: torch.Tensor) -> "Pose":
        assert R.shape[-2:] == (3, 3) and t.shape[-1] == 3 and R.shape[:-2] == t.shape[:-1]
        return cls(torch.cat([R.flatten(start_dim=-2), t], -1))

    @property
    def R(self) -> torch.Tensor:
        r = self._data[..., :9]
        return r.reshape(r.shape[:-1] + (3, 3))

    @property
    def t(self) -> torch.Tensor:
        return self._data[..., -3:]

    def inverse(self) -> "Pose":
        R = self.R.transpose(-1, -2)
        t = -(R @ self.t.unsqueeze(-1)).squeeze(-1)
        return Pose.from_Rt(R, t)

    def compose(self, other: "Pose") -> "Pose":
        R = self.R @ other.R
        t = self.t + (self.R @ other.t.unsqueeze(-1)).squeeze(-1)
        return Pose.from_Rt(R, t)

    __matmul__ = compose

    def transform(self, p3d: torch.Tensor) -> torch.Tensor:
        return p3d @ self.R.transpose(-1, -2) + self.t.unsqueeze(-2)


class Camera(_Wrapped):
    """Pinhole camera, (..., 6) = [w, h, fx, fy, cx, cy] (reference wrappers.py:441-522)."""
    _width = 6
    eps = 1e-3

    @property
    def size(self) -> torch.Tensor:
        return self._data[..., :2]

    @property
    def f(self) -> torch.Tensor:
        return self._data[..., 2:4]

    @property
    def c(self) -> torch.Tensor:
        return self._data[..., 4:6]

    def scale(self, scales) -> "Camera":
        """Camera after resizing the image (reference wrappers.py:478-488)."""
        if isinstance(scales, (int, float)):
            scales = (scales, scales)
        s = self._data.new_tensor(scales)
        return Camera(torch.cat([self.size * s, self.f * s, (self.c + 0.5) * s - 0.5], -1))

    def project(self, p3d: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        z = p3d[..., -1]
        in_front = z > self.eps
        z = z.clamp(min=self.eps)
        p2d = p3d[..., :-1] / z.unsqueeze(-1)
        p2d = p2d * self.f.unsqueeze(-2) + self.c.unsqueeze(-2)
        size = self.size.unsqueeze(-2)
        valid = in_front & torch.all((p2d >= 0) & (p2d <= (size - 1)), -1)
        return p2d, valid


class Obb3D(_Wrapped):
    """Oriented 3D boxes, (..., 19) = bb3_object [xmin,xmax,ymin,ymax,zmin,zmax] | T_world_object (12) | sem_id
    (reference wrappers.py:297-392); produced on the device by ``parq_parse_pred``."""
    _width = 19

    @property
    def bb3_object(self) -> torch.Tensor:
        return self._data[..., :6]

    @property
    def T_world_object(self) -> Pose:
        return Pose(self._data[..., 6:18])

    @property
    def sem_id(self) -> torch.Tensor:
        return self._data[..., 18].unsqueeze(-1)

    @property
    def bb3_size(self) -> torch.Tensor:
        return self._data[..., 1:6:2] - self._data[..., 0:6:2]


def raw(x) -> torch.Tensor:
    """The underlying tensor of a wrapper (ours or the reference's) or a tensor."""
    if isinstance(x, torch.Tensor):
        return x
    d = getattr(x, "_data", None)
    if d is None:
        raise TypeError("expected a tensor or a Pose/Camera wrapper with ._data, got %r" % type(x))
    return d
