"""Pinhole camera — mirror of the reference's ``rtgs/camera.py`` (camera.py:8-71).

Camera space: +x right, +y up, -z forward.  ``position`` / ``rotation`` are plain mutable
attributes re-read at every ``RayTracer.sample()`` (ray_tracer.py:47-48).  The fused render kernel
generates rays in registers; ``cam_ray_field`` is only materialised when somebody asks for it.
"""
from __future__ import annotations

import numpy as np

from . import _native
from .fields import DeviceField
from .utils.types import vec2, vec2i, vec3, vec4


class Camera:
    def __init__(self, position, rotation, buf_size, focal_length, device: int | None = None):
        self.position = vec3(position)
        self.rotation = vec4(rotation)          # quaternion (x, y, z, w)
        self.buf_size = vec2i(buf_size)
        self.censor_size = self.buf_size        # sic (camera.py:27)
        self.focal_length = vec2(focal_length)
        self.device = device
        self._ray_field = None

    def native(self, position=None, rotation=None) -> "_native.rtgs_camera":
        p = self.position if position is None else position
        r = self.rotation if rotation is None else rotation
        return _native.make_camera(np.asarray(p, np.float32), np.asarray(r, np.float32),
                                   np.asarray(self.focal_length, np.float32), self.buf_size.x, self.buf_size.y)

    def generate_ray_field(self, position=None, rotation=None):
        """Fill ``cam_ray_field`` (camera.py:57-71): (W,H,8) float32 = origin, direction, start 0, end inf."""
        import torch
        from .scene import _default_device
        dev = torch.device("cuda", _default_device() if self.device is None else self.device)
        W, H = self.buf_size.x, self.buf_size.y
        if self._ray_field is None or self._ray_field.tensor.device != dev:
            self._ray_field = DeviceField(torch.empty((W, H, 8), dtype=torch.float32, device=dev))
        cam = self.native(position, rotation)
        _native.check(_native.load().rtgs_generate_rays(cam, dev.index, self._ray_field.data_ptr(),
                                                        torch.cuda.current_stream(dev).cuda_stream))
        return self._ray_field

    @property
    def cam_ray_field(self):
        if self._ray_field is None:
            self.generate_ray_field()
        return self._ray_field
