"""Field-like containers standing in for Taichi fields (``.shape``, ``[i]``, ``.to_numpy()``)."""
from __future__ import annotations

import numpy as np


class StructArrayField:
    """A NumPy structured array with Taichi-field flavoured access: ``f[i]`` returns a record
    object built by ``factory``; ``f.to_numpy()`` the raw structured array."""

    def __init__(self, array: np.ndarray, factory):
        self._a = array
        self._factory = factory

    @property
    def shape(self):
        return self._a.shape

    def __len__(self):
        return self._a.shape[0] if self._a.ndim else 0

    def __getitem__(self, idx):
        rec = self._a[idx]
        if isinstance(rec, np.void):
            return self._factory(rec)
        return StructArrayField(rec, self._factory)

    def __setitem__(self, idx, value):
        self._a[idx] = value.to_record() if hasattr(value, "to_record") else value

    def to_numpy(self):
        return self._a


class DeviceField:
    """A device-resident dense field (torch CUDA tensor) with ``.to_numpy()`` / ``.to_torch()``;
    replaces ``ti.field`` for RayTracer.sample_buf / attenuation_buf / disp_buf
    (ray_tracer.py:33-37)."""

    def __init__(self, tensor):
        self.tensor = tensor

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    def to_numpy(self):
        return self.tensor.detach().cpu().numpy()

    def to_torch(self):
        return self.tensor

    def data_ptr(self):
        return self.tensor.data_ptr()

    def fill(self, v):
        self.tensor.fill_(v)

    def __getitem__(self, idx):
        return self.tensor[idx].detach().cpu().numpy()
