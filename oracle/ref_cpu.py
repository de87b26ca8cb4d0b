"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — ctypes wrapper of oracle/ref_cpu.cpp.

`CpuScene.render(..., precision="float")` is the reference-shaped CPU path in the reference's own
arithmetic type (the timed CPU baseline); `precision="double"` is the large-scene oracle;
`render_brute` is the BVH-free float64 cross-check.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libref_cpu.so"


def build(force: bool = False) -> Path:
    """Compile ref_cpu.cpp with OpenMP.  Tries the system g++ first (the image's $CXX wrapper lacks
    libgomp.spec), then $CXX, and finally a single-threaded build without OpenMP."""
    src = HERE / "ref_cpu.cpp"
    if not force and LIB.exists() and LIB.stat().st_mtime >= src.stat().st_mtime:
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    base = ["-O3", "-march=native", "-fPIC", "-std=c++17", "-shared", "-o", str(LIB), str(src)]
    tries = [(cxx, ["-fopenmp"]) for cxx in ("/usr/bin/g++", "g++", os.environ.get("CXX", "")) if cxx]
    tries.append(("g++", ["-DRTGS_NO_OMP"]))
    errs = []
    for cxx, extra in tries:
        r = subprocess.run([cxx, *extra, *base], capture_output=True, text=True)
        if r.returncode == 0:
            return LIB
        errs.append(f"{cxx} {extra}: {r.stderr.strip()[-300:]}")
    raise RuntimeError("could not build oracle/ref_cpu.cpp:\n" + "\n".join(errs))


class _Cam(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("rotation", C.c_float * 4), ("focal", C.c_float * 2),
                ("width", C.c_int32), ("height", C.c_int32)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(LIB))
        lib.rc_build.restype = C.c_void_p
        lib.rc_build.argtypes = [C.c_int64] + [C.c_void_p] * 6
        lib.rc_free.argtypes = [C.c_void_p]
        lib.rc_read_lbvh.argtypes = [C.c_void_p] * 6
        lib.rc_max_threads.restype = C.c_int
        lib.rc_set_restart_eps.argtypes = [C.c_double]
        lib.rc_set_restart_eps.restype = None
        lib.rc_render.restype = C.c_int
        lib.rc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.rc_render_brute.restype = C.c_int
        lib.rc_render_brute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(_load().rc_max_threads())


def all_pixels(W, H, stride=1):
    ii, jj = np.meshgrid(np.arange(0, W, stride), np.arange(0, H, stride), indexing="ij")
    return np.ascontiguousarray(np.stack([ii.ravel(), jj.ravel()], axis=-1), dtype=np.int32)


class CpuScene:
    def __init__(self, pos, rot, scale, color, opacity, sh=None):
        f = np.float32
        self.n = int(np.asarray(pos).shape[0])
        self._a = [np.ascontiguousarray(x, dtype=f) for x in (pos, rot, scale, color, opacity)]
        self._sh = None if sh is None else np.ascontiguousarray(sh, dtype=f).reshape(self.n, 45)
        self._h = _load().rc_build(self.n, *[a.ctypes.data for a in self._a],
                                   None if self._sh is None else self._sh.ctypes.data)
        if not self._h:
            raise RuntimeError("rc_build failed")

    def __del__(self):
        try:
            if self._h:
                _load().rc_free(self._h)
                self._h = None
        except Exception:
            pass

    def read_lbvh(self):
        n = self.n
        out = dict(morton=np.empty(n, np.uint32), sorted_idx=np.empty(n, np.uint32),
                   child=np.empty((max(n - 1, 0), 2), np.int32), parent=np.empty(2 * n - 1, np.int32),
                   aabb=np.empty((2 * n - 1, 6), np.float32))
        _load().rc_read_lbvh(self._h, out["morton"].ctypes.data, out["sorted_idx"].ctypes.data,
                             out["child"].ctypes.data if n > 1 else None, out["parent"].ctypes.data,
                             out["aabb"].ctypes.data)
        return out

    @staticmethod
    def _cam(cam):
        c = _Cam()
        c.position[:] = [float(v) for v in cam.position]
        c.rotation[:] = [float(v) for v in cam.rotation]
        c.focal[:] = [float(cam.focal[0]), float(cam.focal[1])]
        c.width, c.height = int(cam.width), int(cam.height)
        return c

    def render(self, cam, depth=16, pixels=None, precision="double", threads=0, brute=False, restart_eps=None):
        """cam: oracle.ref_numpy.CameraParams.  Returns dict(rgb (npix,3) f64, T, nlayers|nhit, counters).

        restart_eps: what is added to ray.start after every layer (ray_tracer.py:100-102 has 1e-8).  Default: 0 for
        the float64 image oracle (every distinct crossing counts, like the NumPy brute force; see ref_cpu.cpp),
        the reference's literal 1e-8 for the float32 timed baseline."""
        if restart_eps is None:
            restart_eps = 1e-8 if precision == "float" else 0.0
        _load().rc_set_restart_eps(C.c_double(float(restart_eps)))
        if pixels is None:
            pixels = all_pixels(cam.width, cam.height)
        pixels = np.ascontiguousarray(pixels, dtype=np.int32)
        npix = pixels.shape[0]
        rgb = np.empty((npix, 3), np.float64)
        T = np.empty(npix, np.float64)
        cnt = np.empty(npix, np.int32)
        c = self._cam(cam)
        prec = 0 if precision == "float" else 1
        if brute:
            st = _load().rc_render_brute(self._h, C.byref(c), int(depth), prec, npix, pixels.ctypes.data,
                                         rgb.ctypes.data, T.ctypes.data, cnt.ctypes.data, int(threads))
            out = dict(rgb=rgb, T=T, nhit=cnt)
        else:
            counters = np.zeros(2, np.uint64)
            st = _load().rc_render(self._h, C.byref(c), int(depth), prec, npix, pixels.ctypes.data, rgb.ctypes.data,
                                   T.ctypes.data, cnt.ctypes.data, counters.ctypes.data, int(threads))
            out = dict(rgb=rgb, T=T, nlayers=cnt, node_visits=int(counters[0]), gaussian_tests=int(counters[1]))
        if st != 0:
            raise RuntimeError("ref_cpu render failed")
        return out
