"""Minimal binary-little-endian PLY reader/writer for 3DGS point clouds.

Replaces the reference's ``PyntCloud.from_file`` call (scene.py:95-97): properties are looked up
BY NAME, as a DataFrame would, so any column order works.  Only float32 ("float") vertex
properties are supported — that is what 3DGS exporters and tests/data/test.ply contain.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
          "char": "i1", "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
          "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}

GS_PROPERTIES = (["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
                 + [f"f_rest_{i}" for i in range(45)] + ["opacity", "scale_0", "scale_1", "scale_2",
                                                         "rot_0", "rot_1", "rot_2", "rot_3"])


def read_ply(path) -> dict:
    """Return {property name: 1-D array} for the ``vertex`` element of a binary LE PLY file."""
    data = Path(path).read_bytes()
    end = data.find(b"end_header")
    if not data.startswith(b"ply") or end < 0:
        raise ValueError(f"{path}: not a PLY file")
    header_end = data.index(b"\n", end) + 1
    lines = data[:end].decode("ascii", "replace").splitlines()
    fmt = None
    n_vertex = None
    props = []
    in_vertex = False
    for line in lines:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            in_vertex = tok[1] == "vertex"
            if in_vertex:
                n_vertex = int(tok[2])
            elif n_vertex is None:
                raise ValueError(f"{path}: elements before `vertex` are not supported")
        elif tok[0] == "property" and in_vertex:
            if tok[1] == "list":
                raise ValueError(f"{path}: list properties on vertices are not supported")
            if tok[1] not in _TYPES:
                raise ValueError(f"{path}: unknown property type {tok[1]}")
            props.append((tok[2], _TYPES[tok[1]]))
    if fmt != "binary_little_endian":
        raise ValueError(f"{path}: only binary_little_endian PLY is supported (got {fmt})")
    if n_vertex is None:
        raise ValueError(f"{path}: no vertex element")
    dt = np.dtype(props)
    need = header_end + n_vertex * dt.itemsize
    if len(data) < need:
        raise ValueError(f"{path}: truncated ({len(data)} < {need} bytes)")
    arr = np.frombuffer(data, dtype=dt, count=n_vertex, offset=header_end)
    return {name: np.ascontiguousarray(arr[name]) for name, _ in props}


def write_gs_ply(path, cols: dict):
    """Write the 62-float 3DGS vertex layout of tests/data/test.ply (missing columns = 0)."""
    n = len(next(iter(cols.values())))
    dt = np.dtype([(p, "<f4") for p in GS_PROPERTIES])
    arr = np.zeros(n, dtype=dt)
    for p in GS_PROPERTIES:
        if p in cols:
            arr[p] = np.asarray(cols[p], dtype=np.float32)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"]
    header += [f"property float {p}" for p in GS_PROPERTIES] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(arr.tobytes())
