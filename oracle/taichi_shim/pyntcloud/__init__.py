"""TEST INFRASTRUCTURE ONLY — stand-in for pyntcloud.PyntCloud.from_file(path).points
(scene.py:95-97): a pandas DataFrame with one float32 column per PLY vertex property."""
import numpy as np
import pandas as pd

_T = {"float": "<f4", "float32": "<f4", "double": "<f8", "uchar": "u1", "int": "<i4", "uint": "<u4",
      "short": "<i2", "ushort": "<u2", "char": "i1"}


class PyntCloud:
    def __init__(self, points):
        self.points = points

    @classmethod
    def from_file(cls, path):
        data = open(path, "rb").read()
        end = data.index(b"end_header")
        hdr_end = data.index(b"\n", end) + 1
        props, n, in_v = [], None, False
        for line in data[:end].decode("ascii").splitlines():
            t = line.split()
            if t[:1] == ["element"]:
                in_v = t[1] == "vertex"
                if in_v:
                    n = int(t[2])
            elif t[:1] == ["property"] and in_v:
                props.append((t[2], _T[t[1]]))
        arr = np.frombuffer(data, dtype=np.dtype(props), count=n, offset=hdr_end)
        return cls(pd.DataFrame({name: arr[name] for name, _ in props}))
