"""SRPSNAP1: a tiny named-array container for post-init loop state and per-iteration dumps.

Layout (little endian):  magic "SRPSNAP1" | int32 count | count x { char name[24] | int32 dtype
| int32 ndim | int64 dims[4] | raw data, zero-padded to a multiple of 8 bytes }.
dtype: 0=float32 1=int32 2=uint8 3=float64.  The same format is read and written by the C++
host (src/host/snapshot.h) and by the reference replay driver (oracle/ref/ref_replay.cu).
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"SRPSNAP1"
_DT = {0: np.float32, 1: np.int32, 2: np.uint8, 3: np.float64}
_CODE = {np.dtype(v): k for k, v in _DT.items()}


def write_snapshot(path, arrays: dict):
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<i", len(arrays)))
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            if a.dtype not in _CODE:
                raise TypeError(f"{name}: unsupported dtype {a.dtype}")
            if a.ndim > 4:
                raise ValueError(f"{name}: ndim > 4")
            nb = name.encode()
            if len(nb) > 23:
                raise ValueError(f"{name}: name too long")
            dims = list(a.shape) + [1] * (4 - a.ndim)
            f.write(nb.ljust(24, b"\0"))
            f.write(struct.pack("<ii4q", _CODE[a.dtype], a.ndim, *dims))
            raw = a.tobytes()
            f.write(raw)
            f.write(b"\0" * ((-len(raw)) % 8))


def read_snapshot(path) -> dict:
    out = {}
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not an SRPSNAP1 file")
        (count,) = struct.unpack("<i", f.read(4))
        for _ in range(count):
            name = f.read(24).split(b"\0", 1)[0].decode()
            code, ndim, *dims = struct.unpack("<ii4q", f.read(40))
            shape = tuple(dims[:ndim])
            dt = np.dtype(_DT[code])
            nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
            raw = f.read(nbytes)
            f.read((-nbytes) % 8)
            out[name] = np.frombuffer(raw, dtype=dt).reshape(shape).copy()
    return out
