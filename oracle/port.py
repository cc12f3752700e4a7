"""ctypes driver for oracle/srps_oracle_port.c -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Builds the stencil index arrays the C transcription consumes from the reference-style
operators of oracle/srps_oracle.py (make_gradient: SRPS.cu:23-71; KT: SRPS.cu:172-193) and
calls one outer iteration (SRPS.cu:276-317).  Never imported by the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libsrps_oracle_port.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "srps_oracle_port.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC",
                               "-o", _LIB, src, "-lm"])
    return _LIB


class Geom(C.Structure):
    _fields_ = [("npix", C.c_int), ("npixs", C.c_int), ("n", C.c_int), ("c", C.c_int), ("sf", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float),
                ("dx_nb", C.c_void_p), ("dy_nb", C.c_void_p), ("dx_sg", C.c_void_p), ("dy_sg", C.c_void_p),
                ("x_prev_f", C.c_void_p), ("x_next_b", C.c_void_p), ("y_prev_f", C.c_void_p), ("y_next_b", C.c_void_p),
                ("kt_idx", C.c_void_p), ("lr_of", C.c_void_p), ("xx", C.c_void_p), ("yy", C.c_void_p)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Port:
    """Holds the geometry arrays (kept alive) and the loaded library."""

    def __init__(self, ops, n, c, fx, fy, xx, yy):
        self.lib = C.CDLL(build())
        h, w, sf = ops["h"], ops["w"], ops["sf"]
        imask = ops["imask"]
        npix = imask.size
        idx2 = np.full((h, w), -1, dtype=np.int64)
        idx2.ravel(order="K")  # no-op; documentation: idx2[i, j] below
        idx_lin = np.full(h * w, -1, dtype=np.int64)
        idx_lin[imask] = np.arange(npix)
        idx2 = idx_lin.reshape((h, w), order="F")
        t = ops["types"]
        me = np.arange(npix)
        ii, jj = imask % h, imask // h

        def direction(fwd, bwd, di, dj):
            f = fwd[ii, jj]; b = bwd[ii, jj]
            nb = me.copy(); sg = np.zeros(npix, dtype=np.float32)
            nb[f] = idx2[ii[f] + di, jj[f] + dj]; sg[f] = 1
            nb[b] = idx2[ii[b] - di, jj[b] - dj]; sg[b] = -1
            prev_f = np.full(npix, -1, dtype=np.int64)
            next_b = np.full(npix, -1, dtype=np.int64)
            # prev pixel exists in mask and is a forward row -> it references me with +1
            ip, jp = ii - di, jj - dj
            ok = (ip >= 0) & (jp >= 0)
            cand = np.where(ok, idx2[np.clip(ip, 0, h - 1), np.clip(jp, 0, w - 1)], -1)
            isf = np.zeros(npix, dtype=bool)
            isf[cand >= 0] = fwd[np.clip(ip, 0, h - 1), np.clip(jp, 0, w - 1)][cand >= 0]
            prev_f[isf] = cand[isf]
            inn, jn = ii + di, jj + dj
            ok = (inn < h) & (jn < w)
            cand = np.where(ok, idx2[np.clip(inn, 0, h - 1), np.clip(jn, 0, w - 1)], -1)
            isb = np.zeros(npix, dtype=bool)
            isb[cand >= 0] = bwd[np.clip(inn, 0, h - 1), np.clip(jn, 0, w - 1)][cand >= 0]
            next_b[isb] = cand[isb]
            return nb.astype(np.int32), sg, prev_f.astype(np.int32), next_b.astype(np.int32)

        self.dy_nb, self.dy_sg, self.y_prev_f, self.y_next_b = direction(t["yf"], t["yb"], 1, 0)
        self.dx_nb, self.dx_sg, self.x_prev_f, self.x_next_b = direction(t["xf"], t["xb"], 0, 1)
        KT = ops["KT"].tocsr()
        npixs = KT.shape[0]
        self.kt_idx = np.ascontiguousarray(KT.indices.reshape(npixs, sf * sf).astype(np.int32)) if npixs else np.zeros((0, sf * sf), np.int32)
        lr_of = np.full(npix, -1, dtype=np.int32)
        if npixs:
            lr_of[self.kt_idx.ravel()] = np.repeat(np.arange(npixs, dtype=np.int32), sf * sf)
        self.lr_of = lr_of
        self.xx = np.ascontiguousarray(xx, dtype=np.float32)
        self.yy = np.ascontiguousarray(yy, dtype=np.float32)
        self.g = Geom(npix, npixs, n, c, sf, fx, fy, _ptr(self.dx_nb), _ptr(self.dy_nb), _ptr(self.dx_sg),
                      _ptr(self.dy_sg), _ptr(self.x_prev_f), _ptr(self.x_next_b), _ptr(self.y_prev_f),
                      _ptr(self.y_next_b), _ptr(self.kt_idx), _ptr(self.lr_of), _ptr(self.xx), _ptr(self.yy))
        self.lib.srps_port_outer_iteration.restype = C.c_float
        self.lib.srps_port_depth.restype = C.c_float
        self.lib.srps_port_threads.restype = C.c_int

    def threads(self):
        return int(self.lib.srps_port_threads())

    def normals(self, z):
        P = self.g.npix
        N = np.empty((4, P), np.float32); dz = np.empty(P, np.float32)
        self.lib.srps_port_normals(C.byref(self.g), _ptr(z), _ptr(N), _ptr(dz))
        return N, dz

    def outer_iteration(self, st, albedo_closed_form=False):
        """st: dict with float32 C-contiguous s, rho, z, N, dz, I, z0s (updated in place)."""
        for k in ("s", "rho", "z", "N", "dz", "I", "z0s"):
            assert st[k].dtype == np.float32 and st[k].flags.c_contiguous, k
        dk = C.c_int(0)
        ak = (C.c_int * 3)()
        e = self.lib.srps_port_outer_iteration(C.byref(self.g), _ptr(st["s"]), _ptr(st["rho"]), _ptr(st["z"]),
                                               _ptr(st["N"]), _ptr(st["dz"]), _ptr(st["I"]), _ptr(st["z0s"]),
                                               C.c_int(int(albedo_closed_form)), C.byref(dk), ak)
        return float(e), int(dk.value), list(ak)
