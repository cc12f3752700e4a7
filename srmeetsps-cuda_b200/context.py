"""Python host wrapper over the C ABI (include/srps_c_api.h): one `Context` per scene.

All arrays crossing this boundary are HOST numpy arrays in the reference's masked-vector
layouts (SRPS.cu:157-162, 209-260); the context owns every device buffer.  Torch is optional and
only used by callers that want pinned host memory.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class SRPSError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    def __init__(self, mask, n_images, sf, K, device=0, albedo_mode="closed_form", cg_max_iter=0, cg_tol=0.0,
                 n_channels=3, strip=None, rank=0, world=1):
        """mask: (h, w) array, non-zero = inside.  K: the reference's column-major 3x3
        (K[0]=fx, K[4]=fy, K[6]=cx, K[7]=cy; Utilities.cpp:364-373).
        strip=(j0, j1), rank, world: this context owns image columns [j0, j1) of the GLOBAL mask
        (strip partition across GPUs, see dist.py); call dist_connect before any operator."""
        self.lib = L.load()
        mask = np.asarray(mask)
        h, w = mask.shape
        self._mask_cm = np.ascontiguousarray((mask != 0).astype(np.uint8).ravel(order="F"))
        K = np.asarray(K, dtype=np.float64).ravel()
        mode = {"closed_form": L.SRPS_ALBEDO_CLOSED_FORM, "reference_cg": L.SRPS_ALBEDO_REFERENCE_CG}[albedo_mode]
        self.prob = L.Problem(h, w, int(n_images), int(n_channels), int(sf), float(K[0]), float(K[4]), float(K[6]),
                              float(K[7]), _ptr(self._mask_cm), int(device), mode, int(cg_max_iter), float(cg_tol),
                              int(strip[0]) if strip else 0, int(strip[1]) if strip else 0, int(rank), int(world))
        self.h, self.w, self.n, self.c, self.sf = h, w, int(n_images), int(n_channels), int(sf)
        self._ctx = C.c_void_p()
        rc = self.lib.srps_ctx_create(C.byref(self.prob), C.byref(self._ctx))
        if rc != 0:
            raise SRPSError(f"srps_ctx_create failed ({rc}): {self.lib.srps_last_error(None).decode()}")
        self.npix = self.lib.srps_npix(self._ctx)
        self.npixs = self.lib.srps_npixs(self._ctx)

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.srps_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, what):
        if rc != 0:
            raise SRPSError(f"{what} failed ({rc}): {self.lib.srps_last_error(self._ctx).decode()}")

    # -- strip partition ------------------------------------------------------------------------
    def dist_export(self) -> bytes:
        buf = C.create_string_buffer(self.lib.srps_dist_blob_size())
        self._ck(self.lib.srps_dist_export(self._ctx, buf), "srps_dist_export")
        return buf.raw

    def dist_connect(self, blobs):
        raw = b"".join(blobs)
        assert len(raw) == len(blobs) * self.lib.srps_dist_blob_size()
        self._ck(self.lib.srps_dist_connect(self._ctx, raw, len(blobs)), "srps_dist_connect")

    def pixel_range(self):
        v = [C.c_longlong(0) for _ in range(4)]
        self._ck(self.lib.srps_pixel_range(self._ctx, *[C.byref(x) for x in v]), "srps_pixel_range")
        return tuple(int(x.value) for x in v)

    def upload_state_strided(self, I_first, plane_stride, z, z0s):
        """I_first: 1-D float32 view starting at this strip's first pixel of plane 0 of the global stack."""
        z = np.ascontiguousarray(z, dtype=np.float32); z0s = np.ascontiguousarray(z0s, dtype=np.float32)
        assert I_first.dtype == np.float32 and z.shape == (self.npix,) and z0s.shape == (self.npixs,)
        self._ck(self.lib.srps_upload_state_strided(self._ctx, _ptr(I_first), int(plane_stride), _ptr(z), _ptr(z0s)),
                 "srps_upload_state_strided")

    # -- state ----------------------------------------------------------------------------------
    def upload_state(self, I, z, z0s):
        """I: (n, c, npix) float32 or None (after upload_images_u8); z: (npix,); z0s: (npixs,)."""
        if I is not None:
            I = np.ascontiguousarray(I, dtype=np.float32)
            assert I.shape == (self.n, self.c, self.npix), (I.shape, (self.n, self.c, self.npix))
        z = np.ascontiguousarray(z, dtype=np.float32)
        z0s = np.ascontiguousarray(z0s, dtype=np.float32)
        assert z.shape == (self.npix,) and z0s.shape == (self.npixs,)
        self._ck(self.lib.srps_upload_state(self._ctx, _ptr(I) if I is not None else None, _ptr(z), _ptr(z0s)),
                 "srps_upload_state")

    def upload_images_u8(self, I8):
        """The stack as 8-bit samples (kept 8-bit on the device; I = v/255 is formed in registers)."""
        I8 = np.ascontiguousarray(I8, dtype=np.uint8)
        assert I8.shape == (self.n, self.c, self.npix)
        self._ck(self.lib.srps_upload_images_u8(self._ctx, _ptr(I8)), "srps_upload_images_u8")

    def upload_images_u8_strided(self, I8_first, plane_stride):
        """Strip contexts: I8_first is a 1-D uint8 view starting at this strip's first pixel of plane 0 of the global stack."""
        assert I8_first.dtype == np.uint8
        self._ck(self.lib.srps_upload_images_u8_strided(self._ctx, _ptr(I8_first), int(plane_stride)),
                 "srps_upload_images_u8_strided")

    _SHAPES = {L.BUF_S: lambda s: (s.n, s.c, 4), L.BUF_RHO: lambda s: (s.c, s.npix), L.BUF_Z: lambda s: (s.npix,),
               L.BUF_N: lambda s: (4, s.npix), L.BUF_DZ: lambda s: (s.npix,), L.BUF_Z0S: lambda s: (s.npixs,),
               L.BUF_W: lambda s: (3, s.npix), L.BUF_G: lambda s: (3, s.npix), L.BUF_E0: lambda s: (s.npix,),
               L.BUF_R: lambda s: (s.npix,)}
    _NAMES = {"s": L.BUF_S, "rho": L.BUF_RHO, "z": L.BUF_Z, "N": L.BUF_N, "dz": L.BUF_DZ, "z0s": L.BUF_Z0S,
              "w": L.BUF_W, "g": L.BUF_G, "e0": L.BUF_E0, "r": L.BUF_R}

    def download(self, name, out=None):
        which = self._NAMES[name]
        shape = self._SHAPES[which](self)
        if out is None:
            out = np.empty(shape, dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == shape
        self._ck(self.lib.srps_download(self._ctx, which, _ptr(out)), f"srps_download({name})")
        return out

    def set_state(self, name, value):
        which = self._NAMES[name]
        v = np.ascontiguousarray(value, dtype=np.float32)
        assert v.shape == self._SHAPES[which](self), (v.shape, self._SHAPES[which](self))
        self._ck(self.lib.srps_set_state(self._ctx, which, _ptr(v)), f"srps_set_state({name})")

    # -- operators (the loop body, SRPS.cu:276-317) ---------------------------------------------
    def lighting(self):
        self._ck(self.lib.srps_lighting(self._ctx), "srps_lighting")

    def albedo(self):
        self._ck(self.lib.srps_albedo(self._ctx), "srps_albedo")

    def depth(self):
        e = C.c_float(0)
        k = C.c_int(0)
        self._ck(self.lib.srps_depth(self._ctx, C.byref(e), C.byref(k)), "srps_depth")
        return float(e.value), int(k.value)

    def normals(self):
        self._ck(self.lib.srps_normals(self._ctx), "srps_normals")

    def outer_iteration(self):
        e = C.c_float(0)
        k = C.c_int(0)
        self._ck(self.lib.srps_outer_iteration(self._ctx, C.byref(e), C.byref(k)), "srps_outer_iteration")
        return float(e.value), int(k.value)

    def run(self, max_outer=10, tol=5e-3, fixed_iters=0):
        cap = max(max_outer + 1, fixed_iters, 1) + 1
        energies = np.zeros(cap, dtype=np.float32)
        n = C.c_int(0)
        self._ck(self.lib.srps_run(self._ctx, int(max_outer), float(tol), int(fixed_iters), _ptr(energies), cap,
                                   C.byref(n)), "srps_run")
        return energies[: n.value].copy()

    def timings(self):
        t = L.Timings()
        self._ck(self.lib.srps_get_timings(self._ctx, C.byref(t)), "srps_get_timings")
        return dict(ms_lighting=t.ms_lighting, ms_albedo=t.ms_albedo, ms_depth=t.ms_depth, ms_normals=t.ms_normals,
                    ms_total=t.ms_total, ms_depth_cg=t.ms_depth_cg, cg_iters=t.cg_iters,
                    albedo_cg_iters=list(t.albedo_cg_iters), launches=int(t.launches), cg_deferred=int(t.cg_deferred), cg_zskip=int(t.cg_zskip))

    def synchronize(self):
        self._ck(self.lib.srps_synchronize(self._ctx), "srps_synchronize")

    def timer_start(self):
        self._ck(self.lib.srps_timer_start(self._ctx), "srps_timer_start")

    def timer_stop(self):
        ms = C.c_float(0)
        self._ck(self.lib.srps_timer_stop(self._ctx, C.byref(ms)), "srps_timer_stop")
        return float(ms.value)

    def profile_kernels(self, reps=20):
        """Average device ms of the dominant kernels timed alone (state is undefined afterwards)."""
        out = (C.c_float * 6)()
        self._ck(self.lib.srps_profile_kernels(self._ctx, int(reps), out), "srps_profile_kernels")
        return dict(cg_stencil=out[0], cg_update=out[1], lighting_pass=out[2], project_pass=out[3], cg_fused=out[4],
                    cg_driver={0: "graph", 1: "persistent", 2: "fused", 3: "persistent_fused"}[int(out[5])])

    def apply_depth_operator(self, p):
        p = np.ascontiguousarray(p, dtype=np.float32)
        assert p.shape == (self.npix,)
        y = np.empty(self.npix, dtype=np.float32)
        self._ck(self.lib.srps_apply_depth_operator(self._ctx, _ptr(p), _ptr(y)), "srps_apply_depth_operator")
        return y
