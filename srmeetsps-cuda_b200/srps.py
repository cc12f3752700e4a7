"""Host-side mirror of the reference's public surface for this path: `Preferences`,
`DataHandler` / `MatFileDataHandler` / `ImageDataHandler` (Utilities.h:166-230) and
`class SRPS { SRPS(DataHandler&); void execute(); }` (SRPS.h:10-18).

`SRPS.execute()` = one-shot init (SRPS.cu:105-270, host side, OpenCV as in the reference) +
the outer loop (SRPS.cu:272-335) through the C ABI -- every loop operator runs in the
hand-written CUDA library; nothing here computes on the CPU inside the loop.
"""
from __future__ import annotations

import glob
import os
import sys
import time

import numpy as np

from .context import Context


class Preferences:                      # Utilities.h:224-230, Main.cpp:5-7
    blockX = 256                        # accepted for CLI compatibility; launch shapes are fixed per kernel
    blockY = 4
    deviceId = 0
    albedo_mode = "closed_form"         # extension: "reference_cg" reproduces devicecalls.cu:540
    headless = True                     # extension: the reference opens GUI windows (SRPS.cu:319-327)


class DataHandler:                      # Utilities.h:166-181
    """I: (h, w, c, n) column-major semantics kept as numpy (n, c, h, w) float32 in [0,1];
    K: 9 floats column-major; mask: (h, w) float32 {0,1}; z0: (z0_n, h/sf, w/sf) float32."""

    def __init__(self):
        self.I = None; self.K = None; self.mask = None; self.sf = None; self.z0 = None
        self.I_w = self.I_h = self.I_c = self.I_n = 0
        self.z0_w = self.z0_h = self.z0_n = 0

    def _finish(self):
        self.I_n, self.I_c, self.I_h, self.I_w = self.I.shape
        self.z0_n, self.z0_h, self.z0_w = self.z0.shape


class MatFileDataHandler(DataHandler):  # Utilities.cpp:159-199
    def loadDataFromMatFiles(self, filename):
        from scipy.io import loadmat
        m = loadmat(filename)
        for k in ("I", "K", "mask", "sf", "z0"):
            if k not in m:
                print("Variable not found, or error reading MAT file", file=sys.stderr)
                raise RuntimeError("Failed reading MAT file")
        I = np.asarray(m["I"], dtype=np.float64)                       # h x w x c x n doubles
        self.I = np.ascontiguousarray(I.transpose(3, 2, 0, 1)).astype(np.float32)
        self.K = np.asarray(m["K"], dtype=np.float64).ravel(order="F").astype(np.float32)
        self.mask = np.asarray(m["mask"]).astype(np.float32)          # uint8 -> float (Utilities.cpp:134-141)
        self.sf = int(np.asarray(m["sf"]).ravel()[0])
        z0 = np.asarray(m["z0"], dtype=np.float64)
        if z0.ndim == 2:
            z0 = z0[:, :, None]
        self.z0 = np.ascontiguousarray(z0.transpose(2, 0, 1)).astype(np.float32)
        self._finish()
        return self


class ImageDataHandler(DataHandler):    # Utilities.cpp:349-395
    def loadDataFromImages(self, dataFolder):
        import cv2
        rgb = sorted(glob.glob(os.path.join(dataFolder, "RGB", "*")))
        if not rgb:
            raise RuntimeError(f"no images under {dataFolder}/RGB")
        imgs = [cv2.imread(f)[:, :, ::-1].transpose(2, 0, 1) for f in rgb]      # BGR -> reversed channel order
        self.I8 = np.stack(imgs).astype(np.uint8)
        self.I = (self.I8.astype(np.float32) / np.float32(255.0)).astype(np.float32)
        with open(os.path.join(dataFolder, "K.txt")) as fh:
            lines = [ln.strip() for ln in fh.read().strip().splitlines()]
        K = np.zeros(9, dtype=np.float32)
        for i in range(3):
            vals = [float(v) for v in lines[i].split(",")]
            for j in range(3):
                K[i + 3 * j] = vals[j]
        sf, min_z, max_z = [np.float32(float(v)) for v in lines[3].split(",")]
        self.K = K
        self.sf = int(sf)
        mask8 = cv2.imread(os.path.join(dataFolder, "mask.png"), cv2.IMREAD_GRAYSCALE)
        self.mask = mask8.astype(np.float32) / np.float32(255.0)
        depth = sorted(glob.glob(os.path.join(dataFolder, "Depth", "*")))
        self.z0 = np.stack([min_z + (cv2.imread(f, cv2.IMREAD_ANYDEPTH).astype(np.float32) / np.float32(65535.0)) * (max_z - min_z)
                            for f in depth]).astype(np.float32)
        self._finish()
        return self


def preprocess_depth(z0, h, w, sf):
    """Host one-shot (SRPS.cu:117-149, devicecalls.cu:95-125): mean over frames (always /nc, zero
    frames flagged), TELEA inpaint r=16, bilateral(-1,2,2) on depth/max, bicubic upsample -- on the
    transposed image, as the reference's cv::Mat(z0_w, z0_h) view does."""
    import cv2
    hs, ws = h // sf, w // sf
    fr = np.stack([f.ravel(order="F") for f in np.asarray(z0, dtype=np.float32)])
    flags = (fr == 0).any(axis=0).astype(np.uint8)
    zs = (fr.sum(axis=0, dtype=np.float32) / np.float32(fr.shape[0])).astype(np.float32)
    zs_mat = cv2.inpaint(zs.reshape(ws, hs).copy(), flags.reshape(ws, hs).copy(), 16, cv2.INPAINT_TELEA)
    mx = float(zs_mat.max())
    sm = cv2.bilateralFilter((zs_mat / mx).astype(np.float32), -1, 2, 2) * np.float32(mx)
    z_full = cv2.resize(sm, (h, w), interpolation=cv2.INTER_CUBIC)
    return sm.reshape(-1).astype(np.float32), z_full.reshape(-1).astype(np.float32)


class SRPS:                             # SRPS.h:10-18
    def __init__(self, dh: DataHandler):
        self.dh = dh
        self.history = []
        self.result = None

    def execute(self, out=sys.stdout, out_dir=None):
        """SRPS::execute (SRPS.cu:84-349).  out_dir (extension): write the reference's dumps and renderings there
        (output.save_results: s/rho/z/N.mat, normals/albedo/depth.png) instead of opening windows (SRPS.cu:319-333)."""
        dh = self.dh
        TOLERANCE, MAX_ITERATIONS = 5e-3, 10                       # SRPS.cu:85-86
        h, w, sf = dh.I_h, dh.I_w, int(dh.sf)
        mask = (np.asarray(dh.mask) != 0)
        print("Small mask calculation", file=out)                  # SRPS.cu:106
        ctx = Context(mask, dh.I_n, sf, dh.K, device=Preferences.deviceId, albedo_mode=Preferences.albedo_mode)
        print("Mean of depth values", file=out)                    # SRPS.cu:119
        print("Inpainting depth values", file=out)                 # SRPS.cu:129
        print("Smoothing depth", file=out)                         # SRPS.cu:135
        zs, z_full = preprocess_depth(dh.z0, h, w, sf)
        print("Resample depths", file=out)                         # SRPS.cu:146
        print("Mask index calculation", file=out)                  # SRPS.cu:152
        mflat = mask.ravel(order="F")
        hs, ws = h // sf, w // sf
        lr = mask.reshape(hs, sf, ws, sf).all(axis=(1, 3))         # LR mask (SRPS.cu:110-111)
        print("Masked resample matrix", file=out)                  # SRPS.cu:171
        print("Masked gradient matrix", file=out)                  # SRPS.cu:196
        print("Initialization", file=out)                          # SRPS.cu:206
        # masked stack in the reference layout [n][c][npix] (SRPS.cu:223-232), column-major pixel order
        I_masked = np.ascontiguousarray(dh.I.transpose(0, 1, 3, 2).reshape(dh.I_n, dh.I_c, h * w)[:, :, mflat])
        z = z_full[mflat]
        z0s = zs[lr.ravel(order="F")]
        ctx.upload_state(I_masked, z, z0s)
        last_error = float("nan")
        iteration = 1
        self.history = []
        while True:                                                # SRPS.cu:276-335
            t0 = time.perf_counter()
            energy, cg = ctx.outer_iteration()
            tm = ctx.timings()
            print("\n%-25s: %-6.6fs" % ("Lightning Estimation", tm["ms_lighting"] * 1e-3), file=out)
            print("%-25s: %-6.6fs" % ("Albedo Estimation", tm["ms_albedo"] * 1e-3), file=out)
            print("%-25s: %-6.6fs" % ("Depth Estimation", tm["ms_depth"] * 1e-3), file=out)
            rel_err = abs(last_error - energy) / abs(energy)
            stop = (energy > last_error) or (rel_err < TOLERANCE) or (iteration > MAX_ITERATIONS)
            last_error = energy
            print("\nIteration %02d summary" % iteration, file=out)
            print("%-25s: %-6.3f" % ("Error", energy), file=out)
            print("%-25s: %-6.3f" % ("Relative Error", rel_err), file=out)
            self.history.append(dict(iteration=iteration, energy=energy, rel_err=rel_err,
                                     wall_s=time.perf_counter() - t0, **tm))
            iteration += 1
            if stop:
                break
        print("Done!", file=out)                                   # SRPS.cu:337
        self.result = dict(z=ctx.download("z"), rho=ctx.download("rho"), N=ctx.download("N"), s=ctx.download("s"),
                           mask=mask)
        ctx.close()
        if out_dir:
            from .output import save_results
            save_results(self.result, out_dir)
        return self.result
