"""Dataset loading + one-shot init for the oracle -- TEST INFRASTRUCTURE ONLY.

Restates the reference's image-folder loader (Utilities.cpp:322-395) and the one-shot
pre-processing of SRPS::execute (SRPS.cu:105-149, 151-260) with python cv2, producing the
post-init loop state ("snapshot") every implementation starts from.  Reads
/root/reference/dataset only in the build container; the snapshot it writes
(tests/golden/mitten_init.npz) is what travels.
"""
from __future__ import annotations

import glob
import os

import numpy as np

from . import srps_oracle as o


def load_image_folder(folder):
    """ImageDataHandler::loadDataFromImages (Utilities.cpp:349-395).  Returns I (n, c, h, w)
    in [0,1] with channel order reversed from OpenCV's BGR (Utilities.cpp:341-345 -> RGB),
    K (column-major 3x3, 9 floats), mask (h, w) in {0,1}, sf, z0 (z0_n, h/sf, w/sf)."""
    import cv2
    rgb = sorted(glob.glob(os.path.join(folder, "RGB", "*")))          # cv::glob is lexicographic
    imgs = []
    for f in rgb:
        bgr = cv2.imread(f)                                            # 8-bit BGR
        imgs.append(bgr[:, :, ::-1].transpose(2, 0, 1))                # (c=RGB, h, w)
    I8 = np.stack(imgs).astype(np.uint8)
    with open(os.path.join(folder, "K.txt")) as fh:
        lines = [ln.strip() for ln in fh.read().strip().splitlines()]
    K = np.zeros(9, dtype=np.float32)
    for i in range(3):
        vals = [np.float32(float(v)) for v in lines[i].split(",")]
        for j in range(3):
            K[i + 3 * j] = vals[j]                                     # Utilities.cpp:364-373
    sf, min_z, max_z = [np.float32(float(v)) for v in lines[3].split(",")]
    mask8 = cv2.imread(os.path.join(folder, "mask.png"), cv2.IMREAD_GRAYSCALE)
    mask = (mask8.astype(np.float32) / np.float32(255.0))              # quantizer 255 -> {0,1}
    depth = sorted(glob.glob(os.path.join(folder, "Depth", "*")))
    z0 = []
    for f in depth:
        d16 = cv2.imread(f, cv2.IMREAD_ANYDEPTH).astype(np.float32)
        z0.append(min_z + (d16 / np.float32(65535.0)) * (max_z - min_z))   # Utilities.cpp:330
    return dict(I8=I8, K=K, mask=mask, sf=int(sf), z0=np.stack(z0).astype(np.float32))


def post_init_snapshot(I8, K, mask, sf, z0):
    """SRPS.cu:105-260 up to (not including) the first normal_init: LR mask, depth mean ->
    inpaint -> bilateral -> bicubic, masked gathers.  I stays uint8 (I = I8/255 exactly as the
    loader computes it)."""
    n, c, h, w = I8.shape
    ops = o.build_operators(mask, sf)
    z0_cm = np.stack([f.ravel(order="F") for f in z0])                 # column-major frames
    zs, z_full = o.preprocess_depth(z0_cm, h, w, sf)
    I8_masked = np.stack([[I8[i, ch].ravel(order="F")[ops["imask"]] for ch in range(c)] for i in range(n)])
    return dict(h=h, w=w, sf=sf, mask=(mask != 0).astype(np.uint8), K=np.asarray(K, np.float32),
                I8=I8_masked.astype(np.uint8), z=z_full[ops["imask"]].astype(np.float32),
                z0s=zs[ops["imasks"]].astype(np.float32))


def scene_from_snapshot(snap):
    """npz/dict -> the scene dict used by tests (same keys as srps_oracle.synth_scene)."""
    mask = np.asarray(snap["mask"]).astype(np.float32)
    sf = int(snap["sf"])
    ops = o.build_operators(mask, sf)
    if "I8" in snap:
        I = (np.asarray(snap["I8"]).astype(np.float32) / np.float32(255.0)).astype(np.float32)
    else:
        I = np.asarray(snap["I"], np.float32)
    return dict(h=mask.shape[0], w=mask.shape[1], sf=sf, n=I.shape[0], c=I.shape[1], mask=mask,
                K=np.asarray(snap["K"], np.float64), I=I, z=np.asarray(snap["z"], np.float32),
                z0s=np.asarray(snap["z0s"], np.float32), ops=ops)


def replay_snapshot_arrays(scene):
    """Arrays the reference replay driver (oracle/ref/ref_replay.cu) consumes: state in the
    reference's masked layouts + Dx, Dy, KT as COO triplets (SRPS.cu:23-71, 172-193)."""
    ops = scene["ops"]
    K = np.asarray(scene["K"], np.float32)
    xx, yy = o.meshgrid_masked(ops, float(K[6]), float(K[7]), np.float32)
    out = dict(I=scene["I"].astype(np.float32), z=scene["z"].astype(np.float32),
               z0s=scene["z0s"].astype(np.float32), xx=xx, yy=yy, K=K,
               dims=np.array([scene["h"], scene["w"], scene["sf"]], np.int32))
    for name in ("Dx", "Dy", "KT"):
        m = ops[name].tocoo()
        out[name + "_shape"] = np.array(m.shape, np.int32)
        out[name + "_row"] = m.row.astype(np.int32)
        out[name + "_col"] = m.col.astype(np.int32)
        out[name + "_val"] = m.data.astype(np.float32)
    return out
