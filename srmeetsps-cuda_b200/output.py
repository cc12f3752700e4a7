"""Result output: the reference's dumps and visualisations, headless.

The reference shows three windows per outer iteration and writes four MAT files through a macro that is compiled out on
Linux (SRPS.cu:319-333).  Here the same content goes to files after the loop, with numpy + zlib only (no OpenCV, no matio):

  normals.png   N_as_opencv_mat     Utilities.cpp:277-298   (0.5 + 0.5 N0, 0.5 + 0.5 N1, 0.5 - 0.5 N2), clipped to [0,1],
                                                             min-max normalised over the whole image
  albedo.png    rho_as_opencv_mat   Utilities.cpp:242-275   per channel capped at median + 5 sigma, clipped to [0,1]
  depth.png     z_as_opencv_mat     Utilities.cpp:300-320   -z min-max normalised over the mask, 8 bit, BONE colour map
  s.mat rho.mat z.mat N.mat         SRPS.cu:329-332         one single-precision column vector "x" each, in the reference's
                                                             masked layouts; MAT v5 (the reference asks matio for v7.3 = HDF5)

Pixels outside the mask are black, images are full resolution (the reference scales its windows by 0.425 for the screen).
Masked vectors are in column-major pixel order (SRPS.cu:157-162); PNGs are written as RGB.
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np


def _scatter(values, mask, channels):
    """values[channels][npix] (masked, column-major order) -> float32 image h x w x channels, zero outside the mask."""
    mask = np.asarray(mask) != 0
    h, w = mask.shape
    img = np.zeros((w, h, channels), np.float32)                 # transposed view: column-major pixel order is row-major here
    img[mask.T] = np.asarray(values, np.float32).reshape(channels, -1).T
    return np.ascontiguousarray(img.transpose(1, 0, 2))


def normals_image(N, mask):
    """Utilities.cpp:277-298.  N: [4][npix] (or [3][npix]).  Returns float32 h x w x 3 RGB in [0, 1]."""
    N = np.asarray(N, np.float32).reshape(-1, np.count_nonzero(mask))
    v = np.stack([0.5 + 0.5 * N[0], 0.5 + 0.5 * N[1], 0.5 - 0.5 * N[2]]).clip(0.0, 1.0)
    img = _scatter(v, mask, 3)
    lo, hi = float(img.min()), float(img.max())                 # cv::normalize(.., 0, 1, CV_MINMAX) over all channels
    return (img - lo) / (hi - lo) if hi > lo else np.zeros_like(img)


def albedo_image(rho, mask):
    """Utilities.cpp:242-275.  rho: [3][npix].  Returns float32 h x w x 3 RGB in [0, 1]."""
    rho = np.asarray(rho, np.float32).reshape(3, -1)
    out = np.empty_like(rho)
    for c in range(3):
        mean = rho[c].sum(dtype=np.float32) / np.float32(rho[c].size)
        std = np.sqrt(np.float32(np.dot(rho[c], rho[c])) / np.float32(rho[c].size) - mean * mean)
        cap = np.float32(np.median(rho[c])) + np.float32(5) * std
        out[c] = np.minimum(cap, rho[c]).clip(0.0, 1.0)
    return _scatter(out, mask, 3)


def bone_colormap():
    """cv::COLORMAP_BONE as RGB uint8 [256][3]: MATLAB's bone = (7 gray + fliplr(hot)) / 8 (within one level of OpenCV's table)."""
    m = 256
    n1 = 3 * m // 8
    ramp = np.arange(1, n1 + 1) / n1
    hot = np.stack([np.concatenate([ramp, np.ones(m - n1)]),
                    np.concatenate([np.zeros(n1), ramp, np.ones(m - 2 * n1)]),
                    np.concatenate([np.zeros(2 * n1), np.arange(1, m - 2 * n1 + 1) / (m - 2 * n1)])], axis=1)
    gray = np.repeat(np.linspace(0.0, 1.0, m)[:, None], 3, axis=1)
    return np.round((7.0 * gray + hot[:, ::-1]) / 8.0 * 255.0).astype(np.uint8)


def depth_image(z, mask):
    """Utilities.cpp:300-320.  z: [npix].  Returns uint8 h x w x 3 RGB."""
    v = -np.asarray(z, np.float32).reshape(-1)
    lo, hi = float(v.min()), float(v.max())
    g = (v - lo) / (hi - lo) if hi > lo else np.zeros_like(v)
    idx = np.clip(np.rint(g * 255.0), 0, 255).astype(np.uint8)                  # convertTo(CV_8U) rounds to nearest
    rgb = bone_colormap()[idx].astype(np.float32).T                             # [3][npix]
    return _scatter(rgb, mask, 3).astype(np.uint8)


def to_u8(img):
    """float image in [0,1] -> uint8, rounded to nearest (what cv::imwrite's convertTo does)."""
    return np.clip(np.rint(np.asarray(img, np.float32) * 255.0), 0, 255).astype(np.uint8)


def write_png(path, img):
    """8-bit gray (h x w) or RGB (h x w x 3) PNG, filter 0, one zlib stream."""
    img = np.ascontiguousarray(img)
    if img.dtype != np.uint8:
        raise TypeError("write_png wants uint8")
    if img.ndim == 2:
        img = img[:, :, None]
    h, w, c = img.shape
    if c not in (1, 3):
        raise ValueError("1 or 3 channels")
    raw = np.concatenate([np.zeros((h, 1), np.uint8), img.reshape(h, w * c)], axis=1).tobytes()

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 0, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(chunk(b"IEND", b""))


def write_mat_vector(path, x, name="x"):
    """MAT v5 file with one single-precision column vector (what WRITE_MAT_FROM_DEVICE would have written, SRPS.cu:11-17)."""
    x = np.ascontiguousarray(np.asarray(x, "<f4").reshape(-1))

    def element(mi_type, payload):
        pad = (-len(payload)) % 8
        return struct.pack("<II", mi_type, len(payload)) + payload + b"\0" * pad

    body = (element(6, struct.pack("<II", 7, 0))                       # array flags: mxSINGLE_CLASS
            + element(5, struct.pack("<ii", x.size, 1))                # dimensions
            + element(1, name.encode("ascii"))                         # array name
            + element(7, x.tobytes()))                                 # real part, miSINGLE
    header = b"MATLAB 5.0 MAT-file, written by srmeetsps-cuda_b200".ljust(116) + b"\0" * 8 + struct.pack("<H", 0x0100) + b"IM"
    with open(path, "wb") as f:
        f.write(header)
        f.write(struct.pack("<II", 14, len(body)))                     # miMATRIX
        f.write(body)


def save_results(result, out_dir):
    """result: dict(z, rho, N, s, mask) as returned by SRPS.execute.  Writes the seven files listed in the module docstring."""
    os.makedirs(out_dir, exist_ok=True)
    mask = result["mask"]
    write_png(os.path.join(out_dir, "normals.png"), to_u8(normals_image(result["N"], mask)))
    write_png(os.path.join(out_dir, "albedo.png"), to_u8(albedo_image(result["rho"], mask)))
    write_png(os.path.join(out_dir, "depth.png"), depth_image(result["z"], mask))
    for name in ("s", "rho", "z", "N"):
        write_mat_vector(os.path.join(out_dir, name + ".mat"), result[name])
    return out_dir
