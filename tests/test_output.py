"""Result dumps and renderings against the reference's OpenCV recipe (Utilities.cpp:242-320), restated
here with cv2 calls exactly as the reference makes them."""
import os

import cv2
import numpy as np
import scipy.io

from srmeetsps_cuda_b200 import output


def _scene(h=48, w=64, seed=3):
    rng = np.random.default_rng(seed)
    ii, jj = np.mgrid[0:h, 0:w]
    mask = (((ii - h / 2) / (0.45 * h)) ** 2 + ((jj - w / 2) / (0.45 * w)) ** 2 < 1).astype(np.uint8)
    npix = int(mask.sum())
    N = rng.normal(size=(3, npix)).astype(np.float32)
    N /= np.linalg.norm(N, axis=0, keepdims=True)
    N = np.concatenate([N, np.ones((1, npix), np.float32)])
    rho = np.abs(rng.normal(0.5, 0.2, size=(3, npix))).astype(np.float32)
    rho[:, :5] = 7.0                                               # outliers: capped at median + 5 sigma, then at 1
    z = (700 + 30 * rng.random(npix)).astype(np.float32)
    s = rng.normal(size=(6, 3, 4)).astype(np.float32)
    return dict(mask=mask, N=N, rho=rho, z=z, s=s)


def _imask(mask):
    return np.flatnonzero(mask.ravel(order="F"))                   # SRPS.cu:157-162


def test_normals_rendering_matches_the_opencv_recipe():
    r = _scene()
    h, w = r["mask"].shape
    im = _imask(r["mask"])
    ref = np.zeros((h, w, 3), np.float32)                          # BGR, as the reference fills it
    vals = np.stack([0.5 + 0.5 * r["N"][0], 0.5 + 0.5 * r["N"][1], 0.5 - 0.5 * r["N"][2]]).clip(0, 1)
    ref[im % h, im // h] = vals[::-1].T
    ref = cv2.normalize(ref, None, 0.0, 1.0, cv2.NORM_MINMAX)
    got = output.normals_image(r["N"], r["mask"])
    assert np.abs(got[:, :, ::-1] - ref).max() < 1e-6


def test_albedo_rendering_caps_at_median_plus_five_sigma():
    r = _scene()
    h, w = r["mask"].shape
    im = _imask(r["mask"])
    got = output.albedo_image(r["rho"], r["mask"])
    for c in range(3):
        x = r["rho"][c].astype(np.float64)
        cap = np.median(x) + 5 * np.sqrt((x * x).mean() - x.mean() ** 2)
        want = np.clip(np.minimum(cap, x), 0, 1)
        assert np.abs(got[im % h, im // h, c] - want).max() < 1e-5
    assert got[r["mask"] == 0].max() == 0


def test_depth_rendering_matches_bone_colormap_within_one_level():
    r = _scene()
    h, w = r["mask"].shape
    im = _imask(r["mask"])
    zm = (-r["z"]).reshape(-1, 1)
    zm = cv2.normalize(zm, None, 0, 1, cv2.NORM_MINMAX) * 255.0
    col = cv2.applyColorMap(zm.astype(np.float32).round().astype(np.uint8), cv2.COLORMAP_BONE).reshape(-1, 3)   # BGR
    got = output.depth_image(r["z"], r["mask"])
    assert np.abs(got[im % h, im // h].astype(int) - col[:, ::-1].astype(int)).max() <= 1
    lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(-1, 1), cv2.COLORMAP_BONE).reshape(256, 3)[:, ::-1]
    assert np.abs(output.bone_colormap().astype(int) - lut.astype(int)).max() <= 1


def test_files_round_trip(tmp_path):
    r = _scene()
    d = output.save_results(r, str(tmp_path / "out"))
    assert sorted(os.listdir(d)) == ["N.mat", "albedo.png", "depth.png", "normals.png", "rho.mat", "s.mat", "z.mat"]
    for name in ("s", "rho", "z", "N"):
        x = scipy.io.loadmat(os.path.join(d, name + ".mat"))["x"]
        assert x.dtype == np.float32 and x.shape == (r[name].size, 1)
        assert np.array_equal(x[:, 0], r[name].reshape(-1))
    for name, img in (("normals", output.to_u8(output.normals_image(r["N"], r["mask"]))),
                      ("albedo", output.to_u8(output.albedo_image(r["rho"], r["mask"]))),
                      ("depth", output.depth_image(r["z"], r["mask"]))):
        back = cv2.imread(os.path.join(d, name + ".png"), cv2.IMREAD_COLOR)
        assert back is not None and np.array_equal(back[:, :, ::-1], img)
