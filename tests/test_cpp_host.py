"""C++ host (src/host: reference-compatible SRPS class, DataHandlers, CLI) -- CPU parts: the loaders are
bit-exact against the python/cv2 mirror of the reference loaders, the OpenCV-free init stays close to cv2's."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "src", "host", "srps_cli")


@pytest.fixture(scope="module")
def cli():
    import srmeetsps_cuda_b200.build as b
    b.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "src", "host")], stdout=subprocess.DEVNULL)
    return CLI


def write_image_folder(root, h=48, w=64, sf=2, n=4, seed=0, dropout=0.0):
    import cv2
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "RGB")); os.makedirs(os.path.join(root, "Depth"))
    yy, xx = np.mgrid[0:h, 0:w]
    mask = (((xx - w / 2) / (0.4 * w)) ** 2 + ((yy - h / 2) / (0.45 * h)) ** 2 < 1).astype(np.uint8) * 255
    cv2.imwrite(os.path.join(root, "mask.png"), mask)
    for i in range(n):
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        cv2.imwrite(os.path.join(root, "RGB", f"img_{i:02d}.png"), img)
    hs, ws = h // sf, w // sf
    for i in range(3):
        d = (30000 + 8000 * np.sin(np.arange(ws) / 7.0)[None, :] + 3000 * np.cos(np.arange(hs) / 5.0)[:, None]
             + rng.integers(0, 200, size=(hs, ws))).astype(np.uint16)
        if dropout > 0:
            d[rng.random((hs, ws)) < dropout] = 0
        cv2.imwrite(os.path.join(root, "Depth", f"d_{i:02d}.png"), d)
    with open(os.path.join(root, "K.txt"), "w") as fh:
        fh.write("80.5,0,31.5\n0,80.5,23.5\n0,0,1\n%d,0,9870" % sf)     # min_z = 0: a zero sample stays a zero depth
    return root


def run_init(cli, dstype, dsloc, out):
    res = subprocess.run([cli, f"--dstype={dstype}", f"--dsloc={dsloc}", "--init-only", "--init=host", f"--dump-init={out}"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lines = res.stdout.splitlines()
    assert lines[:2] == ["Small mask calculation", "Mean of depth values"] and lines[-1] == "Done!"   # SRPS.cu:106,119,337
    from srmeetsps_cuda_b200.snapshot import read_snapshot
    return read_snapshot(out)


def python_init(folder):
    from srmeetsps_cuda_b200 import ImageDataHandler, preprocess_depth
    dh = ImageDataHandler().loadDataFromImages(folder)
    h, w, sf = dh.I_h, dh.I_w, int(dh.sf)
    mask = dh.mask != 0
    zs, z_full = preprocess_depth(dh.z0, h, w, sf)
    mflat = mask.ravel(order="F")
    lr = mask.reshape(h // sf, sf, w // sf, sf).all(axis=(1, 3)).ravel(order="F")
    I = dh.I.transpose(0, 1, 3, 2).reshape(dh.I_n, dh.I_c, h * w)[:, :, mflat]
    return dh, I, z_full[mflat], zs[lr], mflat


def test_cli_help_without_dsloc(cli):
    res = subprocess.run([cli], capture_output=True, text=True)
    assert res.returncode == 0 and "--dsloc" in res.stdout and "--dstype" in res.stdout      # Main.cpp:23-26


def test_image_loader_and_init_match_python_mirror(cli, tmp_path):
    folder = write_image_folder(str(tmp_path / "scene"))
    snap = run_init(cli, "images", folder, str(tmp_path / "init.snap"))
    dh, I, z, z0s, mflat = python_init(folder)
    assert list(snap["dims"]) == [dh.I_h, dh.I_w, int(dh.sf)]
    assert np.array_equal(snap["mask"].astype(bool), mflat)
    assert np.array_equal(snap["K"], dh.K)
    assert np.array_equal(snap["I"], I)                       # PNG decode + /255 + channel order: bit-exact
    assert snap["z0s"].shape == z0s.shape and snap["z"].shape == z.shape
    # no zero depth samples -> no inpainting: bilateral + bicubic only
    assert np.abs(snap["z0s"] - z0s).max() <= 2e-6 * np.abs(z0s).max()
    assert np.abs(snap["z"] - z).max() <= 2e-6 * np.abs(z).max()


def test_init_with_depth_dropout_matches_cv2(cli, tmp_path):
    """Telea inpainting in OpenCV's parametrisation (Preprocess.cpp) + bilateral + bicubic against python cv2:
    agreement to a few ulp even with 2 % / 10 % of the depth samples missing (holes on the image border included)."""
    for dropout, seed in ((0.02, 3), (0.10, 4)):
        folder = write_image_folder(str(tmp_path / f"scene{seed}"), dropout=dropout, seed=seed)
        snap = run_init(cli, "images", folder, str(tmp_path / "init.snap"))
        dh, I, z, z0s, mflat = python_init(folder)
        assert np.abs(snap["z0s"] - z0s).max() <= 2e-6 * np.abs(z0s).max()
        assert np.abs(snap["z"] - z).max() <= 2e-6 * np.abs(z).max()


def test_mat_loader(cli, tmp_path):
    from scipy.io import savemat
    folder = write_image_folder(str(tmp_path / "scene"))
    from srmeetsps_cuda_b200 import ImageDataHandler
    dh = ImageDataHandler().loadDataFromImages(folder)
    for compress in (False, True):
        path = str(tmp_path / f"scene_{int(compress)}.mat")
        savemat(path, {"I": dh.I.transpose(2, 3, 1, 0).astype(np.float64), "K": dh.K.reshape(3, 3, order="F").astype(np.float64),
                       "mask": (dh.mask != 0).astype(np.uint8), "sf": float(dh.sf), "z0": dh.z0.transpose(1, 2, 0).astype(np.float64)},
                do_compression=compress)
        a = run_init(cli, "matlab", path, str(tmp_path / "a.snap"))
        b = run_init(cli, "images", folder, str(tmp_path / "b.snap"))
        for k in ("dims", "mask", "K", "I", "z", "z0s"):
            assert np.array_equal(a[k], b[k]), k
    res = subprocess.run([cli, "--dstype=matlab", f"--dsloc={tmp_path}/missing.mat"], capture_output=True, text=True)
    assert res.returncode != 0 and "Error opening MAT file" in res.stderr                    # Utilities.cpp:165-168


@pytest.mark.skipif(not os.path.isdir("/root/reference/dataset/Images/Mitten"), reason="reference dataset not on this box")
def test_mitten_loaders_bit_exact(cli, tmp_path):
    snap = run_init(cli, "images", "/root/reference/dataset/Images/Mitten", str(tmp_path / "m.snap"))
    ref = np.load(os.path.join(ROOT, "tests", "golden", "mitten_init.npz"))
    assert np.array_equal(snap["mask"], ref["mask"].ravel(order="F"))
    assert np.array_equal(snap["I"], (ref["I8"].astype(np.float32) / np.float32(255)))
    d = snap["z"] - ref["z"]
    assert np.sqrt((d ** 2).mean()) < 5e-4 and np.abs(d).max() < 2e-2       # z in [535, 570]; measured: rms 9e-5, max 3e-3


@pytest.mark.gpu
def test_cli_loop_matches_python_host(cli, tmp_path, mitten_scene):
    """The C++ SRPS::execute and the python Context drive the same library: identical energies."""
    from srmeetsps_cuda_b200 import Context
    from srmeetsps_cuda_b200.snapshot import read_snapshot, write_snapshot
    sc = mitten_scene
    snap_in = str(tmp_path / "mitten.snap")
    write_snapshot(snap_in, {"dims": np.array([sc["h"], sc["w"], sc["sf"]], np.int32), "K": np.asarray(sc["K"], np.float32),
                             "mask": (sc["mask"] != 0).astype(np.uint8).ravel(order="F"), "I": sc["I"], "z": sc["z"], "z0s": sc["z0s"]})
    out = str(tmp_path / "res.snap")
    files = str(tmp_path / "files")
    res = subprocess.run([cli, "--dstype=snapshot", f"--dsloc={snap_in}", "--iters=3", f"--out={out}", f"--outdir={files}"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr + res.stdout
    assert sorted(os.listdir(files)) == ["N.mat", "albedo.png", "depth.png", "normals.png", "rho.mat", "s.mat", "z.mat"]
    assert "Iteration 03 summary" in res.stdout and "Lightning Estimation" in res.stdout          # SRPS.cu:283,303
    r = read_snapshot(out)
    with Context(sc["mask"], sc["n"], sc["sf"], sc["K"]) as ctx:
        ctx.upload_state(sc["I"], sc["z"], sc["z0s"])
        e = [ctx.outer_iteration()[0] for _ in range(3)]
        assert np.array_equal(np.asarray(e, np.float32), r["energy"])
        assert np.array_equal(ctx.download("z"), r["z"])


def test_result_files_match_python_output(cli, tmp_path):
    """N4: `srps_cli --render=result.snap --outdir=DIR` (the files `--outdir` writes after a run) against
    srmeetsps_cuda_b200.output on the same result: MAT vectors and the normals / depth PNGs byte-identical, the albedo
    PNG within one level (its cap uses an fp32 mean whose summation order differs)."""
    import cv2
    import scipy.io
    from srmeetsps_cuda_b200 import output
    from srmeetsps_cuda_b200.snapshot import write_snapshot
    from test_output import _scene
    r = _scene(h=40, w=56, seed=11)
    h, w = r["mask"].shape
    snap = str(tmp_path / "result.snap")
    write_snapshot(snap, dict(hw=np.array([h, w], np.int32), mask=np.ascontiguousarray(r["mask"].T.astype(np.uint8)),
                              z=r["z"], rho=r["rho"], N=r["N"], s=r["s"], energy=np.zeros(1, np.float32)))
    d_cpp, d_py = str(tmp_path / "cpp"), str(tmp_path / "py")
    res = subprocess.run([cli, f"--render={snap}", f"--outdir={d_cpp}"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    output.save_results(r, d_py)
    for name in ("s", "rho", "z", "N"):
        a = scipy.io.loadmat(os.path.join(d_cpp, name + ".mat"))["x"]
        b = scipy.io.loadmat(os.path.join(d_py, name + ".mat"))["x"]
        assert a.dtype == np.float32 and np.array_equal(a, b)
    for name, tol in (("normals", 0), ("depth", 0), ("albedo", 1)):
        a = cv2.imread(os.path.join(d_cpp, name + ".png"), cv2.IMREAD_COLOR)
        b = cv2.imread(os.path.join(d_py, name + ".png"), cv2.IMREAD_COLOR)
        assert a is not None and a.shape == b.shape == (h, w, 3)
        assert np.abs(a.astype(int) - b.astype(int)).max() <= tol, name


def test_images_to_mat_round_trip(tmp_path):
    """BASELINE config 1 substitute: the image folder re-encoded as MAT v5 loads to the same DataHandler (python host)."""
    from srmeetsps_cuda_b200 import ImageDataHandler, MatFileDataHandler
    from srmeetsps_cuda_b200.images_to_mat import images_to_mat
    folder = write_image_folder(str(tmp_path / "scene"))
    path = images_to_mat(folder, str(tmp_path / "scene.mat"))
    a = ImageDataHandler().loadDataFromImages(folder)
    b = MatFileDataHandler().loadDataFromMatFiles(path)
    assert (a.I_h, a.I_w, a.I_c, a.I_n, int(a.sf), a.z0.shape) == (b.I_h, b.I_w, b.I_c, b.I_n, int(b.sf), b.z0.shape)
    assert np.array_equal(a.I, b.I) and np.array_equal(a.K, b.K) and np.array_equal(a.z0, b.z0)
    assert np.array_equal(a.mask != 0, b.mask != 0)


def _write_interlaced_png(path, img):
    """8-bit gray or RGB Adam7-interlaced PNG (neither cv2 nor Pillow writes one)."""
    import struct
    import zlib
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else 3
    raw = b""
    for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
        sub = img[y0::dy, x0::dx]
        if sub.size:
            raw += b"".join(b"\0" + sub[r].tobytes() for r in range(sub.shape[0]))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 0, 0, 0, 1))
                + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


def test_png_flavours_load_like_cv_imread(cli, tmp_path):
    """N3: the zlib PNG reader against cv::imread semantics (python cv2 on the same files): palette, RGBA, gray and
    Adam7-interlaced images in RGB/, 1-bit / palette / colour masks."""
    import cv2
    from PIL import Image
    folder = write_image_folder(str(tmp_path / "scene"), n=4, seed=5)
    rgb = sorted(os.listdir(os.path.join(folder, "RGB")))
    imgs = [cv2.imread(os.path.join(folder, "RGB", f))[:, :, ::-1] for f in rgb]
    Image.fromarray(imgs[0]).quantize(32).save(os.path.join(folder, "RGB", rgb[0]))                     # palette, 8 bit
    Image.fromarray(np.dstack([imgs[1], np.full(imgs[1].shape[:2], 77, np.uint8)])).save(os.path.join(folder, "RGB", rgb[1]))   # RGBA
    Image.fromarray(imgs[2][:, :, 1]).save(os.path.join(folder, "RGB", rgb[2]))                          # gray
    _write_interlaced_png(os.path.join(folder, "RGB", rgb[3]), imgs[3])                                  # Adam7
    mask = cv2.imread(os.path.join(folder, "mask.png"), cv2.IMREAD_GRAYSCALE)
    for kind in ("bit1", "palette4", "colour", "interlaced"):
        mp = os.path.join(folder, "mask.png")
        if kind == "bit1":
            Image.fromarray(mask > 0).save(mp)
        elif kind == "palette4":
            Image.fromarray(np.dstack([mask, mask, mask])).quantize(4).save(mp, bits=2)
        elif kind == "colour":
            Image.fromarray(np.dstack([mask, mask, mask])).save(mp)
        else:
            _write_interlaced_png(mp, mask)
        snap = run_init(cli, "images", folder, str(tmp_path / "init.snap"))
        dh, I, z, z0s, mflat = python_init(folder)
        assert list(snap["dims"]) == [dh.I_h, dh.I_w, int(dh.sf)], kind
        assert np.array_equal(snap["mask"].astype(bool), mflat), kind
        assert np.array_equal(snap["I"], I), kind


def test_truncated_and_corrupt_datasets_fail_cleanly(cli, tmp_path):
    """Sizes stored in a file are not trusted: a truncated MAT file, a MAT file whose dims disagree with its data and a
    PNG with an absurd header end in the reference's error exit (Utilities.cpp:37-40, 165-168), not in a crash."""
    import struct
    import zlib
    from scipy.io import savemat
    folder = write_image_folder(str(tmp_path / "scene"))
    from srmeetsps_cuda_b200 import ImageDataHandler
    dh = ImageDataHandler().loadDataFromImages(folder)
    good = str(tmp_path / "good.mat")
    payload = {"I": dh.I.transpose(2, 3, 1, 0).astype(np.float64), "K": dh.K.reshape(3, 3, order="F").astype(np.float64),
               "mask": (dh.mask != 0).astype(np.uint8), "sf": float(dh.sf), "z0": dh.z0.transpose(1, 2, 0).astype(np.float64)}
    savemat(good, payload, do_compression=False)
    raw = open(good, "rb").read()
    cases = {"cut_half.mat": raw[: len(raw) // 2], "cut_tag.mat": raw[:128 + 12], "cut_tail.mat": raw[:-9]}
    small_mask = dict(payload, mask=payload["mask"][:-1])             # mask is not h x w
    savemat(str(tmp_path / "badmask.mat"), small_mask, do_compression=True)
    for name, blob in cases.items():
        open(str(tmp_path / name), "wb").write(blob)
    for name in list(cases) + ["badmask.mat"]:
        res = subprocess.run([cli, "--dstype=matlab", f"--dsloc={tmp_path}/{name}", "--init-only", "--init=host"], capture_output=True, text=True)
        # 134 = the reference's escaped std::runtime_error (Main.cpp has no handler), returned deliberately; a signal
        # (negative code: SIGSEGV / SIGABRT from a wild read or a failed huge allocation) is the failure this guards against
        assert res.returncode == 134 and "MAT file" in res.stderr, (name, res.returncode, res.stderr[-300:])

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    huge = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 0x7fffffff, 0x7fffffff, 8, 0, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0")) + chunk(b"IEND", b"")
    open(os.path.join(folder, "mask.png"), "wb").write(huge)
    res = subprocess.run([cli, "--dstype=images", f"--dsloc={folder}", "--init-only", "--init=host"], capture_output=True, text=True)
    assert res.returncode == 134 and "implausible PNG size" in res.stderr, (res.returncode, res.stderr[-300:])
