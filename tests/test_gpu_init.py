"""N2: the device kernels of the one-shot depth pre-processing (csrc/srps_init.cuh, SRPS.cu:117-149) against
python cv2 -- the OpenCV calls the reference makes -- and against the host implementation of the C++ CLI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_depth_mean_and_flags_match_reference_kernel_semantics():
    from srmeetsps_cuda_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    n, frames = 48 * 64, 5
    z0 = (600 + 100 * rng.random((frames, n))).astype(np.float32)
    z0[rng.random((frames, n)) < 0.05] = 0.0
    mean = np.empty(n, np.float32); hole = np.empty(n, np.uint8)
    assert lib.srps_init_depth_mean(0, _p(z0), n, frames, _p(mean), _p(hole)) == 0
    acc = np.zeros(n, np.float32)
    for c in range(frames):                                   # same accumulation order as devicecalls.cu:100-108
        acc = (acc + np.where(z0[c] != 0, z0[c], np.float32(0))).astype(np.float32)
    assert np.array_equal(mean, acc / np.float32(frames))
    assert np.array_equal(hole, (z0 == 0).any(axis=0).astype(np.uint8))


@pytest.mark.parametrize("shape,sf", [((64, 48), 2), ((135, 240), 4), ((33, 57), 8)])
def test_bilateral_and_bicubic_match_cv2(shape, sf):
    import cv2
    from srmeetsps_cuda_b200 import _lib
    lib = _lib.load()
    rows, cols = shape
    rng = np.random.default_rng(1)
    yy, xx = np.mgrid[0:rows, 0:cols]
    depth = (700 + 60 * np.exp(-((xx - cols / 2) ** 2 + (yy - rows / 2) ** 2) / (0.1 * rows * cols)) + rng.standard_normal(shape)).astype(np.float32)
    zs = np.empty(rows * cols, np.float32); zf = np.empty(rows * sf * cols * sf, np.float32)
    assert lib.srps_init_depth_smooth_upsample(0, _p(depth), rows, cols, rows * sf, cols * sf, 2.0, 2.0, _p(zs), _p(zf)) == 0
    mx = float(depth.max())
    ref_s = cv2.bilateralFilter((depth / np.float32(mx)).astype(np.float32), -1, 2, 2) * np.float32(mx)        # SRPS.cu:137-140
    ref_f = cv2.resize(ref_s, (cols * sf, rows * sf), interpolation=cv2.INTER_CUBIC)                            # SRPS.cu:149
    assert np.abs(zs.reshape(rows, cols) - ref_s).max() <= 2e-6 * mx
    assert np.abs(zf.reshape(rows * sf, cols * sf) - ref_f).max() <= 3e-6 * mx


def test_cli_device_init_equals_host_init(tmp_path):
    """srps_cli --init=device (default: kernels + host Telea) and --init=host produce the same post-init snapshot."""
    from test_cpp_host import write_image_folder
    from srmeetsps_cuda_b200.snapshot import read_snapshot
    import srmeetsps_cuda_b200.build as b
    b.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "src", "host")], stdout=subprocess.DEVNULL)
    cli = os.path.join(ROOT, "src", "host", "srps_cli")
    folder = write_image_folder(str(tmp_path / "scene"), dropout=0.05, seed=9)
    snaps = {}
    for mode in ("device", "host"):
        out = str(tmp_path / f"{mode}.snap")
        res = subprocess.run([cli, "--dstype=images", f"--dsloc={folder}", "--init-only", f"--init={mode}", f"--dump-init={out}"],
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        snaps[mode] = read_snapshot(out)
    for k in ("I", "mask", "K"):
        assert np.array_equal(snaps["device"][k], snaps["host"][k])
    for k in ("z", "z0s"):
        assert np.abs(snaps["device"][k] - snaps["host"][k]).max() <= 2e-6 * np.abs(snaps["host"][k]).max(), k
