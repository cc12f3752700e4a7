"""The drop-in seam at operator level: tests/adapter_replay.cu runs the reference's loop against the
reference's own function names and signatures (include/srps_devicecalls_adapter.h: raw device pointers, masked layouts,
cuBLAS / cuSPARSE handles and CSR operands accepted and ignored).  Its results must equal the context API's, bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from oracle import srps_oracle as o

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def replay(tmp_path_factory):
    import srmeetsps_cuda_b200.build as b
    b.build()
    exe = str(tmp_path_factory.mktemp("adapter") / "adapter_replay")
    lib = os.path.join(ROOT, "srmeetsps-cuda_b200")
    subprocess.check_call(["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "adapter_replay.cu"), "-o", exe, "-L", lib, "-lsrps_b200",
                           "-Xlinker", f"-rpath={lib}"])
    return exe


@pytest.mark.parametrize("mode", ["closed_form", "reference_cg"])
@pytest.mark.parametrize("cfg", [dict(h=96, w=128, sf=2, n=6, seed=7, mask_kind="ellipse"), dict(h=64, w=96, sf=4, n=5, seed=11, mask_kind="random")],
                         ids=["ellipse", "random"])
def test_reference_named_operators_equal_context_api(replay, tmp_path, cfg, mode):
    from srmeetsps_cuda_b200 import Context
    from srmeetsps_cuda_b200.snapshot import read_snapshot, write_snapshot
    sc = o.synth_scene(**cfg)
    snap = str(tmp_path / "in.snap")
    write_snapshot(snap, {"dims": np.array([sc["h"], sc["w"], sc["sf"]], np.int32), "K": np.asarray(sc["K"], np.float32),
                          "mask": (sc["mask"] != 0).astype(np.uint8).ravel(order="F"), "I": sc["I"], "z": sc["z"], "z0s": sc["z0s"]})
    out = str(tmp_path / "out.snap")
    res = subprocess.run([replay, snap, out, "3", "1" if mode == "reference_cg" else "0"], capture_output=True, text=True)
    assert res.returncode == 0 and "adapter_replay ok" in res.stdout, res.stdout + res.stderr
    r = read_snapshot(out)
    with Context(sc["mask"], sc["n"], sc["sf"], sc["K"], albedo_mode=mode) as ctx:
        ctx.upload_state(sc["I"], sc["z"], sc["z0s"])
        e = [ctx.outer_iteration()[0] for _ in range(3)]
        assert np.array_equal(np.asarray(e, np.float32), r["energy"])
        for name in ("z", "rho", "s", "N"):
            assert np.array_equal(ctx.download(name), r[name]), name
