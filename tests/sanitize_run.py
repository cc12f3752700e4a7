"""Workload for compute-sanitizer (memcheck / racecheck / initcheck): every CG driver, both albedo modes, the 8-bit stack, a
partial last group, early convergence with deferred passes, the device-pointer operators and the depth pre-processing
kernels on small scenes.   compute-sanitizer --tool memcheck python tests/sanitize_run.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import srps_oracle as o            # noqa: E402
from srmeetsps_cuda_b200 import Context, _lib  # noqa: E402

SCENES = [dict(h=64, w=96, sf=2, n=6, seed=12, mask_kind="random95"), dict(h=300, w=40, sf=4, n=17, seed=9, mask_kind="ellipse"),
          dict(h=48, w=272, sf=8, n=6, seed=6, mask_kind="ellipse"), dict(h=40, w=24, sf=1, n=6, seed=5, mask_kind="random95"),
          dict(h=64, w=50, sf=2, n=5, seed=14, mask_kind="full"), dict(h=40, w=48, sf=1, n=6, seed=5, mask_kind="random95", dark=0.1)]
DRIVERS = (("strip", "persistent_fused"), ("strip", "persistent"), ("strip", "graph"), ("strip", "fused"), ("strip", "fused_tma"), ("tile", "graph"))
if os.environ.get("SRPS_SAN_DRIVERS"):          # e.g. SRPS_SAN_DRIVERS=persistent_fused,fused: only these CG drivers
    DRIVERS = tuple(d for d in DRIVERS if d[1] in os.environ["SRPS_SAN_DRIVERS"].split(",") and d[0] == "strip")
for cfg in SCENES:
    sc = o.synth_scene(cfg["h"], cfg["w"], cfg["sf"], cfg["n"], seed=cfg["seed"], mask_kind=cfg["mask_kind"])
    if "dark" in cfg:
        sc["I"] = (sc["I"] * np.float32(cfg["dark"])).astype(np.float32)
    for mode in ("closed_form", "reference_cg"):
        for stencil, cg in DRIVERS:
            os.environ["SRPS_STENCIL"] = stencil
            os.environ["SRPS_CG"] = cg
            with Context(sc["mask"], sc["n"], sc["sf"], sc["K"], albedo_mode=mode, cg_max_iter=5) as ctx:
                ctx.upload_state(sc["I"], sc["z"], sc["z0s"])
                e, k = ctx.outer_iteration(); e, k = ctx.outer_iteration()
                if mode == "closed_form":
                    ctx.run(fixed_iters=2)
                ctx.apply_depth_operator(np.ones(ctx.npix, np.float32))
                for name in ("z", "rho", "N", "s", "dz", "z0s"):
                    ctx.download(name)
    # 8-bit stack
    os.environ["SRPS_CG"] = "fused"
    with Context(sc["mask"], sc["n"], sc["sf"], sc["K"]) as ctx:
        ctx.upload_images_u8(np.clip(np.rint(sc["I"] * 255), 0, 255).astype(np.uint8))
        ctx.upload_state(None, sc["z"], sc["z0s"])
        ctx.outer_iteration()
    print(cfg, "ok", e, k, flush=True)
# depth pre-processing kernels
lib = _lib.load()
rng = np.random.default_rng(0)
z0 = (600 + 100 * rng.random((3, 24 * 32))).astype(np.float32)
z0[0, :5] = 0
mean = np.empty(24 * 32, np.float32); hole = np.empty(24 * 32, np.uint8)
assert lib.srps_init_depth_mean(0, z0.ctypes.data_as(C.c_void_p), 24 * 32, 3, mean.ctypes.data_as(C.c_void_p), hole.ctypes.data_as(C.c_void_p)) == 0
zs = np.empty(24 * 32, np.float32); zf = np.empty(24 * 32 * 4, np.float32)
assert lib.srps_init_depth_smooth_upsample(0, mean.ctypes.data_as(C.c_void_p), 32, 24, 64, 48, 2.0, 2.0, zs.ctypes.data_as(C.c_void_p),
                                           zf.ctypes.data_as(C.c_void_p)) == 0
print("init kernels ok", flush=True)
