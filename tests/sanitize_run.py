import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import srps_oracle as o
from srmeetsps_cuda_b200 import Context
for cfg in [dict(h=64, w=96, sf=2, n=6, seed=12, mask_kind="random95"), dict(h=300, w=40, sf=4, n=17, seed=9, mask_kind="ellipse"),
            dict(h=48, w=272, sf=8, n=6, seed=6, mask_kind="ellipse"), dict(h=40, w=24, sf=1, n=6, seed=5, mask_kind="random95")]:
    sc = o.synth_scene(cfg["h"], cfg["w"], cfg["sf"], cfg["n"], seed=cfg["seed"], mask_kind=cfg["mask_kind"])
    for mode in ("closed_form", "reference_cg"):
        for stencil, cg in (("strip", "persistent_fused"), ("strip", "persistent"), ("strip", "graph"), ("strip", "fused"), ("tile", "graph")):
            os.environ["SRPS_STENCIL"] = stencil
            os.environ["SRPS_CG"] = cg
            with Context(sc["mask"], sc["n"], sc["sf"], sc["K"], albedo_mode=mode, cg_max_iter=5) as ctx:
                ctx.upload_state(sc["I"], sc["z"], sc["z0s"])
                e, k = ctx.outer_iteration(); e, k = ctx.outer_iteration()
                ctx.apply_depth_operator(np.ones(ctx.npix, np.float32))
                for name in ("z", "rho", "N", "s", "dz", "z0s"):
                    ctx.download(name)
    print(cfg, "ok", e, k, flush=True)
