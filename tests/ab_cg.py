"""A/B helper: the CG drivers (SRPS_CG=graph|fused|persistent) on one scene -- per-iteration timing and the agreement of
z / energy between them.  python tests/ab_cg.py [h w n]   (default 4096 4096 32; SRPS_LIB selects an alternative build)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srmeetsps_cuda_b200 import Context                       # noqa: E402
from srmeetsps_cuda_b200.synth import synth_scene_torch       # noqa: E402

h, w, n = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 4096, 32)
modes = sys.argv[4].split(",") if len(sys.argv) >= 5 else ["graph", "fused"]
sc = synth_scene_torch(h, w, 4, n, 2000)
out = {"lib": os.environ.get("SRPS_LIB", "default").split("/")[-1], "scene": [h, w, n]}
zs = {}
for mode in modes:
    os.environ["SRPS_CG"] = mode
    with Context(sc["mask"], n, 4, sc["K"]) as ctx:
        ctx.upload_state(sc["I"], sc["z"], sc["z0s"])
        es = []
        for _ in range(3):
            e, k = ctx.outer_iteration()
            es.append(e)
        zs[mode] = ctx.download("z")
        ms, cg = [], []
        for _ in range(5):
            ctx.outer_iteration()
            t = ctx.timings()
            ms.append(t["ms_total"]); cg.append(t["ms_depth_cg"])
        out[mode] = {"energies": es, "cg_iters": k, "ms_total": float(np.median(ms)), "ms_depth_cg": float(np.median(cg))}
ref = zs[modes[0]]
for mode in modes[1:]:
    d = zs[mode] - ref
    out[mode]["z_rel_rmse_vs_" + modes[0]] = float(np.sqrt(np.mean(d.astype(np.float64) ** 2)) / np.sqrt(np.mean(ref.astype(np.float64) ** 2)))
print(json.dumps(out))
