"""Debug aid: CG scalars per pass (SRPS_TRACE=1) of the early-convergence scene for one driver.  usage: python tests/trace_guard.py <driver>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SRPS_CG"] = sys.argv[1]
os.environ["SRPS_TRACE"] = "1"
from oracle import srps_oracle as o            # noqa: E402
from srmeetsps_cuda_b200 import Context        # noqa: E402

sc = o.synth_scene(40, 48, 1, 6, seed=5, mask_kind="random95")
sc["I"] = (sc["I"] * np.float32(0.1)).astype(np.float32)
with Context(sc["mask"], sc["n"], sc["sf"], sc["K"]) as ctx:
    ctx.upload_state(sc["I"], sc["z"], sc["z0s"])
    e, k = ctx.outer_iteration()
    print("energy", e, "k", k, "deferred", ctx.timings()["cg_deferred"])
