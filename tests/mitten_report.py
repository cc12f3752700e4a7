"""Mitten (BASELINE configs 1/2 data, post-init snapshot): per-iteration timing of the CUDA path and its parity
against the reference-CUDA goldens.  Prints one JSON line; run on a GPU box: python tests/mitten_report.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLDEN, rel_rmse                         # noqa: E402
from oracle import datasets as ds                              # noqa: E402
from srmeetsps_cuda_b200 import Context                        # noqa: E402

sc = ds.scene_from_snapshot(np.load(os.path.join(GOLDEN, "mitten_init.npz")))
g = np.load(os.path.join(GOLDEN, "ref_mitten.npz"))
stride = int(g["stride"])
out = {"scene": "Mitten", "npix": int(sc["ops"]["npix"]), "n": int(sc["n"]), "modes": {}}
for mode in ("closed_form", "reference_cg"):
    rows = []
    with Context(sc["mask"], sc["n"], sc["sf"], sc["K"], albedo_mode=mode) as ctx:
        ctx.upload_state(sc["I"], sc["z"], sc["z0s"])
        for it in range(1, 4):
            e, k = ctx.outer_iteration()
            t = ctx.timings()
            rows.append({"iteration": it, "energy": e, "energy_ref": float(g[f"energy_{it}"][0]), "cg_iters": k,
                         "z_rel_rmse_vs_ref": rel_rmse(ctx.download("z"), g[f"z_{it}"]),
                         "rho_maxabs_vs_ref": float(np.abs(ctx.download("rho")[:, ::stride] - g[f"rho_{it}"]).max()),
                         "ms_total": t["ms_total"], "ms_lighting": t["ms_lighting"], "ms_albedo": t["ms_albedo"], "ms_depth": t["ms_depth"]})
        ms = []
        for _ in range(10):
            ctx.outer_iteration(); ms.append(ctx.timings()["ms_total"])
    out["modes"][mode] = {"iterations": rows, "ms_per_outer_iteration_steady": float(np.median(ms))}
with open(os.path.join(GOLDEN, "ref_replay_log.json")) as fh:
    ref = json.load(fh)["mitten"]
out["reference_cuda_ms_per_outer_iteration"] = [r["ms_total"] for r in ref if "ms_total" in r]
print(json.dumps(out))
