"""How far do two CG drivers of ONE GPU drift apart over many free-running outer iterations?  (Scale for the strip
soak of tests/dist_strip_check.py: the drivers differ only in the order of fp32 / fp64 operations.)
usage: python tests/drift_probe.py [iters]"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(cg, iters, out):
    os.environ["SRPS_CG"] = cg
    from srmeetsps_cuda_b200 import Context
    from srmeetsps_cuda_b200.synth import synth_scene_torch
    sc = synth_scene_torch(1024, 1024, 4, 8, 77, device="cuda:0", pin=False)
    zs = []
    with Context(sc["mask"], 8, 4, sc["K"]) as ctx:
        ctx.upload_state(sc["I"], sc["z"], sc["z0s"])
        for it in range(iters):
            ctx.outer_iteration()
            zs.append(ctx.download("z")); zs.append(ctx.download("rho").reshape(-1))
    np.savez(out, *zs)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        run(sys.argv[2], int(sys.argv[3]), sys.argv[4])
        sys.exit(0)
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    outs = {}
    for cg in ("fused", "graph", "persistent_fused"):
        out = f"/tmp/drift_{cg}.npz"
        subprocess.check_call([sys.executable, __file__, "--child", cg, str(iters), out])
        outs[cg] = np.load(out)
    ref = outs["fused"]
    for cg in ("graph", "persistent_fused"):
        d = outs[cg]
        row = []
        for it in range(iters):
            z0, z1 = ref[f"arr_{2 * it}"].astype(np.float64), d[f"arr_{2 * it}"].astype(np.float64)
            r0, r1 = ref[f"arr_{2 * it + 1}"], d[f"arr_{2 * it + 1}"]
            row.append(f"{np.sqrt(np.mean((z0 - z1) ** 2)) / np.sqrt(np.mean(z0 ** 2)):.1e}/{np.abs(r0 - r1).max():.1e}")
        print(f"1 GPU, {cg} vs fused, per outer iteration (z relRMSE / rho maxabs): " + " ".join(row))
