"""A/B of the CG drivers on one GPU: bit-identity of the results and per-pass times.
usage: python tests/ab_drivers.py [workload] [drivers...]   (bench.py workloads: 4k, 1080p, 1k)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(workload, cg, out):
    os.environ["SRPS_CG"] = cg
    import numpy as np
    import torch
    from bench import WORKLOADS
    from srmeetsps_cuda_b200 import Context
    from srmeetsps_cuda_b200.synth import synth_scene_torch
    h, w, sf, n, seed = WORKLOADS[workload]
    sc = synth_scene_torch(h, w, sf, n, seed, device="cuda:0")
    with Context(sc["mask"], n, sf, sc["K"]) as ctx:
        ctx.upload_state(sc["I"], sc["z"], sc["z0s"])
        for _ in range(3):
            ctx.outer_iteration()
        ts, cgs = [], []
        for _ in range(8):
            e, k = ctx.outer_iteration()
            t = ctx.timings()
            ts.append(t["ms_total"]); cgs.append(t["ms_depth_cg"])
        z = ctx.download("z")
        prof = ctx.profile_kernels(reps=30)
    np.save(out, z)
    print(json.dumps({"cg": cg, "ms_total": float(np.mean(ts)), "ms_cg": float(np.mean(cgs)), "k": k, "energy": e,
                      "pass_alone_ms": prof["cg_fused"], "driver": prof["cg_driver"]}))


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3], sys.argv[4])
        sys.exit(0)
    import numpy as np
    workload = sys.argv[1] if len(sys.argv) > 1 else "4k"
    drivers = sys.argv[2:] or ["fused", "fused_tma"]
    zs = {}
    for cg in drivers:
        out = f"/tmp/ab_{cg}.npy"
        subprocess.check_call([sys.executable, __file__, "--child", workload, cg, out])
        zs[cg] = np.load(out)
    for cg in drivers[1:]:
        print(f"{workload}: z of {cg} bit-identical to {drivers[0]}: {np.array_equal(zs[cg], zs[drivers[0]])}, max abs diff {np.abs(zs[cg] - zs[drivers[0]]).max():.3e}")
