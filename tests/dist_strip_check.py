"""2+-GPU check of the strip partition (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_strip_check.py [--soak ITERS]

Every rank owns a strip of image columns; ghost lines and CG scalars travel through the library's own
kernels over NVLink peer memory.  The strips' results are gathered and compared with a single-GPU
context on the same scene (only the fp64 summation order of the dot products differs).

--soak ITERS: additionally run ITERS outer iterations of a 1024x1024x8 full-mask scene TWICE on the strips from the
same upload and require bit-identical results on every rank (a stale ghost line, a lost mailbox word or a race in
the in-kernel all-reduce shows up as a difference between two runs of the same program), then compare the end state
with the single-GPU run.  SRPS_CG selects the CG driver (fused | graph | persistent_fused; default: by size)."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import rel_rmse                      # noqa: E402
from oracle import srps_oracle as o               # noqa: E402
from srmeetsps_cuda_b200 import Context           # noqa: E402
from srmeetsps_cuda_b200.dist import local_ranges, make_strip_context, strip_bounds   # noqa: E402

SCENES = [dict(h=96, w=128, sf=2, n=6, seed=7, mask_kind="ellipse"),
          dict(h=64, w=96, sf=2, n=6, seed=12, mask_kind="random95"),
          dict(h=48, w=72, sf=2, n=6, seed=13, mask_kind="random95"),     # 36 lines per rank at 2 ranks: ghost line inside a partial tile
          dict(h=300, w=64, sf=4, n=9, seed=9, mask_kind="full"),
          dict(h=40, w=48, sf=1, n=6, seed=5, mask_kind="random95"),
          dict(h=40, w=48, sf=1, n=6, seed=5, mask_kind="random95", dark=0.1),   # early convergence: the fused CG's guard (deferred passes)
          dict(h=256, w=512, sf=4, n=8, seed=21, mask_kind="full")]


def make_scene(cfg):
    sc = o.synth_scene(cfg["h"], cfg["w"], cfg["sf"], cfg["n"], seed=cfg["seed"], mask_kind=cfg["mask_kind"])
    if "dark" in cfg:
        sc["I"] = (sc["I"] * np.float32(cfg["dark"])).astype(np.float32)
    return sc


def strip_upload(ctx, sc, rank, world):
    j0, j1 = strip_bounds(sc["mask"].shape[1], world)[rank]
    p0, p1, q0, q1 = local_ranges(sc["mask"], sc["sf"], j0, j1)
    assert (p0, p1, q0, q1) == ctx.pixel_range(), ((p0, p1, q0, q1), ctx.pixel_range())
    I = np.ascontiguousarray(sc["I"])
    ctx.upload_state_strided(I.reshape(-1)[p0:], sc["ops"]["npix"], sc["z"][p0:p1], sc["z0s"][q0:q1])


def check_scenes(rank, world, local):
    ok = True
    for cfg in SCENES:
        if cfg["w"] // 4 < world:
            continue
        sc = make_scene(cfg)
        ref = None
        if rank == 0:                       # single-GPU reference result first (not collective)
            with Context(sc["mask"], sc["n"], sc["sf"], sc["K"], device=local) as c1:
                c1.upload_state(sc["I"], sc["z"], sc["z0s"])
                ref = []
                for it in range(3):
                    e, k = c1.outer_iteration()
                    ref.append((e, k, c1.download("z"), c1.download("rho"), c1.download("s"), c1.timings()["cg_deferred"]))
        dist.barrier()
        ctx = make_strip_context(sc["mask"], sc["n"], sc["sf"], sc["K"], rank, world, local)
        strip_upload(ctx, sc, rank, world)
        for it in range(3):
            e, k = ctx.outer_iteration()
            parts = [None] * world
            dist.all_gather_object(parts, (ctx.download("z"), ctx.download("rho"), ctx.download("s"), e, k, ctx.timings()["cg_deferred"]))
            if rank == 0:
                z = np.concatenate([p[0] for p in parts]); rho = np.concatenate([p[1] for p in parts], axis=1)
                e_ref, k_ref, z_ref, rho_ref, s_ref, d_ref = ref[it]
                same_scalars = all(p[3] == parts[0][3] and p[4] == parts[0][4] and np.array_equal(p[2], parts[0][2]) for p in parts)
                zr = rel_rmse(z, z_ref); rr = float(np.abs(rho - rho_ref).max()); er = abs(e - e_ref) / abs(e_ref)
                good = same_scalars and np.all(np.isfinite(z)) and zr <= 2e-5 and rr <= 3e-4 and er <= 1e-4 and abs(k - k_ref) <= 1
                print(f"{cfg} it={it} world={world}: z relRMSE {zr:.2e} rho maxabs {rr:.2e} energy rel {er:.2e} cg {k}/{k_ref} "
                      f"deferred {parts[0][5]}/{d_ref} ranks-agree {same_scalars} -> {'ok' if good else 'FAIL'}", flush=True)
                ok = ok and bool(good)
        ctx.close()
        dist.barrier()
    return ok


def soak(rank, world, local, iters):
    """Two runs of `iters` outer iterations on the strips from the same upload must be bit-identical on every rank;
    the first iterations are also held to the single-GPU run.  (Far beyond the reference's <= 11 iterations the two
    partitions drift apart by amplified summation-order noise, as two single-GPU CG drivers do: reported, not judged.)"""
    from srmeetsps_cuda_b200.synth import synth_scene_torch
    h, w, sf, n, seed = 1024, 1024, 4, 8, 77
    keep = min(iters, 10)
    full = synth_scene_torch(h, w, sf, n, seed, device=f"cuda:{local}", pin=False)
    j0, j1 = strip_bounds(w, world)[rank]
    p0, p1, q0, q1 = j0 * h, j1 * h, (j0 // sf) * (h // sf), (j1 // sf) * (h // sf)
    runs, curve = [], []
    ctx = make_strip_context(full["mask"], n, sf, full["K"], rank, world, local)
    for rep in range(2):
        ctx.upload_state_strided(np.ascontiguousarray(full["I"]).reshape(-1)[p0:], h * w, full["z"][p0:p1], full["z0s"][q0:q1])
        es = []
        for it in range(iters):
            es.append(ctx.outer_iteration())
            if rep == 0 and it < keep:
                curve.append((ctx.download("z"), ctx.download("rho")))
        runs.append((es, ctx.download("z"), ctx.download("rho"), ctx.download("s")))
    ctx.close()
    same = (runs[0][0] == runs[1][0] and np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2])
            and np.array_equal(runs[0][3], runs[1][3]))
    parts = [None] * world
    dist.all_gather_object(parts, (same, curve, runs[1][1], runs[1][2], runs[1][0][-1]))
    ok = True
    if rank == 0:
        drift = []
        with Context(full["mask"], n, sf, full["K"], device=local) as c1:
            c1.upload_state(full["I"], full["z"], full["z0s"])
            for it in range(iters):
                e1 = c1.outer_iteration()
                if it < keep:
                    z = np.concatenate([p[1][it][0] for p in parts]); rho = np.concatenate([p[1][it][1] for p in parts], axis=1)
                    drift.append((rel_rmse(z, c1.download("z")), float(np.abs(rho - c1.download("rho")).max())))
            z1, rho1 = c1.download("z"), c1.download("rho")
        z = np.concatenate([p[2] for p in parts]); rho = np.concatenate([p[3] for p in parts], axis=1)
        all_same = all(p[0] for p in parts)
        zr = rel_rmse(z, z1); rr = float(np.abs(rho - rho1).max())
        er = abs(parts[0][4][0] - e1[0]) / abs(e1[0])
        # measured scale (tests/drift_probe.py, profiles/r2_drift_1gpu.log): two CG drivers on ONE GPU, which differ in
        # nothing but the order of floating-point operations, are 1e-4 / 3e-4 / 1e-3 / 6e-3 apart in albedo after
        # 3 / 6 / 7 / 9 free-running iterations of this scene, 2e-2 after 20 -- the loop amplifies round-off in the
        # pixels the data barely determine (the two-kernel driver on 4 GPUs reached 1.1e-3 at the sixth).  So: the first
        # three iterations sharp, five within the north-star bound (z 1e-4, albedo 1e-3), the rest reported.
        early = all(d[0] <= 1e-5 and d[1] <= 2e-4 for d in drift[:3]) and all(d[0] <= 1e-4 and d[1] <= 1e-3 for d in drift[:5])
        ok = all_same and early and bool(np.all(np.isfinite(z))) and er <= 1e-3
        print(f"soak world={world} iters={iters}: two strip runs bit-identical on every rank {all_same}; strips vs 1 GPU per iteration "
              f"(z relRMSE / rho maxabs): " + " ".join(f"{a:.1e}/{b:.1e}" for a, b in drift)
              + f"; after {iters}: {zr:.2e}/{rr:.2e} energy rel {er:.2e} -> {'ok' if ok else 'FAIL'}", flush=True)
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--soak", type=int, default=0)
    ap.add_argument("--skip-scenes", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        print(f"dist_strip_check: world={world} SRPS_CG={os.environ.get('SRPS_CG', '(default)')}", flush=True)
    ok = True
    if not args.skip_scenes:
        ok = check_scenes(rank, world, local) and ok
    if args.soak > 0:
        ok = soak(rank, world, local, args.soak) and ok
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0:
        print("DIST_OK" if int(flag.item()) == 1 else "DIST_FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
