"""2+-GPU check of the strip partition (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_strip_check.py

Every rank owns a strip of image columns; ghost lines and CG scalars travel through the library's own
kernels over NVLink peer memory.  The strips' results are gathered and compared with a single-GPU
context on the same scene (only the fp64 summation order of the dot products differs)."""
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
          dict(h=48, w=72, sf=2, n=6, seed=13, mask_kind="random95"),     # 36 lines per rank: ghost line inside a partial tile
          dict(h=300, w=64, sf=4, n=9, seed=9, mask_kind="full"),
          dict(h=40, w=48, sf=1, n=6, seed=5, mask_kind="random95")]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    ok = True
    for cfg in SCENES:
        sc = o.synth_scene(cfg["h"], cfg["w"], cfg["sf"], cfg["n"], seed=cfg["seed"], mask_kind=cfg["mask_kind"])
        npix = sc["ops"]["npix"]
        ref = None
        if rank == 0:                       # single-GPU reference result first (not collective)
            with Context(sc["mask"], sc["n"], sc["sf"], sc["K"], device=local) as c1:
                c1.upload_state(sc["I"], sc["z"], sc["z0s"])
                ref = []
                for it in range(3):
                    e, k = c1.outer_iteration()
                    ref.append((e, k, c1.download("z"), c1.download("rho"), c1.download("s")))
        dist.barrier()
        ctx = make_strip_context(sc["mask"], sc["n"], sc["sf"], sc["K"], rank, world, local)
        j0, j1 = strip_bounds(cfg["w"], world)[rank]
        p0, p1, q0, q1 = local_ranges(sc["mask"], sc["sf"], j0, j1)
        assert (p0, p1, q0, q1) == ctx.pixel_range(), ((p0, p1, q0, q1), ctx.pixel_range())
        I = np.ascontiguousarray(sc["I"])
        ctx.upload_state_strided(I.reshape(-1)[p0:], npix, sc["z"][p0:p1], sc["z0s"][q0:q1])
        for it in range(3):
            e, k = ctx.outer_iteration()
            parts = [None] * world
            dist.all_gather_object(parts, (ctx.download("z"), ctx.download("rho"), ctx.download("s"), e, k))
            if rank == 0:
                z = np.concatenate([p[0] for p in parts]); rho = np.concatenate([p[1] for p in parts], axis=1)
                e_ref, k_ref, z_ref, rho_ref, s_ref = ref[it]
                same_scalars = all(p[3] == parts[0][3] and p[4] == parts[0][4] and np.array_equal(p[2], parts[0][2]) for p in parts)
                zr = rel_rmse(z, z_ref); rr = float(np.abs(rho - rho_ref).max()); er = abs(e - e_ref) / abs(e_ref)
                good = same_scalars and zr <= 2e-5 and rr <= 3e-4 and er <= 1e-4 and abs(k - k_ref) <= 1
                print(f"{cfg} it={it} world={world}: z relRMSE {zr:.2e} rho maxabs {rr:.2e} energy rel {er:.2e} cg {k}/{k_ref} "
                      f"ranks-agree {same_scalars} -> {'ok' if good else 'FAIL'}", flush=True)
                ok = ok and good
        ctx.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0:
        print("DIST_OK" if ok else "DIST_FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
