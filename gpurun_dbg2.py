import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from oracle import srps_oracle as o
from srmeetsps_cuda_b200 import Context
from srmeetsps_cuda_b200.dist import local_ranges, make_strip_context, strip_bounds
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
cfg=dict(h=40, w=48, sf=1, n=6, seed=5, mask_kind="random95")
sc=o.synth_scene(cfg["h"], cfg["w"], cfg["sf"], cfg["n"], seed=cfg["seed"], mask_kind=cfg["mask_kind"])
if rank==0:
    with Context(sc["mask"], sc["n"], sc["sf"], sc["K"], device=local) as c1:
        c1.upload_state(sc["I"], sc["z"], sc["z0s"])
        for it in range(3):
            print('single it', it, c1.outer_iteration(), file=sys.stderr, flush=True)
dist.barrier()
ctx=make_strip_context(sc["mask"], sc["n"], sc["sf"], sc["K"], rank, world, local)
j0,j1=strip_bounds(cfg["w"], world)[rank]; p0,p1,q0,q1=local_ranges(sc["mask"], sc["sf"], j0, j1)
I=np.ascontiguousarray(sc["I"]); ctx.upload_state_strided(I.reshape(-1)[p0:], sc["ops"]["npix"], sc["z"][p0:p1], sc["z0s"][q0:q1])
for it in range(3):
    print('dist rank', rank, 'it', it, ctx.outer_iteration(), file=sys.stderr, flush=True)
ctx.close(); dist.destroy_process_group()
