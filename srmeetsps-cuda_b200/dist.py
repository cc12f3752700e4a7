"""Strip partition of one scene across the GPUs of a box: host-side helpers.

One process per GPU (torchrun); `torch.distributed` is used only as the rendezvous that carries the
CUDA-IPC blobs between the ranks -- all data-path exchanges (ghost lines, CG scalars, lighting sums)
happen inside the library's own kernels over NVLink peer memory (csrc/srps_comm.cuh)."""
from __future__ import annotations

import numpy as np


def strip_bounds(w: int, world: int, align: int = 4):
    """Cut image columns [0, w) into `world` contiguous strips whose boundaries are multiples of
    `align` (sf and the 4-line group of the operator kernel), as equal as possible."""
    if w % align:
        raise ValueError("image width must be a multiple of the alignment")
    units = w // align
    if units < world:
        raise ValueError("more ranks than aligned column groups")
    base, extra = divmod(units, world)
    bounds, j = [], 0
    for r in range(world):
        n = (base + (1 if r < extra else 0)) * align
        bounds.append((j, j + n))
        j += n
    return bounds


def local_ranges(mask, sf: int, j0: int, j1: int):
    """Half-open ranges of the global masked vector ([p0,p1)) and of the global LR masked vector
    ([q0,q1)) owned by image columns [j0, j1): strips cut the column-major order into runs."""
    m = np.asarray(mask) != 0
    h, w = m.shape
    cols = m.sum(axis=0)
    p0, p1 = int(cols[:j0].sum()), int(cols[:j1].sum())
    lr = m.reshape(h // sf, sf, w // sf, sf).all(axis=(1, 3)).sum(axis=0)
    q0, q1 = int(lr[: j0 // sf].sum()), int(lr[: j1 // sf].sum())
    return p0, p1, q0, q1


def exchange_blobs(blob: bytes, group=None):
    """all-gather of the per-rank IPC blobs in rank order (any torch.distributed backend)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, blob, group=group)
    return out


def make_strip_context(mask, n_images, sf, K, rank, world, device, **kw):
    """Create this rank's strip context and wire it to its peers."""
    from .context import Context
    h, w = np.asarray(mask).shape
    j0, j1 = strip_bounds(w, world, align=4)[rank]
    ctx = Context(mask, n_images, sf, K, device=device, strip=(j0, j1), rank=rank, world=world, **kw)
    ctx.dist_connect(exchange_blobs(ctx.dist_export()))
    return ctx
