"""Synthetic scenes of the BASELINE.json shapes (SURVEY §8d generator), built with torch on the
GPU and returned as HOST arrays in the reference's masked layouts (full mask => masked vector ==
column-major image).  Used by bench.py; data only, no solver code."""
from __future__ import annotations

import numpy as np


def synth_scene_torch(h, w, sf, n, seed, device="cuda", pin=True, j0=0, j1=None):
    """Full-mask scene; with (j0, j1) only image columns [j0, j1) are generated (one strip of a
    strip-partitioned scene: same analytic surface / albedo / lights, per-strip noise stream).
    Returned arrays cover the strip's pixels; `mask` is always the GLOBAL mask."""
    import torch
    import torch.nn.functional as F
    j1 = w if j1 is None else j1
    wl = j1 - j0
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) * 1000003 + j0)
    rng = np.random.default_rng(seed)
    fx = fy = 1.2 * w
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    ii = torch.arange(h, device=device, dtype=torch.float64)[None, :]     # cols   (image rows i, contiguous)
    v = (ii - cy) / h

    def surface(ja, jb):
        jj_ = torch.arange(ja, jb, device=device, dtype=torch.float64)[:, None]     # lines (image columns j)
        u_ = (jj_ - cx) / w
        return jj_, u_, 700 + 60 * torch.exp(-9 * (u_ * u_ + v * v)) + 8 * torch.sin(9 * u_) * torch.cos(7 * v)

    jj, u, zt = surface(j0, j1)                                                       # [wl][h]
    # normals of z* with the reference's forward-else-backward differences (SRPS.cu:23-71)
    zx = torch.empty_like(zt); zy = torch.empty_like(zt)
    zx[:-1] = zt[1:] - zt[:-1]
    zx[-1] = (surface(j1, j1 + 1)[2][0] - zt[-1]) if j1 < w else (zt[-1] - zt[-2])
    zy[:, :-1] = zt[:, 1:] - zt[:, :-1]; zy[:, -1] = zt[:, -1] - zt[:, -2]
    xx = (jj - cx).expand(wl, h); yy = (ii - cy).expand(wl, h)
    n0 = fx * zx; n1 = fy * zy; n2 = -zt - xx * zx - yy * zy
    nrm = torch.sqrt(n0 * n0 + n1 * n1 + n2 * n2).clamp_min(1e-10)
    Nt = torch.stack([n0 / nrm, n1 / nrm, n2 / nrm]).float()            # [3][w][h]
    L = rng.standard_normal((n, 3))
    L[:, 2] = -(np.abs(L[:, 2]) + 1.5)
    L /= np.linalg.norm(L, axis=1, keepdims=True)
    s_true = np.concatenate([L, np.full((n, 1), 0.2)], axis=1).astype(np.float32)
    rho_t = torch.stack([0.55 + 0.3 * torch.sin(20 * u + k) * torch.cos(17 * v) for k in range(3)]).float()   # [3][w][h]
    npix = h * wl
    I = torch.empty((n, 3, npix), dtype=torch.float32, pin_memory=pin)
    for i in range(n):
        sv = torch.tensor(s_true[i], device=device)
        shade = sv[0] * Nt[0] + sv[1] * Nt[1] + sv[2] * Nt[2] + sv[3]
        img = rho_t * shade[None] + 0.01 * torch.randn((3, wl, h), device=device, generator=gen)
        I[i].copy_(img.clamp_(0, 1).reshape(3, npix))
    z0 = F.avg_pool2d(zt.float()[None, None], sf)[0, 0] + torch.randn((wl // sf, h // sf), device=device, generator=gen)
    sm = F.avg_pool2d(F.pad(z0[None, None], (1, 1, 1, 1), mode="replicate"), 3, stride=1)
    z_init = F.interpolate(sm, size=(wl, h), mode="bicubic", align_corners=False)[0, 0]
    z = torch.empty(npix, dtype=torch.float32, pin_memory=pin); z.copy_(z_init.reshape(-1))
    z0s = torch.empty((wl // sf) * (h // sf), dtype=torch.float32, pin_memory=pin); z0s.copy_(z0.reshape(-1))
    K = np.array([fx, 0, 0, 0, fy, 0, cx, cy, 1], dtype=np.float64)
    torch.cuda.synchronize()
    return dict(h=h, w=w, sf=sf, n=n, c=3, K=K, mask=np.ones((h, w), np.uint8), I=I.numpy(), z=z.numpy(), z0s=z0s.numpy(),
                j0=j0, j1=j1, _keep=(I, z, z0s))
