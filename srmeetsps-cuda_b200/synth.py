"""Synthetic scenes of the BASELINE.json shapes (SURVEY §8d generator), built with torch on the
GPU and returned as HOST arrays in the reference's masked layouts (full mask => masked vector ==
column-major image).  Used by bench.py; data only, no solver code.

The noise is a counter-based hash of (seed, stream, global pixel index): a strip of a scene holds exactly
the values the whole scene holds on those columns, so a strip-partitioned run at any GPU count solves the
SAME scene as the single-GPU run (energies and results are comparable across N)."""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1


def _s64(v: int) -> int:
    """two's-complement int64 view of a 64-bit constant (torch has no uint64 arithmetic)"""
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x, k: int):
    """logical shift right of an int64 tensor"""
    return (x >> k) & ((1 << (64 - k)) - 1)


def _mix(x):
    """splitmix64 finaliser on int64 tensors (wrap-around multiplication)"""
    x = (x ^ _lsr(x, 30)) * _s64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _s64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def hash_normal(idx, seed: int, stream: int):
    """Standard normal deviates as a pure function of (seed, stream, idx): idx is an int64 tensor of global
    element indices.  Box-Muller on two 53-bit uniforms from two splitmix64 rounds."""
    import torch
    key = _s64(seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03 + 0x2545F4914F6CDD1D)
    a = _mix(idx * _s64(0x9E3779B97F4A7C15) + key)
    b = _mix(a + _s64(0x9E3779B97F4A7C15))
    u1 = (_lsr(a, 11).to(torch.float64) + 0.5) * (1.0 / (1 << 53))
    u2 = (_lsr(b, 11).to(torch.float64) + 0.5) * (1.0 / (1 << 53))
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * np.pi * u2)


def synth_scene_torch(h, w, sf, n, seed, device="cuda", pin=True, j0=0, j1=None):
    """Full-mask scene; with (j0, j1) only image columns [j0, j1) are generated (one strip of a
    strip-partitioned scene: the same analytic surface / albedo / lights AND the same noise values as the
    whole scene has on those columns).  Returned arrays cover the strip's pixels; `mask` is always the GLOBAL mask."""
    import torch
    import torch.nn.functional as F
    j1 = w if j1 is None else j1
    wl = j1 - j0
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
    # global linear pixel index (column-major: i + j*h) of the strip's pixels, and the per-plane offset
    gidx = (torch.arange(j0, j1, device=device, dtype=torch.int64)[:, None] * h
            + torch.arange(h, device=device, dtype=torch.int64)[None, :])                 # [wl][h]
    I = torch.empty((n, 3, npix), dtype=torch.float32, pin_memory=pin)
    for i in range(n):
        sv = torch.tensor(s_true[i], device=device)
        shade = sv[0] * Nt[0] + sv[1] * Nt[1] + sv[2] * Nt[2] + sv[3]
        noise = torch.stack([hash_normal(gidx, seed, 1 + i * 3 + c) for c in range(3)]).float()
        img = rho_t * shade[None] + 0.01 * noise
        I[i].copy_(img.clamp_(0, 1).reshape(3, npix))
    hl = h // sf
    lidx = (torch.arange(j0 // sf, j1 // sf, device=device, dtype=torch.int64)[:, None] * hl
            + torch.arange(hl, device=device, dtype=torch.int64)[None, :])
    z0 = F.avg_pool2d(zt.float()[None, None], sf)[0, 0] + hash_normal(lidx, seed, 0).float()
    # z initialisation = smoothed, bicubically upsampled z0 (the role of SRPS.cu:133-149).  Smoothing and resampling
    # look one LR pixel across a strip boundary: a strip computes them on its LR columns plus a one-column apron
    # of the neighbours' z0 (same hash => same values), so the result does not depend on the partition either.
    a0, a1 = max(j0 // sf - 3, 0), min(j1 // sf + 3, w // sf)
    if (a0, a1) != (j0 // sf, j1 // sf):
        _, _, zt_a = surface(a0 * sf, a1 * sf)
        lidx_a = (torch.arange(a0, a1, device=device, dtype=torch.int64)[:, None] * hl
                  + torch.arange(hl, device=device, dtype=torch.int64)[None, :])
        z0_a = F.avg_pool2d(zt_a.float()[None, None], sf)[0, 0] + hash_normal(lidx_a, seed, 0).float()
    else:
        z0_a = z0
    sm = F.avg_pool2d(F.pad(z0_a[None, None], (1, 1, 1, 1), mode="replicate"), 3, stride=1)
    z_full = F.interpolate(sm, size=((a1 - a0) * sf, h), mode="bicubic", align_corners=False)[0, 0]
    z_init = z_full[(j0 // sf - a0) * sf:(j0 // sf - a0) * sf + wl]
    z = torch.empty(npix, dtype=torch.float32, pin_memory=pin); z.copy_(z_init.reshape(-1))
    z0s = torch.empty((wl // sf) * hl, dtype=torch.float32, pin_memory=pin); z0s.copy_(z0.reshape(-1))
    K = np.array([fx, 0, 0, 0, fy, 0, cx, cy, 1], dtype=np.float64)
    torch.cuda.synchronize() if str(device).startswith("cuda") else None
    return dict(h=h, w=w, sf=sf, n=n, c=3, K=K, mask=np.ones((h, w), np.uint8), I=I.numpy(), z=z.numpy(), z0s=z0s.numpy(),
                j0=j0, j1=j1, _keep=(I, z, z0s))
