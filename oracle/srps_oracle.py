"""CPU restatement (numpy/scipy) of the SRmeetsPS outer loop -- TEST INFRASTRUCTURE ONLY.

This module is the *oracle*: a plain restatement of the reference algorithm
(nihalsid/SRmeetsPS-CUDA, `SRmeetsPS-GPU/SRPS.cu` + `devicecalls.cu`) used to check the
CUDA product path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.  The product path
(`srmeetsps-cuda_b200/`) never imports anything from `oracle/`.

Parity pin: the reference ships no tests / golden vectors for this path (SURVEY.md §4,
§8c).  The pin is the reference's own *unmodified* device code (`devicecalls.cu`),
compiled by `oracle/ref/Makefile` into `oracle/_ref/ref_replay` and run on a B200; its
per-iteration outputs are committed as `tests/golden/ref_*.npz` by
`oracle/ref/make_goldens.py`, and `tests/test_oracle_vs_ref_goldens.py` checks this
restatement against them.  Until those goldens exist for a scene, parity for it is
"unpinned".

Layouts are the reference's (column-major images, masked vectors):
    pixel (row i, col j) <-> linear i + j*h            Utilities.cpp:330,343
    masked vector        = mask pixels in ascending linear order   SRPS.cu:157-162
    I   [n][c][npix]     SRPS.cu:223-232        s   [n][c][4]    SRPS.cu:209-217
    rho [c][npix]        devicecalls.cu:133-149 N   [4][npix]    devicecalls.cu:194-223
    z   [npix], z0s [npixs], xx/yy [npix]       SRPS.cu:237-260

Every function takes `dt` (np.float32 = reference-faithful arithmetic, np.float64 =
ground truth) and cites the reference lines it follows.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

CG_TOL = 1e-9          # devicecalls.cu:230
CG_MAX_ITER = 100      # devicecalls.cu:231  (loop runs while k <= max_iter -> 101 passes)
OUTER_TOL = 5e-3       # SRPS.cu:85
OUTER_MAX_ITER = 10    # SRPS.cu:86
LAMBDA = 1.0           # devicecalls.cu:644


# --------------------------------------------------------------------------------------
# Operators (one-shot, CPU in the reference)
# --------------------------------------------------------------------------------------
def mask_indices(mask: np.ndarray):
    """imask + index_in_masked_matrix.  SRPS.cu:153-162.  mask is (h, w), non-zero = in."""
    h, w = mask.shape
    flat = (np.asarray(mask) != 0).ravel(order="F")          # column-major linearisation
    imask = np.flatnonzero(flat).astype(np.int64)
    index_in_masked = np.zeros(h * w, dtype=np.int64)
    index_in_masked[imask] = np.arange(imask.size)
    return imask, index_in_masked


def make_gradient(mask: np.ndarray):
    """Dx, Dy on the mask: forward difference if the next pixel is in the mask, else
    backward if the previous one is, else an empty row.  SRPS.cu:23-71.
    'y' runs along rows i (the fast axis), 'x' along columns j."""
    h, w = mask.shape
    m = np.asarray(mask) != 0
    imask, idx = mask_indices(mask)
    npix = imask.size
    idx2 = idx.reshape((h, w), order="F")

    def one_dir(axis):
        nxt = np.zeros_like(m)
        prv = np.zeros_like(m)
        if axis == 0:
            nxt[:-1, :] = m[1:, :]
            prv[1:, :] = m[:-1, :]
        else:
            nxt[:, :-1] = m[:, 1:]
            prv[:, 1:] = m[:, :-1]
        fwd = m & nxt                                   # SRPS.cu:31 / :39
        bwd = m & ~nxt & prv                            # SRPS.cu:35 / :43 (else-if)
        rows, cols, vals = [], [], []
        sh = (1, 0) if axis == 0 else (0, 1)
        ii, jj = np.nonzero(fwd)
        me = idx2[ii, jj]
        nb = idx2[ii + sh[0], jj + sh[1]]
        rows += [me, me]; cols += [nb, me]
        vals += [np.ones(me.size), -np.ones(me.size)]   # SRPS.cu:51,57  (k1=1 @nb, k2=-1 @self)
        ii, jj = np.nonzero(bwd)
        me = idx2[ii, jj]
        nb = idx2[ii - sh[0], jj - sh[1]]
        rows += [me, me]; cols += [nb, me]
        vals += [-np.ones(me.size), np.ones(me.size)]   # SRPS.cu:54,60  (k1=-1 @nb, k2=1 @self)
        D = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(npix, npix))
        return D, fwd, bwd

    Dy, yf, yb = one_dir(0)
    Dx, xf, xb = one_dir(1)
    return Dx, Dy, dict(xf=xf, xb=xb, yf=yf, yb=yb)


def downsampling_matrix(h: int, w: int, sf: int):
    """D: (h*w/sf^2) x (h*w), block average with weight 1/sf^2.  Utilities.cpp:201-220."""
    hs, ws = h // sf, w // sf
    r = np.arange(hs * ws)
    q, rr = r // hs, r % hs
    base = q * h * sf + rr * sf
    jj, kk = np.meshgrid(np.arange(sf), np.arange(sf), indexing="ij")
    cols = (base[:, None] + (jj.ravel() * h + kk.ravel())[None, :]).ravel()
    rows = np.repeat(r, sf * sf)
    vals = np.full(rows.size, 1.0 / (sf * sf))
    return sp.csr_matrix((vals, (rows, cols)), shape=(hs * ws, h * w))


def build_operators(mask: np.ndarray, sf: int):
    """LR mask, imask/imasks, masked resample matrix KT, Dx, Dy.
    SRPS.cu:105-115 (LR mask = D*mask, <1 -> 0), :153-193 (KT), :197-203 (Dx, Dy)."""
    h, w = mask.shape
    assert h % sf == 0 and w % sf == 0, "image size must be a multiple of sf"
    imask, idx = mask_indices(mask)
    D = downsampling_matrix(h, w, sf)
    mflat = (np.asarray(mask) != 0).ravel(order="F").astype(np.float64)
    masks = D @ mflat
    masks[masks < 1.0] = 0.0                                  # SRPS.cu:111
    imasks = np.flatnonzero(masks != 0)                       # SRPS.cu:163-166
    KT = D[imasks][:, imask].tocsr()                          # SRPS.cu:176-189
    Dx, Dy, types = make_gradient(mask)
    return dict(h=h, w=w, sf=sf, imask=imask, imasks=imasks, npix=imask.size,
                npixs=imasks.size, KT=KT, Dx=Dx, Dy=Dy, types=types,
                masks=(masks != 0).reshape((h // sf, w // sf), order="F"))


def meshgrid_masked(ops, cx: float, cy: float, dt=np.float32):
    """xx = j - K[6], yy = i - K[7] on the mask.  devicecalls.cu:151-158 with the w/h
    launch swap of :164-166 corrected (identical wherever the reference's
    grid covers the mask)."""
    h = ops["h"]
    lin = ops["imask"]
    i, j = lin % h, lin // h
    return (j.astype(dt) - dt(cx)).astype(dt), (i.astype(dt) - dt(cy)).astype(dt)


# --------------------------------------------------------------------------------------
# Normals                                                     devicecalls.cu:171-223
# --------------------------------------------------------------------------------------
def normals(z, xx, yy, ops, fx, fy, dt=np.float32):
    zx = (ops["Dx"].astype(dt) @ z.astype(dt)).astype(dt)     # SRPS.cu:264,310
    zy = (ops["Dy"].astype(dt) @ z.astype(dt)).astype(dt)     # SRPS.cu:265,311
    N = np.empty((4, z.size), dtype=dt)
    N[0] = dt(fx) * zx                                        # devicecalls.cu:204
    N[1] = dt(fy) * zy                                        # devicecalls.cu:211
    N[2] = -z - xx * zx - yy * zy                             # devicecalls.cu:174
    N[3] = 1                                                  # devicecalls.cu:175
    dz = np.maximum(dt(1e-10), np.sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2])).astype(dt)
    N[:3] /= dz                                               # devicecalls.cu:186-192
    return N, dz, zx, zy


# --------------------------------------------------------------------------------------
# CG exactly as the reference runs it                         devicecalls.cu:229-279
# --------------------------------------------------------------------------------------
def cg_reference(matvec, x, b, dt=np.float32, max_iter=CG_MAX_ITER, tol=CG_TOL):
    """x, b are updated in place semantics-wise (returned); b holds the residual on entry.
    Returns (x, iterations)."""
    x = x.astype(dt).copy()
    b = b.astype(dt).copy()
    k = 0
    r0 = dt(0)
    r1 = dt(np.dot(b, b))
    p = None
    tol2 = dt(tol) * dt(tol)
    while r1 > tol2 and k <= max_iter:
        k += 1
        if k == 1:
            p = b.copy()
        else:
            beta = dt(r1 / r0)
            p = (beta * p + b).astype(dt)
        om = matvec(p).astype(dt)
        dot = dt(np.dot(p, om))
        alpha = dt(r1 / dot)
        x = (x + alpha * p).astype(dt)
        b = (b - alpha * om).astype(dt)
        r0 = r1
        r1 = dt(np.dot(b, b))
    return x, k


def cg_fused_reference(matvec, x, b, dt=np.float32, max_iter=CG_MAX_ITER, tol=CG_TOL, defer_rel=1e-6, stats=None, lazy_z=False,
                       lazy_min_beta=1e-3):
    """The recurrence of the CUDA path's one-kernel-per-pass CG (csrc/srps_cg.cuh: cg_fused_kernel), restated to
    show that it is the reference's CG (cg_reference above, devicecalls.cu:229-279) in another order of operations:
    pass k first applies the step of pass k-1 (r -= alpha y, x += alpha p), then forms p and y = A p and four dots;
    alpha_k = r.r / p.y with r.r MEASURED; beta_{k+1} = |r_{k+1}|^2 / r.r with |r_{k+1}|^2 = |r - alpha y|^2 expanded
    from the dots, r.y = p.y - beta (y_prev . p) (A symmetric).  A pass that measures r.r <= tol^2 is void.
    Guard: when the expansion cancels (|r_{k+1}|^2 < defer_rel * r.r) the next pass slot only applies the pending step
    and measures r.r; beta then is the reference's r1/r0 of measured norms.  stats (dict) receives the number of
    deferred passes.

    lazy_z=True restates the schedule of cg_persistent_fused_kernel: a pass with only one step pending may leave x
    untouched; the next slot (pass, deferred slot or tail) applies both steps, the older direction recovered from its
    own operands as (p_in - r_in) / beta_link.  stats["zskip"] counts the passes that skipped x."""
    f64 = np.float64
    x = x.astype(dt).copy()
    r = b.astype(dt).copy()
    tol2 = dt(tol) * dt(tol)
    k = 0
    alpha = dt(0); beta = dt(0)
    alpha_old = dt(0); beta_link = dt(1)
    p = np.zeros_like(r); y = np.zeros_like(r)
    active = dt(np.dot(r.astype(f64), r.astype(f64))) > tol2
    deferred = False
    n_deferred = n_zskip = 0
    r0 = 0.0
    tail = None                          # (c1, c2, p, r) set by a void pass that still owes a step
    slots = max_iter + 1 + 2            # FUSED_SPARE_PASSES
    while active and slots > 0:
        slots -= 1
        # what this slot does to x
        zc1, zc2, zskip = alpha, dt(0), False
        if alpha_old != 0:
            zc2 = dt(-alpha_old / beta_link); zc1 = dt(alpha - zc2)
        elif alpha == 0:
            zskip = True
        elif lazy_z and not deferred and beta >= dt(lazy_min_beta):
            zskip = True
        r_in, p_in = r, p
        if not zskip:
            x = (x + (zc1 * p_in + zc2 * r_in).astype(dt)).astype(dt)
        r = (r_in - alpha * y).astype(dt)
        if deferred:                     # fused_update_only: p, y unchanged, r.r measured
            S0 = float(np.dot(r.astype(f64), r.astype(f64)))
            alpha = dt(0); alpha_old = dt(0)
            if not (dt(S0) > tol2):
                break
            beta = dt(dt(S0) / dt(r0))
            deferred = False
            n_deferred += 1
            continue
        y_prev = y
        p = (r + beta * p_in).astype(dt)
        y = matvec(p).astype(dt)
        S0 = float(np.dot(r.astype(f64), r.astype(f64)))
        S1 = float(np.dot(p.astype(f64), y.astype(f64)))
        S3 = float(np.dot(y.astype(f64), y.astype(f64)))
        C = float(np.dot(y_prev.astype(f64), p.astype(f64)))
        if not (dt(S0) > tol2):          # the reference left its loop before this pass
            if zskip and alpha != 0:     # ... and this pass had left its step to the next one: the tail applies it
                tail = (alpha, dt(0), p_in, r_in)
            alpha = dt(0); alpha_old = dt(0)
            break
        al = dt(dt(S0) / dt(S1))
        S2 = S1 - float(beta) * C
        rr = S0 - 2.0 * float(al) * S2 + float(al) ** 2 * S3
        deferred = not (rr > defer_rel * S0)
        r0 = S0
        if zskip and alpha != 0:
            alpha_old, beta_link = alpha, beta      # p = r + beta p_in links the two directions
            n_zskip += 1
        else:
            alpha_old = dt(0)
        beta = dt(0) if deferred else dt(dt(rr) / dt(S0))
        alpha = al
        k += 1
        active = k <= max_iter
    if tail is None:                     # the step(s) still pending after the last pass (cg_tail_kernel / the kernel's tail loop)
        zc2 = dt(-alpha_old / beta_link) if alpha_old != 0 else dt(0)
        tail = (dt(alpha - zc2), zc2, p, r)
    x = (x + (tail[0] * tail[2] + tail[1] * tail[3]).astype(dt)).astype(dt)
    if stats is not None:
        stats["deferred"] = n_deferred
        stats["zskip"] = n_zskip
    return x, k


# --------------------------------------------------------------------------------------
# Lighting                                                    devicecalls.cu:376-444
# --------------------------------------------------------------------------------------
def lighting_update(s, rho, N, I, dt=np.float32, direct=False):
    """For every (image i, channel c): 4x4 normal equations of A = rho_c (.) N, solved by
    the warm-started reference CG (or directly if direct=True)."""
    n, c, _ = s.shape
    s = s.astype(dt).copy()
    for ch in range(c):
        A = (rho[ch][None, :] * N).astype(dt)                 # [4][npix]   devicecalls.cu:381
        ATA = (A @ A.T).astype(dt)                            # devicecalls.cu:422
        for i in range(n):
            ATb = (A @ I[i, ch].astype(dt)).astype(dt)        # devicecalls.cu:423
            if direct:
                s[i, ch] = np.linalg.solve(ATA.astype(np.float64), ATb.astype(np.float64)).astype(dt)
                continue
            res = (ATb - ATA @ s[i, ch]).astype(dt)           # devicecalls.cu:424
            s[i, ch], _ = cg_reference(lambda v: ATA @ v, s[i, ch], res, dt)   # :437
    return s


# --------------------------------------------------------------------------------------
# Albedo                                                      devicecalls.cu:497-548
# --------------------------------------------------------------------------------------
def albedo_terms(s, N, I, ch, dt=np.float32):
    """Shading A[i][p] = N_p . s_{i,ch}  (devicecalls.cu:507), diagonal AtA and At b (:395-406)."""
    A = (s[:, ch, :].astype(dt) @ N.astype(dt)).astype(dt)    # [n][npix]
    d = np.einsum("ip,ip->p", A, A).astype(dt)
    b = np.einsum("ip,ip->p", A, I[:, ch, :].astype(dt)).astype(dt)
    return d, b


def albedo_update(s, rho, N, I, dt=np.float32, closed_form=False):
    """Per channel: CG (101 passes) on the *diagonal* system diag(d) rho = b, warm-started
    (devicecalls.cu:531,540); closed_form=True returns b/d (what that CG converges to)."""
    rho = rho.astype(dt).copy()
    iters = []
    for ch in range(rho.shape[0]):
        d, b = albedo_terms(s, N, I, ch, dt)
        if closed_form:
            ok = d > 0
            rho[ch][ok] = (b[ok] / d[ok]).astype(dt)
            iters.append(0)
            continue
        res = (b - d * rho[ch]).astype(dt)                    # devicecalls.cu:404-405
        rho[ch], k = cg_reference(lambda v: d * v, rho[ch], res, dt)
        iters.append(k)
    return rho, iters


# --------------------------------------------------------------------------------------
# Depth                                                       devicecalls.cu:550-786
# --------------------------------------------------------------------------------------
def depth_rows(s, rho, dz, xx, yy, fx, fy, dt=np.float32):
    """Dense coefficient planes a1,a2,a3 [c][n][npix]  (devicecalls.cu:583-620)."""
    r = (rho / dz[None, :]).astype(dt)                                    # [c][npix]
    s0 = s[:, :, 0].T.astype(dt); s1 = s[:, :, 1].T.astype(dt); s2 = s[:, :, 2].T.astype(dt)  # [c][n]
    a1 = r[:, None, :] * (dt(fx) * s0[:, :, None] - xx[None, None, :] * s2[:, :, None])
    a2 = r[:, None, :] * (dt(fy) * s1[:, :, None] - yy[None, None, :] * s2[:, :, None])
    a3 = r[:, None, :] * s2[:, :, None]
    return a1.astype(dt), a2.astype(dt), a3.astype(dt)


def depth_B(s, rho, I, dt=np.float32):
    """B[c][j][p] = I[j][c][p] - rho[c][p] * N3 * s[j][c][3], N3 == 1  (devicecalls.cu:550-581)."""
    return (I.transpose(1, 0, 2).astype(dt) - rho[:, None, :].astype(dt) * s[:, :, 3].T[:, :, None].astype(dt)).astype(dt)


def depth_update_assembled(s, rho, I, xx, yy, dz, ops, z0s, z, fx, fy, dt=np.float32):
    """Literal restatement: stack A (c*n*npix rows, 3 nnz/row), form KtK + lambda AtA, rhs,
    residual, 101-pass CG, energy with lagged A,B and the new z (devicecalls.cu:636-786).
    Memory ~ c*n*npix rows: use for small scenes."""
    npix = z.size
    n, c, _ = s.shape
    a1, a2, a3 = depth_rows(s, rho, dz, xx, yy, fx, fy, dt)
    B = depth_B(s, rho, I, dt).reshape(-1)
    Dx = ops["Dx"].astype(dt); Dy = ops["Dy"].astype(dt); KT = ops["KT"].astype(dt)
    blocks = []
    for ch in range(c):
        for j in range(n):
            blocks.append(sp.diags(a1[ch, j]) @ Dx + sp.diags(a2[ch, j]) @ Dy - sp.diags(a3[ch, j]))
    A = sp.vstack(blocks).tocsr().astype(dt)                               # :668-723
    A_ = (KT.T @ KT + dt(LAMBDA) * (A.T @ A)).tocsr().astype(dt)          # :734-736
    rhs = (KT.T @ z0s.astype(dt) + dt(LAMBDA) * (A.T @ B)).astype(dt)      # :743-745
    res = (rhs - A_ @ z.astype(dt)).astype(dt)                             # :758
    znew, k = cg_reference(lambda v: A_ @ v, z, res, dt)                   # :759
    t1 = np.sum(((KT @ znew) - z0s.astype(dt)) ** 2, dtype=dt)             # :762-766
    t2 = np.sum(((A @ znew) - B) ** 2, dtype=dt)                           # :763-767
    return znew, float(t1 + dt(LAMBDA) * t2), k, dict(A_=A_, rhs=rhs)


def depth_coeffs(s, rho, I, xx, yy, dz, fx, fy, dt=np.float64):
    """Per-pixel 3x3 M_p (6 unique), g_p (3) and the constant e0_p = sum_{c,j} B^2 such that
    AtA = G^T diag(M) G,  AtB = G^T g,  ||Az-B||^2 = sum_p (Gz)^T M (Gz) - 2 g.(Gz) + e0
    with G_p z = [(Dx z)_p, (Dy z)_p, z_p]  (same algebra as devicecalls.cu:583-620,
    :550-581 without materialising the c*n*npix-row matrix)."""
    a1, a2, a3 = depth_rows(s, rho, dz, xx, yy, fx, fy, dt)
    B = depth_B(s, rho, I, dt)
    t = np.stack([a1, a2, -a3], axis=0)                                    # [3][c][n][npix]
    M = np.einsum("acjp,bcjp->abp", t, t)
    g = np.einsum("acjp,cjp->ap", t, B)
    e0 = np.einsum("cjp,cjp->p", B, B)
    return M.astype(dt), g.astype(dt), e0.astype(dt)


def depth_update_matfree(s, rho, I, xx, yy, dz, ops, z0s, z, fx, fy, dt=np.float64):
    """Same update through the stencil form y = Kt K p + G^T M G p (no c*n*npix-row matrix)."""
    M, g, e0 = depth_coeffs(s, rho, I, xx, yy, dz, fx, fy, dt)
    Dx = ops["Dx"].astype(dt); Dy = ops["Dy"].astype(dt); KT = ops["KT"].astype(dt)

    def G(v):
        return np.stack([Dx @ v, Dy @ v, v])

    def GT(q):
        return Dx.T @ q[0] + Dy.T @ q[1] + q[2]

    def Aop(v):
        gv = G(v)
        q = np.einsum("abp,bp->ap", M, gv)
        return (KT.T @ (KT @ v) + dt(LAMBDA) * GT(q)).astype(dt)

    rhs = (KT.T @ z0s.astype(dt) + dt(LAMBDA) * GT(g)).astype(dt)
    res = (rhs - Aop(z.astype(dt))).astype(dt)
    znew, k = cg_reference(Aop, z, res, dt)
    gz = G(znew)
    e_photo = np.sum(np.einsum("ap,abp,bp->p", gz, M, gz) - 2 * np.einsum("ap,ap->p", g, gz) + e0, dtype=np.float64)
    e_depth = np.sum(((KT @ znew) - z0s.astype(dt)) ** 2, dtype=np.float64)
    return znew, float(e_depth + LAMBDA * e_photo), k, dict(Aop=Aop, rhs=rhs, M=M, g=g, e0=e0)


# --------------------------------------------------------------------------------------
# Loop state + outer loop                                     SRPS.cu:206-335
# --------------------------------------------------------------------------------------
def init_state(I_masked, z_masked, z0s, ops, K, dt=np.float32):
    """s=(0,0,-1,0), rho=0.5, xx,yy, first normals  (SRPS.cu:209-270).  K is the reference's
    column-major 3x3 (K[0]=fx, K[4]=fy, K[6]=cx, K[7]=cy)."""
    n, c, npix = I_masked.shape
    fx, fy, cx, cy = float(K[0]), float(K[4]), float(K[6]), float(K[7])
    s = np.zeros((n, c, 4), dtype=dt); s[:, :, 2] = -1
    rho = np.full((c, npix), 0.5, dtype=dt)
    xx, yy = meshgrid_masked(ops, cx, cy, dt)
    z = z_masked.astype(dt).copy()
    N, dz, _, _ = normals(z, xx, yy, ops, fx, fy, dt)
    return dict(I=I_masked.astype(dt), s=s, rho=rho, z=z, z0s=z0s.astype(dt), xx=xx, yy=yy,
                N=N, dz=dz, fx=fx, fy=fy, cx=cx, cy=cy)


def outer_iteration(st, ops, dt=np.float32, assembled=False, albedo_closed_form=False,
                    lighting_direct=False):
    """One pass of the do-while body (SRPS.cu:276-317).  Returns energy and CG counts."""
    st["s"] = lighting_update(st["s"], st["rho"], st["N"], st["I"], dt, direct=lighting_direct)
    st["rho"], ak = albedo_update(st["s"], st["rho"], st["N"], st["I"], dt, closed_form=albedo_closed_form)
    fn = depth_update_assembled if assembled else depth_update_matfree
    st["z"], energy, k, _ = fn(st["s"], st["rho"], st["I"], st["xx"], st["yy"], st["dz"], ops,
                               st["z0s"], st["z"], st["fx"], st["fy"], dt)
    st["N"], st["dz"], _, _ = normals(st["z"], st["xx"], st["yy"], ops, st["fx"], st["fy"], dt)
    return energy, k, ak


def run(st, ops, dt=np.float32, max_outer=OUTER_MAX_ITER, tol=OUTER_TOL, fixed_iters=None, **kw):
    """Termination rule of SRPS.cu:273-301 (first comparison is against NaN -> never stops)."""
    last = float("nan")
    it = 1
    hist = []
    while True:
        e, k, ak = outer_iteration(st, ops, dt, **kw)
        rel = abs(last - e) / abs(e)
        stop = (e > last) or (rel < tol) or (it > max_outer)
        if fixed_iters is not None:
            stop = it >= fixed_iters
        last = e
        hist.append(dict(iteration=it, energy=e, rel_err=rel, cg_iters=k, albedo_cg_iters=ak))
        it += 1
        if stop:
            break
    return hist


# --------------------------------------------------------------------------------------
# Synthetic scenes (SURVEY §8d generator) and post-init snapshots
# --------------------------------------------------------------------------------------
def synth_scene(h, w, sf, n, seed, mask_kind="full", noise=True, dt=np.float32):
    """Deterministic synthetic scene in the reference's loop-state layouts (post-init
    snapshot: I masked, z = bicubic-free upsampled LR depth, z0s, mask, K)."""
    rng = np.random.default_rng(seed)
    fx = fy = 1.2 * w
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    jj, ii = np.meshgrid(np.arange(w), np.arange(h))          # (h, w)
    u = (jj - cx) / w
    v = (ii - cy) / h
    zt = 700 + 60 * np.exp(-9 * (u * u + v * v)) + 8 * np.sin(9 * u) * np.cos(7 * v)
    if mask_kind == "full":
        mask = np.ones((h, w), dtype=np.float32)
    elif mask_kind == "ellipse":
        mask = (((u / 0.45) ** 2 + (v / 0.45) ** 2) < 1).astype(np.float32)
    elif mask_kind in ("random", "random95"):
        # irregular mask with holes / thin features: exercises fwd/bwd/none stencils.  "random" (80 %
        # density) leaves very few fully-masked LR blocks -> a barely constrained, ill-conditioned depth
        # solve; "random95" keeps the problem well conditioned.
        m = rng.random((h, w)) < (0.8 if mask_kind == "random" else 0.95)
        m &= ((u / 0.48) ** 2 + (v / 0.48) ** 2) < 1
        mask = m.astype(np.float32)
    else:
        raise ValueError(mask_kind)
    ops = build_operators(mask, sf)
    K = np.array([fx, 0, 0, 0, fy, 0, cx, cy, 1], dtype=np.float64)   # column-major 3x3
    xx, yy = meshgrid_masked(ops, cx, cy, np.float64)
    zt_m = zt.ravel(order="F")[ops["imask"]]
    Nt, _, _, _ = normals(zt_m, xx, yy, ops, fx, fy, np.float64)
    L = rng.standard_normal((n, 3))
    L[:, 2] = -(np.abs(L[:, 2]) + 1.5)
    L /= np.linalg.norm(L, axis=1, keepdims=True)
    s_true = np.concatenate([L, np.full((n, 1), 0.2)], axis=1)         # [n][4]
    c = 3
    npix = ops["npix"]
    um, vm = u.ravel(order="F")[ops["imask"]], v.ravel(order="F")[ops["imask"]]
    rho_t = np.stack([0.55 + 0.3 * np.sin(20 * um + k) * np.cos(17 * vm) for k in range(c)])
    I = np.empty((n, c, npix), dtype=dt)
    shade = s_true @ Nt                                                # [n][npix]
    for i in range(n):
        img = rho_t * shade[i][None, :]
        if noise:
            img = img + 0.01 * rng.standard_normal(img.shape)
        I[i] = np.clip(img, 0, 1).astype(dt)
    D = downsampling_matrix(h, w, sf)
    z0 = D @ zt.ravel(order="F")
    if noise:
        z0 = z0 + 1.0 * rng.standard_normal(z0.shape)
    z0s = z0[ops["imasks"]].astype(dt)
    # initial HR depth: nearest-neighbour upsample of the LR depth smoothed by a 3x3 box
    z0_img = z0.reshape((h // sf, w // sf), order="F")
    pad = np.pad(z0_img, 1, mode="edge")
    sm = sum(pad[a:a + z0_img.shape[0], b:b + z0_img.shape[1]] for a in range(3) for b in range(3)) / 9.0
    z_init = np.kron(sm, np.ones((sf, sf))).ravel(order="F")[ops["imask"]].astype(dt)
    return dict(h=h, w=w, sf=sf, n=n, c=c, mask=mask, K=K, I=I, z=z_init, z0s=z0s, ops=ops,
                truth=dict(z=zt_m, rho=rho_t, s=s_true))


# --------------------------------------------------------------------------------------
# One-shot depth pre-processing (SRPS.cu:117-149, devicecalls.cu:95-125) via python cv2
# --------------------------------------------------------------------------------------
def preprocess_depth(z0, h, w, sf):
    """z0: (z0_n, h/sf * w/sf) column-major frames.  Mean (/nc always, zero frames flagged),
    TELEA inpaint r=16, bilateral(-1,2,2) on depth/max, bicubic upsample -- applied, as the
    reference does, to the TRANSPOSED image (cv::Mat(z0_w rows, z0_h cols), SRPS.cu:130-149)."""
    import cv2
    hs, ws = h // sf, w // sf
    z0 = np.asarray(z0, dtype=np.float32).reshape(-1, hs * ws)
    nc = z0.shape[0]
    flags = (z0 == 0).any(axis=0).astype(np.uint8)                       # devicecalls.cu:101-105
    zs = (np.where(z0 != 0, z0, 0).sum(axis=0, dtype=np.float32) / np.float32(nc)).astype(np.float32)
    zs_mat = zs.reshape(ws, hs).copy()                                   # (z0_w rows, z0_h cols)
    fl_mat = flags.reshape(ws, hs).copy()
    zs_mat = cv2.inpaint(zs_mat, fl_mat, 16, cv2.INPAINT_TELEA)          # SRPS.cu:133
    mx = float(zs_mat.max())                                             # SRPS.cu:137
    out = cv2.bilateralFilter((zs_mat / mx).astype(np.float32), -1, 2, 2) * np.float32(mx)   # :138-140
    z_full = cv2.resize(out, (h, w), interpolation=cv2.INTER_CUBIC)      # cv::Size(I_h, I_w)  :149
    return out.reshape(-1).astype(np.float32), z_full.reshape(-1).astype(np.float32)
