"""CPU tests of the oracle itself (no GPU): the stencil form equals the reference's assembled
sparse system, the C transcription equals the numpy restatement, the reference quirks hold."""
import numpy as np
import pytest

from oracle import srps_oracle as o
from oracle.port import Port

from conftest import rel_rmse


@pytest.mark.parametrize("mask_kind,sf", [("random", 2), ("ellipse", 4), ("full", 2), ("random", 1)])
def test_stencil_form_equals_assembled_system(mask_kind, sf):
    """K^T K + G^T M G == the matrix the reference assembles with SpGEMM/SpGEAM
    (devicecalls.cu:668-736), and G^T g == A^T B (devicecalls.cu:744)."""
    sc = o.synth_scene(24, 32, sf, 4, seed=5, mask_kind=mask_kind)
    ops = sc["ops"]
    st = o.init_state(sc["I"], sc["z"], sc["z0s"], ops, sc["K"], np.float64)
    rng = np.random.default_rng(0)
    st["s"] = rng.standard_normal(st["s"].shape)
    st["rho"] = 0.3 + rng.random(st["rho"].shape)
    _, _, _, asm = o.depth_update_assembled(st["s"], st["rho"], st["I"], st["xx"], st["yy"], st["dz"], ops,
                                            st["z0s"], st["z"], st["fx"], st["fy"], np.float64)
    _, _, _, mf = o.depth_update_matfree(st["s"], st["rho"], st["I"], st["xx"], st["yy"], st["dz"], ops,
                                         st["z0s"], st["z"], st["fx"], st["fy"], np.float64)
    A_ = asm["A_"]
    for _ in range(3):
        v = rng.standard_normal(ops["npix"])
        ref = A_ @ v
        got = mf["Aop"](v)
        assert np.abs(ref - got).max() <= 1e-10 * np.abs(ref).max()
    assert np.abs(asm["rhs"] - mf["rhs"]).max() <= 1e-10 * np.abs(asm["rhs"]).max()
    # <= 9 non-zeros per row for sf <= 2
    if sf <= 2:
        assert np.diff(A_.indptr).max() <= 9 + sf * sf


def test_gradient_semantics():
    """forward-else-backward-else-none, SRPS.cu:23-71."""
    mask = np.zeros((5, 6), np.float32)
    mask[1:4, 1:5] = 1
    mask[2, 2] = 0
    mask[0, 0] = 1                     # isolated pixel: empty rows
    Dx, Dy, t = o.make_gradient(mask)
    imask, idx = o.mask_indices(mask)
    z = np.arange(imask.size, dtype=np.float64) ** 2
    img = np.zeros(30); img[imask] = z; img = img.reshape((5, 6), order="F")
    zx, zy = Dx @ z, Dy @ z
    for p, lin in enumerate(imask):
        i, j = lin % 5, lin // 5
        def m(a, b): return 0 <= a < 5 and 0 <= b < 6 and mask[a, b] != 0
        ey = img[i + 1, j] - img[i, j] if m(i + 1, j) else (img[i, j] - img[i - 1, j] if m(i - 1, j) else 0.0)
        ex = img[i, j + 1] - img[i, j] if m(i, j + 1) else (img[i, j] - img[i, j - 1] if m(i, j - 1) else 0.0)
        assert zy[p] == ey and zx[p] == ex
    assert Dx[idx[0]].nnz == 0 and Dy[idx[0]].nnz == 0


def test_lr_mask_and_KT():
    """LR pixel kept only if all sf^2 HR pixels are masked (SRPS.cu:110-111); KT rows average them."""
    sc = o.synth_scene(16, 24, 4, 2, seed=2, mask_kind="ellipse")
    ops = sc["ops"]
    m = sc["mask"].reshape(4, 4, 6, 4).transpose(0, 2, 1, 3).reshape(4, 6, 16).all(axis=2)
    assert ops["npixs"] == int(m.sum())
    KT = ops["KT"]
    assert np.all(np.diff(KT.indptr) == 16)
    assert np.allclose(KT.data, 1 / 16)


def test_c_port_matches_numpy_oracle():
    sc = o.synth_scene(32, 48, 2, 6, seed=1, mask_kind="random")
    ops = sc["ops"]
    st64 = o.init_state(sc["I"], sc["z"], sc["z0s"], ops, sc["K"], np.float64)
    st = o.init_state(sc["I"], sc["z"], sc["z0s"], ops, sc["K"], np.float32)
    pt = Port(ops, sc["n"], sc["c"], st["fx"], st["fy"], st["xx"], st["yy"])
    stp = {k: np.ascontiguousarray(st[k]).copy() for k in ("s", "rho", "z", "N", "dz", "I", "z0s")}
    N, dz = pt.normals(stp["z"])
    assert np.abs(N - st["N"]).max() < 1e-5
    for it in range(3):
        e64, k64, ak64 = o.outer_iteration(st64, ops, np.float64, assembled=True)
        e, k, ak = pt.outer_iteration(stp)
        assert k == 101 and k64 == 101              # r.r never reaches 1e-18: exactly max_iter+1 passes (devicecalls.cu:252)
        assert abs(e - e64) <= 1e-4 * abs(e64)
        assert rel_rmse(stp["z"], st64["z"]) < 1e-4
        assert np.abs(stp["rho"] - st64["rho"]).max() < 1e-3
        assert np.abs(stp["s"] - st64["s"]).max() < 2e-3


def test_albedo_cg_equals_closed_form():
    """The reference's diagonal CG (devicecalls.cu:531,540) converges to b/d."""
    sc = o.synth_scene(32, 32, 2, 5, seed=4, mask_kind="ellipse")
    ops = sc["ops"]
    st = o.init_state(sc["I"], sc["z"], sc["z0s"], ops, sc["K"], np.float32)
    st["s"] = o.lighting_update(st["s"], st["rho"], st["N"], st["I"], np.float32)
    r_cg, iters = o.albedo_update(st["s"], st["rho"], st["N"], st["I"], np.float32)
    r_cf, _ = o.albedo_update(st["s"], st["rho"], st["N"], st["I"], np.float32, closed_form=True)
    assert max(iters) < 101
    assert np.abs(r_cg - r_cf).max() < 1e-4


def test_synth_generator_is_deterministic():
    a = o.synth_scene(16, 16, 2, 3, seed=9)
    b = o.synth_scene(16, 16, 2, 3, seed=9)
    assert np.array_equal(a["I"], b["I"]) and np.array_equal(a["z0s"], b["z0s"])


def test_mitten_snapshot_matches_survey_anchors(mitten_scene):
    """npix / npixs / gradient-type counts of SURVEY §6 and the fp64 energies of BASELINE.md §5."""
    ops = mitten_scene["ops"]
    assert ops["npix"] == 148600 and ops["npixs"] == 36915
    t = ops["types"]
    assert (int(t["xf"].sum()), int(t["xb"].sum())) == (148005, 595)
    assert (int(t["yf"].sum()), int(t["yb"].sum())) == (148200, 397)
    st = o.init_state(mitten_scene["I"], mitten_scene["z"], mitten_scene["z0s"], ops, mitten_scene["K"], np.float32)
    pt = Port(ops, mitten_scene["n"], mitten_scene["c"], st["fx"], st["fy"], st["xx"], st["yy"])
    stp = {k: np.ascontiguousarray(st[k]).copy() for k in ("s", "rho", "z", "N", "dz", "I", "z0s")}
    anchors = [38058.6, 34266.5, 33475.5]
    for it in range(3):
        e, k, _ = pt.outer_iteration(stp)
        assert k == 101
        assert abs(e - anchors[it]) < 2e-4 * anchors[it]


@pytest.mark.parametrize("cfg", [dict(h=96, w=128, sf=4, n=8, seed=1, mask_kind="ellipse"),
                                 dict(h=128, w=96, sf=2, n=6, seed=3, mask_kind="random95"),
                                 dict(h=40, w=24, sf=1, n=6, seed=5, mask_kind="random95")],
                         ids=lambda c: f"{c['h']}x{c['w']}sf{c['sf']}{c['mask_kind']}")
def test_fused_cg_recurrence_is_the_reference_cg(cfg):
    """The one-kernel-per-pass CG of the CUDA path (oracle.cg_fused_reference restates its recurrence) against the
    reference's CG on the fp32 depth system: same pass count (101, or the early stop of the sf = 1 scene) and a
    solution within fp32 round-off of it -- closer than the fp64-vector run of the reference recurrence is."""
    sc = o.synth_scene(**cfg)
    ops = o.build_operators(sc["mask"], sc["sf"])
    st = o.init_state(sc["I"], sc["z"], sc["z0s"], ops, sc["K"])
    s = o.lighting_update(st["s"], st["rho"], st["N"], st["I"])
    rho, _ = o.albedo_update(s, st["rho"], st["N"], st["I"], closed_form=True)
    _, _, _, aux = o.depth_update_matfree(s, rho, st["I"], st["xx"], st["yy"], st["dz"], ops, st["z0s"], st["z"],
                                          st["fx"], st["fy"], dt=np.float32)
    Aop, rhs = aux["Aop"], aux["rhs"]
    res = (rhs - Aop(st["z"].astype(np.float32))).astype(np.float32)
    z_ref, k_ref = o.cg_reference(Aop, st["z"], res, np.float32)
    z_fus, k_fus = o.cg_fused_reference(Aop, st["z"], res, np.float32)
    z_f64, _ = o.cg_reference(Aop, st["z"], res, np.float64)
    assert k_fus == k_ref
    assert rel_rmse(z_fus, z_ref) <= 5e-7
    assert rel_rmse(z_fus, z_ref) <= 2.0 * rel_rmse(z_f64.astype(np.float32), z_ref) + 1e-9


def test_fused_cg_guard_measures_when_the_expansion_cancels():
    """Early convergence (a well-conditioned SPD system: every step removes most of the residual, the last ones almost
    all of it): the expanded |r - alpha y|^2 of the fused recurrence cancels, the guard defers to a measured r.r and
    the pass count / solution stay those of the reference's CG.  Without the guard beta is noise (or negative)."""
    rng = np.random.default_rng(0)
    n = 400
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    # clustered spectrum: CG needs about one pass per cluster; the residual collapses by > 1e3 per pass at the end
    lam = np.concatenate([np.full(n - 3, 1.0), [2.0, 3.0, 5.0]])
    A = ((Q * lam) @ Q.T).astype(np.float32)
    A = 0.5 * (A + A.T)
    mv = lambda v: (A @ v.astype(np.float32)).astype(np.float32)
    x0 = np.zeros(n, np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    z_ref, k_ref = o.cg_reference(mv, x0, b, np.float32)
    stats = {}
    z_fus, k_fus = o.cg_fused_reference(mv, x0, b, np.float32, stats=stats)
    assert stats["deferred"] >= 1, "the scenario must exercise the guard"
    assert abs(k_fus - k_ref) <= 1
    assert np.all(np.isfinite(z_fus))
    sol = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
    err_ref = np.linalg.norm(z_ref - sol) / np.linalg.norm(sol)
    err_fus = np.linalg.norm(z_fus - sol) / np.linalg.norm(sol)
    assert err_fus <= 2.0 * err_ref + 1e-6


def _spd(rng, n, lam):
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = ((Q * lam) @ Q.T).astype(np.float32)
    return 0.5 * (A + A.T)


@pytest.mark.parametrize("max_iter", [0, 1, 2, 3, 4, 7, 100])
@pytest.mark.parametrize("spectrum", ["spread", "clustered", "identity_like"])
def test_lazy_depth_update_is_the_eager_one(spectrum, max_iter):
    """The persistent kernel's lazy z schedule (oracle.cg_fused_reference(lazy_z=True) restates it: every other pass
    leaves x untouched, the next slot applies both steps with the older direction recovered as (p_in - r_in) / beta)
    gives the solution of the eager recurrence to fp32 round-off -- for every way a solve can end: iteration limit after an
    odd / even number of passes (tail with one or two pending steps), early convergence (void pass that had skipped x),
    deferred slots (clustered spectrum), nothing to do."""
    rng = np.random.default_rng(7)
    n = 300
    lam = {"spread": np.geomspace(1.0, 300.0, n),
           "clustered": np.concatenate([np.full(n - 3, 1.0), [2.0, 3.0, 5.0]]),
           "identity_like": 1.0 + 1e-3 * rng.random(n)}[spectrum]
    A = _spd(rng, n, lam)
    mv = lambda v: (A @ v.astype(np.float32)).astype(np.float32)
    x0 = rng.standard_normal(n).astype(np.float32) * np.float32(10.0)        # steps are small against x, as for depths
    b = rng.standard_normal(n).astype(np.float32)
    se, sl = {}, {}
    xe, ke = o.cg_fused_reference(mv, x0, b, np.float32, max_iter=max_iter, stats=se)
    xl, kl = o.cg_fused_reference(mv, x0, b, np.float32, max_iter=max_iter, stats=sl, lazy_z=True)
    assert kl == ke and sl["deferred"] == se["deferred"]
    assert se["zskip"] == 0
    if ke >= 2 and spectrum == "spread":          # (the other two converge so fast that beta falls below lazy_min_beta)
        assert sl["zskip"] >= 1, "the schedule must skip x somewhere"
    step = np.linalg.norm(xe - x0)
    assert np.linalg.norm(xl - xe) <= 2e-6 * np.linalg.norm(xe) + 1e-5 * step


def test_lazy_depth_update_keeps_a_tiny_beta_eager():
    """The recovery divides by beta: below lazy_min_beta a pass applies its step at once."""
    rng = np.random.default_rng(3)
    n = 200
    A = _spd(rng, n, np.concatenate([np.full(n - 1, 1.0), [1.0 + 1e-2]]))
    mv = lambda v: (A @ v.astype(np.float32)).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    x0 = np.zeros(n, np.float32)
    s0, s1 = {}, {}
    xe, ke = o.cg_fused_reference(mv, x0, b, np.float32, stats=s0, lazy_z=True, lazy_min_beta=10.0)    # never lazy
    xl, kl = o.cg_fused_reference(mv, x0, b, np.float32, stats=s1, lazy_z=True)
    assert s0["zskip"] == 0 and ke == kl
    assert np.linalg.norm(xl - xe) <= 2e-6 * max(np.linalg.norm(xe), 1e-30)
