"""n_scan_normal_reg::Register written a second time, in numpy, straight from the reference source (n_scan_normal.cpp:82-187
outer loop, :215-326 AddScanPairCost, registration.cpp:67-76 weights) with the Ceres loop of test_lm_independent.py, and
compared with the oracle on the same cell sets: outer and inner iteration counts, residual count, final cost, pose and the
covariance of GetCovariance (:392-433).
CPU only.  What is shared with the oracle is the input (cell sets from the oracle's surface-point stage) and nothing else:
the nearest neighbour is a brute-force float32 search, the P2D factor comes from numpy.linalg, the solve is the QR-based loop.
"""
import numpy as np
import pytest

import helpers
from test_lm_independent import _ceres_lm, _loss


def _se2(v):
    c, s = np.cos(v[2]), np.sin(v[2])
    return np.array([[c, -s], [s, c]]), np.asarray(v[:2], dtype=np.float64)


def _similarity(x, y):
    return 2 * np.minimum(x, y) / (x + y)                                           # registration.cpp:73-75


def _associate(tar, src, Ttar, Tsrc, radius, cost, weight_opt, reg, cov_scale):
    """AddScanPairCost for one (keyframe, scan) pair: arrays (p, q, A, w) of the residual blocks, in source-cell order."""
    Rt, tt = Ttar; Rs, ts = Tsrc
    R = Rt.T @ Rs; t = Rt.T @ (ts - tt)                                             # Ttar^-1 * Tsrc             :224
    pm = src["mean"] @ R.T + t                                                      # src_trans_mean             :240
    # MapPointNormal::GetClosestIdx (pointnormal.cpp:238-254): float query, float cell means, 1-NN, float d2 < d*d
    q32 = pm.astype(np.float32); m32 = tar["mean"].astype(np.float32)
    dx = q32[:, None, 0] - m32[None, :, 0]; dy = q32[:, None, 1] - m32[None, :, 1]
    d2 = dx * dx + dy * dy                                                          # float32 throughout (FLANN L2_Simple)
    m = d2.argmin(1)                                                                # first minimum = smallest index on ties
    near = d2[np.arange(m.size), m].astype(np.float64) < radius * radius
    nt = src["normal"] @ R.T                                                        # src_normal_trans           :244
    sim = np.maximum((nt * tar["normal"][m]).sum(1), 0.0)                           #                            :246
    keep = near & (sim > np.cos(np.pi / 6.0))                                       #                            :247
    j = np.nonzero(keep)[0]; m = m[j]; sim = sim[j]
    n1, n2 = src["nsamples"][j].astype(np.float64), tar["nsamples"][m].astype(np.float64)
    p1, p2 = src["planarity"][j], tar["planarity"][m]
    w = {0: np.ones(j.size), 1: _similarity(n1, n2), 2: sim, 3: _similarity(p1, p2),
         4: _similarity(n1, n2) + sim + _similarity(p1, p2)}[weight_opt]              # registration.cpp:67-76
    q = tar["mean"][m] @ Rt.T + tt                                                  # Ttar * tar_mean            :279 / :299
    if cost == "P2L":
        A = np.zeros((j.size, 2, 2)); A[:, 0, :] = tar["normal"][m] @ Rt.T          # Ttar.linear() * tar_normal :285
    elif cost == "P2D":
        C = tar["cov"][m].reshape(-1, 2, 2)
        S = (reg * np.eye(2) + Rt @ C @ Rt.T) * cov_scale                           #                            :292-296
        A = np.linalg.cholesky(np.linalg.inv(S))                                    # tar_cov.inverse().llt().matrixL()  :297
    else:
        A = np.tile(-np.eye(2), (j.size, 1, 1))                                     # P2P: tar - (R p + t)       n_scan_normal.h:336-351
    return src["mean"][j], q, A, w


def register_py(sets, poses, cost, weight_opt, reg=1.0, cov_scale=1.0, radius=2.0, max_outer=8, min_outer=3, max_inner=20,
                loss_limit=0.1, loss="Huber", want_cov=True):
    K = len(sets) - 1
    rows = 1 if cost == "P2L" else 2
    x = np.array(poses[K], dtype=np.float64)
    prev_par, prev_score = x.copy(), np.finfo(np.float64).max
    itr, inner_total, success, nres, final_cost = 1, 0, True, 0, 0.0
    while itr <= max_outer and success:                                             # :102
        cur = 2 * radius if itr == 1 else radius                                    # :222
        parts = [_associate(sets[i], sets[K], _se2(poses[i]), _se2(x), cur, cost, weight_opt, reg, cov_scale) for i in range(K)]
        p, q, A, w = (np.concatenate([pt[k] for pt in parts]) for k in range(4))
        nres = rows * p.shape[0]
        if nres <= 1:                                                               # :370
            success = False
            break

        def fun(y, want_jac):
            c, s = np.cos(y[2]), np.sin(y[2])
            Rp = np.stack([c * p[:, 0] - s * p[:, 1], s * p[:, 0] + c * p[:, 1]], 1)
            res = np.einsum("nij,nj->ni", A, Rp + y[:2] - q)[:, :rows]
            rho, rho1 = _loss(loss, (res * res).sum(1), loss_limit)
            c_ = 0.5 * (w * rho).sum()                                              # ScaledLoss(loss, w)        :277
            if not want_jac:
                return c_, None, None
            Je = np.zeros((p.shape[0], 2, 3)); Je[:, 0, 0] = 1; Je[:, 1, 1] = 1
            Je[:, 0, 2] = -Rp[:, 1]; Je[:, 1, 2] = Rp[:, 0]
            Jr = np.einsum("nij,njk->nik", A, Je)[:, :rows, :]
            sc = np.sqrt(w * rho1)
            return c_, (res * sc[:, None]).ravel(), (Jr * sc[:, None, None]).reshape(-1, 3)

        x, n_it, final_cost, last_rel = _ceres_lm(fun, x, max_inner)               # :117
        inner_total += n_it
        rel_improvement = (prev_score - final_cost) / prev_score                    # :124
        if itr > min_outer:                                                         # :134-149
            if prev_score < final_cost:
                x = prev_par.copy()
                break
            if rel_improvement < 0.00001:
                break
            if last_rel < 0.00001 or n_it == 0:
                break
        prev_score, prev_par = final_cost, x.copy()
        itr += 1
    cov = None
    if success and nres - 3 != 0 and want_cov:                                      # GetCovariance              :392-433
        _, _, Jc = fun(x, True)                                                     # the LAST problem, loss applied, at the final x
        cmat = 30 * (final_cost / (nres - 3)) * np.linalg.inv(Jc.T @ Jc)             #                            :418
        cov = np.eye(6)
        cov[:2, :2] = cmat[:2, :2]; cov[5, 5] = cmat[2, 2]; cov[0, 5] = cmat[0, 2]; cov[5, 0] = cmat[2, 0]   # (1,5) / (5,1) stay 0  :426-430
    return success, x, itr, inner_total, nres, final_cost, cov


@pytest.mark.parametrize("cost,wopt,reg", [("P2L", 0, 1.0), ("P2D", 4, 0.1), ("P2D", 0, 1.0), ("P2P", 2, 1.0), ("P2L", 3, 1.0)])
@pytest.mark.parametrize("seed,K,offset", [(3, 1, (0.4, -0.3, 0.02)), (6, 2, (-0.8, 0.5, -0.03)), (9, 3, (1.5, 1.0, 0.05))])
def test_oracle_register_matches_an_independent_numpy_register(orc, cost, wopt, reg, seed, K, offset):
    im, tp = helpers.scan_images(seed, K)
    sets = [helpers.oracle_cells(orc, im[i], radius=3.0)[1] for i in range(K + 1)]
    P = tp[:K + 1].copy(); P[K] = tp[K] + np.asarray(offset)
    cfg = orc.reg_cfg(cost=cost, loss="Huber", loss_limit=0.1, weight_opt=wopt, regularization=reg, cov_scale=1.0)
    ok, op, ocov, st, _ = orc.register(sets, P, cfg)
    okp, x, itr, inner, nres, fc, cov = register_py(sets, P, cost, wopt, reg=reg)
    assert ok and okp
    assert (itr, inner, nres) == (st.outer_iterations, st.inner_iterations, st.num_residuals), \
        ((itr, inner, nres), (st.outer_iterations, st.inner_iterations, st.num_residuals))
    np.testing.assert_allclose(fc, st.final_cost, rtol=1e-9)
    assert np.hypot(*(x[:2] - op[K, :2])) < 1e-9 and abs(x[2] - op[K, 2]) < 1e-10
    np.testing.assert_allclose(ocov, cov, rtol=1e-6, atol=1e-15)
    assert ocov[1, 5] == 0.0 and ocov[5, 1] == 0.0 and ocov[0, 5] != 0.0


@pytest.mark.parametrize("loss", ["Cauchy", "SoftLOne", "Tukey", "Combined", "None"])
@pytest.mark.parametrize("cost,wopt,reg,seed,K,offset", [("P2D", 4, 0.1, 6, 2, (-0.8, 0.5, -0.03)), ("P2L", 1, 1.0, 9, 3, (0.5, 0.4, 0.02))])
def test_oracle_register_matches_the_numpy_register_for_every_loss(orc, loss, cost, wopt, reg, seed, K, offset):
    """Registration::GetLoss (registration.cpp:78-97): Cauchy, SoftLOne, Tukey, Huber(1) o Cauchy(1), and no loss."""
    im, tp = helpers.scan_images(seed, K)
    sets = [helpers.oracle_cells(orc, im[i], radius=3.0)[1] for i in range(K + 1)]
    P = tp[:K + 1].copy(); P[K] = tp[K] + np.asarray(offset)
    cfg = orc.reg_cfg(cost=cost, loss=loss, loss_limit=0.1, weight_opt=wopt, regularization=reg, cov_scale=1.0)
    ok, op, ocov, st, _ = orc.register(sets, P, cfg)
    okp, x, itr, inner, nres, fc, cov = register_py(sets, P, cost, wopt, reg=reg, loss=loss)
    assert ok == okp
    assert (itr, inner, nres) == (st.outer_iterations, st.inner_iterations, st.num_residuals), \
        ((itr, inner, nres), (st.outer_iterations, st.inner_iterations, st.num_residuals))
    np.testing.assert_allclose(fc, st.final_cost, rtol=1e-9)
    assert np.hypot(*(x[:2] - op[K, :2])) < 1e-9 and abs(x[2] - op[K, 2]) < 1e-10
    if ok:
        np.testing.assert_allclose(ocov, cov, rtol=1e-6, atol=1e-15)


def get_cost_py(sets, poses, cost, weight_opt, reg=1.0, cov_scale=1.0, radius=2.0, loss_limit=0.1, loss="Huber"):
    """n_scan_normal_reg::GetCost (n_scan_normal.cpp:187-213) after a Register (itr_ > 1: the association radius is not doubled):
    one association at the given poses and 1/2 sum w rho(s)."""
    K = len(sets) - 1
    rows = 1 if cost == "P2L" else 2
    x = np.asarray(poses[K], dtype=np.float64)
    parts = [_associate(sets[i], sets[K], _se2(poses[i]), _se2(x), radius, cost, weight_opt, reg, cov_scale) for i in range(K)]
    p, q, A, w = (np.concatenate([pt[k] for pt in parts]) for k in range(4))
    c, s = np.cos(x[2]), np.sin(x[2])
    Rp = np.stack([c * p[:, 0] - s * p[:, 1], s * p[:, 0] + c * p[:, 1]], 1)
    res = np.einsum("nij,nj->ni", A, Rp + x[:2] - q)[:, :rows]
    rho, _ = _loss(loss, (res * res).sum(1), loss_limit)
    return rows * p.shape[0] > 1, 0.5 * (w * rho).sum(), rows * p.shape[0]


@pytest.mark.parametrize("cost,wopt,reg", [("P2D", 4, 0.1), ("P2L", 0, 1.0)])
def test_oracle_get_cost_matches_the_numpy_get_cost_on_the_sampling_grid(orc, cost, wopt, reg):
    """The 27 GetCost evaluations of approximateCovarianceBySampling (odometrykeyframefuser.cpp:261-380: +-0.2 m, +-0.00218 rad,
    three per axis) around a registered pose: every sample re-associates (n_scan_normal.cpp:199-202)."""
    K = 2
    im, tp = helpers.scan_images(12, K)
    sets = [helpers.oracle_cells(orc, im[i], radius=3.0)[1] for i in range(K + 1)]
    cfg = orc.reg_cfg(cost=cost, loss="Huber", loss_limit=0.1, weight_opt=wopt, regularization=reg, cov_scale=1.0)
    P = tp[:K + 1].copy()
    n = 0
    for t in np.linspace(-0.00218125, 0.00218125, 3):
        for dx in np.linspace(-0.2, 0.2, 3):
            for dy in np.linspace(-0.2, 0.2, 3):
                Q = P.copy(); Q[K] = P[K] + [dx, dy, t]
                ok, c, nres = orc.get_cost(sets, Q, cfg)
                okp, cp, nresp = get_cost_py(sets, Q, cost, wopt, reg=reg)
                assert ok and okp and nres == nresp > 100
                np.testing.assert_allclose(c, cp, rtol=1e-11)
                n += 1
    assert n == 27


def test_oracle_register_fuzz_against_the_numpy_register(orc):
    """Sixty seeded random configurations -- cost, loss, loss limit, weight option, regularisation, 1-3 keyframes, start offsets
    from centimetres to several metres (rejected steps, failed associations, exhausted loops): same counts, costs and poses.
    A Tukey start so far off that every residual is an outlier leaves rank-deficient normal equations; there Ceres'
    `Covariance` reports failure by the reciprocal condition number (1e-14) while the oracle and K5 only fail on a
    non-positive Cholesky pivot: such cases are recognised here and only their poses are compared."""
    rng = np.random.Generator(np.random.PCG64(123))
    cache = {}
    n_checked = n_degenerate = 0
    for _trial in range(60):
        cost = str(rng.choice(["P2L", "P2D", "P2P"]))
        loss = str(rng.choice(["Huber", "Cauchy", "SoftLOne", "Tukey", "Combined", "None"]))
        wopt = int(rng.integers(0, 5)); K = int(rng.integers(1, 4)); seed = int(rng.integers(40, 44))
        reg = float(rng.choice([0.1, 1.0])); ll = float(rng.choice([0.05, 0.1, 0.5])); scale = float(rng.choice([0.1, 0.5, 1.5, 4.0]))
        off = np.array([rng.normal(0, 1) * scale, rng.normal(0, 1) * scale, rng.normal(0, 0.03) * scale])
        if (seed, K) not in cache:
            im, tp = helpers.scan_images(seed, K)
            cache[(seed, K)] = ([helpers.oracle_cells(orc, im[i], radius=3.0)[1] for i in range(K + 1)], tp)
        sets, tp = cache[(seed, K)]
        P = tp[:K + 1].copy(); P[K] = tp[K] + off
        cfg = orc.reg_cfg(cost=cost, loss=loss, loss_limit=ll, weight_opt=wopt, regularization=reg, cov_scale=1.0)
        ok, op, ocov, st, _ = orc.register(sets, P, cfg)
        try:
            okp, x, itr, inner, nres, fc, cov = register_py(sets, P, cost, wopt, reg=reg, loss=loss, loss_limit=ll)
            degenerate = cov is None or not np.all(np.isfinite(cov)) or np.abs(cov).max() > 1e6
        except np.linalg.LinAlgError:                                     # exactly singular J^T J in the covariance
            degenerate = True
            okp, x, itr, inner, nres, fc, cov = register_py(sets, P, cost, wopt, reg=reg, loss=loss, loss_limit=ll, want_cov=False)
        assert (itr, inner, nres) == (st.outer_iterations, st.inner_iterations, st.num_residuals), (cost, loss, wopt, K, seed, off)
        if okp:
            assert np.hypot(*(x[:2] - op[K, :2])) < 1e-8 and abs(x[2] - op[K, 2]) < 1e-9, (cost, loss, wopt, K, seed, off)
            np.testing.assert_allclose(fc, st.final_cost, rtol=1e-8, atol=1e-300)
        if degenerate:
            n_degenerate += 1
            continue
        assert ok == okp
        if ok:
            np.testing.assert_allclose(ocov, cov, rtol=1e-5, atol=1e-14)
        n_checked += 1
    assert n_checked >= 50 and n_degenerate <= 8, (n_checked, n_degenerate)
