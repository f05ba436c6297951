"""A second, independent writing of the Ceres trust-region loop (numpy, dense QR) against the oracle's.

The reference hands every outer iteration's residual blocks to `ceres::Solve` (n_scan_normal.cpp:443-452) with default
options apart from `max_num_iterations` (n_scan_normal.cpp:9,18): TRUST_REGION / LEVENBERG_MARQUARDT, Jacobi scaling,
initial radius 1e4, function / gradient / parameter tolerances 1e-6 / 1e-10 / 1e-8, min_relative_decrease 1e-3.  Ceres
is not in the image, so the oracle (oracle/cfear_oracle.cc) and K5 restate that loop; this file restates it once more from
the structure of Ceres 1.13 / 1.14's `TrustRegionMinimizer::Minimize` and `LevenbergMarquardtStrategy` -- residuals and
Jacobians corrected for the loss like `Corrector` does for rho'' <= 0, the damped step from a QR factorisation of the
augmented Jacobian like DENSE_QR, not from normal equations -- and requires the same number of iterations, the same final
cost and the same minimiser.  CPU only; it guards the oracle's bookkeeping (step acceptance, radius schedule, which
iterations are counted, the order of the convergence tests), not Ceres' arithmetic to the last bit.
"""
import numpy as np
import pytest

import helpers


def _ceres_lm(fun, x0, max_num_iterations=20):
    """fun(x, want_jac) -> (cost, r, J): cost = 1/2 sum rho_i, r / J the loss-corrected residuals / Jacobian.
    Returns (x, iterations.size() - 1, final_cost, last relative_decrease)."""
    function_tolerance, gradient_tolerance, parameter_tolerance = 1e-6, 1e-10, 1e-8
    min_relative_decrease, min_diag, max_diag = 1e-3, 1e-6, 1e32
    max_radius, min_radius = 1e16, 1e-32
    radius, decrease_factor, reuse_diagonal = 1e4, 2.0, False
    x = np.array(x0, dtype=np.float64)
    x_cost, r, J = fun(x, True)
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))              # jacobi_scaling, fixed at iteration 0
    Js = J * scale
    g = J.T @ r
    iteration, n_done, last_rel, invalid = 0, 0, 0.0, 0
    diagonal = None
    if np.abs(g).max() <= gradient_tolerance:
        return x, 0, x_cost, 0.0
    while True:
        if iteration >= max_num_iterations or radius <= min_radius:
            break
        iteration += 1
        # LevenbergMarquardtStrategy::ComputeStep
        if not reuse_diagonal:
            diagonal = np.clip((Js * Js).sum(0), min_diag, max_diag)
        lm_diagonal = np.sqrt(diagonal / radius)
        A = np.vstack([Js, np.diag(lm_diagonal)])
        b = np.concatenate([-r, np.zeros(3)])
        Q, R = np.linalg.qr(A)
        step = np.linalg.solve(R, Q.T @ b)
        reuse_diagonal = True
        model = Js @ step
        model_cost_change = -model @ (r + model / 2.0)
        if not np.all(np.isfinite(step)) or not model_cost_change > 0.0:
            invalid += 1
            if invalid >= 5:
                raise RuntimeError("too many invalid steps")
            radius /= decrease_factor; decrease_factor *= 2.0
            n_done += 1; last_rel = 0.0
            continue
        invalid = 0
        delta = step * scale
        cand = x + delta
        cand_cost, _, _ = fun(cand, False)
        # ParameterToleranceReached / FunctionToleranceReached: the terminating iteration is not recorded
        if np.linalg.norm(delta) <= parameter_tolerance * (np.linalg.norm(x) + parameter_tolerance):
            break
        cost_change = x_cost - cand_cost
        if abs(cost_change) <= function_tolerance * x_cost:
            break
        rel = cost_change / model_cost_change
        n_done += 1; last_rel = rel
        if rel > min_relative_decrease:                                   # HandleSuccessfulStep
            x = cand
            x_cost, r, J = fun(x, True)
            Js = J * scale
            g = J.T @ r
            radius = min(max_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * rel - 1.0) ** 3))
            decrease_factor, reuse_diagonal = 2.0, False
            if np.abs(g).max() <= gradient_tolerance:
                break
        else:                                                             # HandleUnsuccessfulStep
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = True
    return x, n_done, x_cost, last_rel


def _huber(s, a):
    out = s > a * a
    rho = np.where(out, 2 * a * np.sqrt(np.maximum(s, 1e-300)) - a * a, s)
    rho1 = np.where(out, a / np.sqrt(np.maximum(s, 1e-300)), 1.0)
    return rho, rho1


def _loss(name, s, a):
    """(rho, rho') of the Ceres loss functions Registration::GetLoss hands out (registration.cpp:78-97; loss_function.cc).  All
    have rho'' <= 0, so Ceres' Corrector only scales rows by sqrt(rho')."""
    tiny = np.finfo(np.float64).tiny
    if name == "Huber":
        return _huber(s, a)
    if name == "Cauchy":
        b = a * a
        return b * np.log1p(s / b), np.maximum(tiny, 1.0 / (1.0 + s / b))
    if name == "SoftLOne":
        b = a * a
        t = np.sqrt(1.0 + s / b)
        return 2 * b * (t - 1.0), np.maximum(tiny, 1.0 / t)
    if name == "Tukey":
        b = a * a
        v = 1.0 - s / b
        return np.where(s <= b, b / 3.0 * (1.0 - v ** 3), b / 3.0), np.where(s <= b, v * v, 0.0)
    if name == "Combined":                                                # ComposedLoss(HuberLoss(1), CauchyLoss(1))
        g, g1 = _loss("Cauchy", s, 1.0)
        f, f1 = _loss("Huber", g, 1.0)
        return f, f1 * g1
    return s, np.ones_like(s)                                             # None: ScaledLoss(nullptr, w)


def _problem(orc, cost, seed, offset):
    """One outer iteration's residual blocks exactly as the reference builds them (n_scan_normal.cpp:215-326), from the
    oracle's association table: (p, q, A) per block with r = A (R(psi) p + t - q)."""
    im, tp = helpers.scan_images(seed, 1)
    sets = [helpers.oracle_cells(orc, im[i], radius=3.0)[1] for i in range(2)]
    P = tp[:2].copy(); P[1] = tp[1] + np.asarray(offset)
    cfg = orc.reg_cfg(cost=cost, loss="Huber", loss_limit=0.1, weight_opt=0, regularization=0.1 if cost == "P2D" else 1.0,
                      max_outer=1, max_inner=20)
    ok, op, _, st, assoc = orc.register(sets, P, cfg, want_assoc=True)
    assert ok
    j = np.nonzero(assoc[0] >= 0)[0]; m = assoc[0][j]
    c0, s0 = np.cos(P[0, 2]), np.sin(P[0, 2]); R0 = np.array([[c0, -s0], [s0, c0]])
    p = sets[1]["mean"][j]
    q = sets[0]["mean"][m] @ R0.T + P[0, :2]
    if cost == "P2L":
        n = sets[0]["normal"][m] @ R0.T
        A = np.zeros((j.size, 2, 2)); A[:, 0, :] = n                     # second row empty: one scalar residual
        rows = 1
    else:
        C = sets[0]["cov"][m].reshape(-1, 2, 2)
        S = (0.1 * np.eye(2) + R0 @ C @ R0.T) * 1.0                      # (reg I + R C R^T) cov_scale   :292-296
        A = np.linalg.cholesky(np.linalg.inv(S))                         # L of the information; r = L e (not L^T)  :297-299
        rows = 2
    return P[1].copy(), p, q, A, rows, op[1], st


@pytest.mark.parametrize("cost", ["P2L", "P2D"])
@pytest.mark.parametrize("seed,offset", [(3, (0.15, -0.1, 0.01)), (5, (-0.4, 0.3, -0.02)), (8, (0.05, 0.02, 0.002)),
                                         (13, (0.8, -0.6, 0.03)),
                                         # far starts: rejected steps, shrinking radius, the 20-iteration cap (7 .. 20 iterations)
                                         (3, (2.0, 1.5, 0.08)), (5, (-3.0, 2.0, -0.1)), (8, (1.5, -2.5, 0.15))])
def test_oracle_lm_matches_an_independent_ceres_loop(orc, cost, seed, offset):
    x0, p, q, A, rows, x_orc, st = _problem(orc, cost, seed, offset)
    a = 0.1

    def fun(x, want_jac):
        c, s = np.cos(x[2]), np.sin(x[2])
        Rp = np.stack([c * p[:, 0] - s * p[:, 1], s * p[:, 0] + c * p[:, 1]], 1)
        e = Rp + x[:2] - q
        res = np.einsum("nij,nj->ni", A, e)[:, :rows]
        sq = (res * res).sum(1)
        rho, rho1 = _huber(sq, a)
        cost_ = 0.5 * rho.sum()
        if not want_jac:
            return cost_, None, None
        dRp = np.stack([-Rp[:, 1], Rp[:, 0]], 1)                          # d(R p)/d psi
        Je = np.zeros((p.shape[0], 2, 3)); Je[:, 0, 0] = 1; Je[:, 1, 1] = 1; Je[:, :, 2] = dRp
        Jr = np.einsum("nij,njk->nik", A, Je)[:, :rows, :]
        w = np.sqrt(rho1)                                                  # Corrector, rho'' <= 0: rows scaled by sqrt(rho')
        return cost_, (res * w[:, None]).ravel(), (Jr * w[:, None, None]).reshape(-1, 3)

    x, n_it, final_cost, _ = _ceres_lm(fun, x0)
    assert st.num_residuals == rows * p.shape[0] > 50
    assert n_it == st.inner_iterations, (n_it, st.inner_iterations)
    np.testing.assert_allclose(final_cost, st.final_cost, rtol=1e-9)
    assert np.hypot(*(x[:2] - x_orc[:2])) < 1e-9 and abs(x[2] - x_orc[2]) < 1e-10
