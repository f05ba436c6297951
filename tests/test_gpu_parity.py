"""GPU parity: the CUDA path (through the C ABI, libcfear_b200.so) against the CPU oracle on the same seeded inputs.

Bars (north_star): k-strongest index sets bit-exact; clouds bit-exact (fp32, table-driven cos/sin from host libm);
surface-point neighbour sets / counts exact, statistics to 1e-9 (fp64 summation order differs);
poses within 1e-4 m / 1e-5 rad of the oracle after the same iteration counts.
"""
import numpy as np
import pytest

from cfear_radarodometry_code_public_b200 import capi
import helpers

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_form")]

POS_TOL, ROT_TOL = 1e-4, 1e-5


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(max_batch=8, max_cellsets=16, max_keyframes=4)
    yield c
    c.close()


@pytest.fixture(scope="module")
def imgs():
    return helpers.scan_images(3)


def test_kstrongest_bit_exact_synthetic(ctx, orc, imgs):
    im, _ = imgs
    idx, cnt = ctx.kstrongest(im)
    for i in range(im.shape[0]):
        oi, oc = orc.kstrongest(im[i], 60, 12)
        assert np.array_equal(cnt[i], oc)
        assert np.array_equal(idx[i], oi)


def test_kstrongest_bit_exact_adversarial(ctx, orc):
    im = helpers.adversarial_image(1)
    idx, cnt = ctx.kstrongest(im[None])
    oi, oc = orc.kstrongest(im, 60, 12)
    assert np.array_equal(cnt[0], oc)
    assert np.array_equal(idx[0], oi)


@pytest.mark.parametrize("k,zmin", [(1, 60), (5, 0), (40, 60), (64, 200), (12, 255), (12, 128)])
def test_kstrongest_k_and_zmin_sweep(orc, k, zmin):
    c = capi.Context(max_batch=2, k_strongest=k, z_min=float(zmin), max_cellsets=2)
    im = np.stack([helpers.adversarial_image(2), helpers.scan_images(5, 0)[0][0]])
    idx, cnt = c.kstrongest(im)
    for i in range(2):
        oi, oc = orc.kstrongest(im[i], zmin, k)
        assert np.array_equal(cnt[i], oc)
        assert np.array_equal(idx[i], oi)
    c.close()


def pretest_stress_image(zmin, A=48, R=3360, seed=11):
    """Rows aimed at K1's streaming any-byte test (inside a word a byte >= 128 + z_min carries into its neighbour, so a
    neighbour equal to z_min - 1 looks like a candidate until the exact per-byte test of the drain) and at the queue of
    flagged vectors (more than 64 flagged vectors in one row force a drain in the middle of the row, with the candidate
    list still below its capacity)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lo = max(zmin - 1, 0)
    img = rng.integers(0, max(lo, 1), (A, R), dtype=np.uint8) if lo > 0 else np.zeros((A, R), np.uint8)
    img[0, 0::2] = 255; img[0, 1::2] = lo           # every byte pair (255, z_min - 1): saturated + false positives
    img[1, 7::16] = 255; img[1, 8::16] = lo         # one such pair per vector: 210 candidates, 210 flagged vectors
    img[2, 3::4] = lo; img[2, 2::128] = 255         # false positives only around a few real candidates
    img[3, 5::32] = min(zmin + 1, 255)              # 105 flagged vectors, one candidate each
    img[4, 15::48] = min(zmin, 255)                 # 70 flagged vectors: just past the forced-drain threshold
    img[5, :] = lo                                  # nothing, everything one below z_min
    img[6, ::3] = 255; img[6, 1::3] = lo            # carries into every third byte
    for a in range(7, A):                           # sparse real candidates beside (>= 128 + z_min, z_min - 1) pairs
        n = int(rng.integers(0, 30))
        cols = rng.integers(1, R - 1, n)
        img[a, cols] = rng.integers(min(zmin, 255), 256, n)
        cols = rng.integers(1, R - 1, 40)
        img[a, cols] = 255; img[a, cols + 1] = lo
    return img


@pytest.mark.parametrize("zmin", [1, 2, 60, 100, 127, 128, 129, 200, 255])
@pytest.mark.parametrize("R", [3360, 3768])
def test_kstrongest_carry_neighbours_and_queue_drains(orc, zmin, R):
    A = 48
    im = np.stack([pretest_stress_image(zmin, A, R, seed=11), pretest_stress_image(zmin, A, R, seed=12)])
    c = capi.Context(max_batch=2, azimuths=A, range_bins=R, z_min=float(zmin), max_cellsets=2)
    idx, cnt = c.kstrongest(im)
    for i in range(2):
        oi, oc = orc.kstrongest(im[i], zmin, 12)
        assert np.array_equal(cnt[i], oc)
        assert np.array_equal(idx[i], oi)
    c.close()


def test_kstrongest_odd_row_length(orc):
    # R not a multiple of 16: rows are not 16-byte aligned
    A, R = 37, 1001
    rng = np.random.Generator(np.random.PCG64(4))
    im = rng.integers(0, 256, (3, A, R), dtype=np.uint8)
    im[0] = (im[0] // 4)
    c = capi.Context(max_batch=3, azimuths=A, range_bins=R, max_cellsets=2)
    idx, cnt = c.kstrongest(im)
    for i in range(3):
        oi, oc = orc.kstrongest(im[i], 60, 12)
        assert np.array_equal(cnt[i], oc) and np.array_equal(idx[i], oi)
    c.close()


def test_cloud_and_peaks_bit_exact(ctx, orc, imgs):
    im, _ = imgs
    both = np.stack([im[0], helpers.adversarial_image(1)])
    out = ctx.filter(both, peaks=True)
    for i in range(2):
        oi, oc = orc.kstrongest(both[i], 60, 12)
        ocl = orc.cloud(both[i], oi, oc)
        assert out["npts"][i] == ocl.shape[0]
        assert np.array_equal(out["clouds"][i].view(np.uint32), ocl.view(np.uint32))
        pi, pc = orc.peaks(both[i], oi, oc)
        opk = orc.cloud(both[i], pi, pc)
        assert out["peaks"][i].shape == opk.shape
        assert np.array_equal(out["peaks"][i].view(np.uint32), opk.view(np.uint32))


def test_compensate(ctx, orc, imgs):
    im, _ = imgs
    oi, oc = orc.kstrongest(im[0], 60, 12)
    cl = orc.cloud(im[0], oi, oc)
    mot = np.array([2.5, 0.03, 0.025])
    for ccw in (False, True):
        g = ctx.compensate(cl, mot, ccw)
        o = orc.compensate(cl, mot, ccw)
        # fp64 atan2/sincos differ from glibc by <= 1-2 ulp(double); after rounding to fp32 that is a rare 1-ulp flip
        ulp = np.abs(g.view(np.int32).astype(np.int64) - o.view(np.int32).astype(np.int64))
        assert ulp.max() <= 1
        assert (ulp > 0).mean() < 1e-3
    z = ctx.compensate(cl, np.zeros(3))
    assert np.array_equal(z.view(np.uint32), cl.view(np.uint32))       # identity motion is exact


def _assert_cells_close(g, o):
    assert g["mean"].shape == o["mean"].shape
    assert np.array_equal(g["nsamples"], o["nsamples"])                # neighbour sets: exact fp32 radius test
    np.testing.assert_allclose(g["mean"], o["mean"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(g["cov"], o["cov"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(g["planarity"], o["planarity"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(g["avg_intensity"], o["avg_intensity"], rtol=1e-12)
    np.testing.assert_allclose(g["normal"], o["normal"], rtol=0, atol=1e-7)


@pytest.mark.parametrize("radius,wint", [(3.5, True), (3.0, True), (3.5, False), (2.0, True), (5.0, False)])
def test_surface_points(orc, imgs, radius, wint):
    im, _ = imgs
    c = capi.Context(max_batch=2, radius=radius, weight_intensity=int(wint), max_cellsets=4)
    for i in (0, 4):
        cl, o = helpers.oracle_cells(orc, im[i], radius=radius, weight_intensity=wint)
        n = c.surface_points(cl, 1)
        assert n == o["mean"].shape[0] and n > 50
        _assert_cells_close(c.cells_download(1), o)
    c.close()


def test_surface_points_edge_cases(ctx, orc):
    assert ctx.surface_points(np.zeros((0, 4), np.float32), 2) == 0     # reference exits on an empty cloud
    few = np.array([[10, 10, 0, 100], [10.5, 10, 0, 90], [11, 10.2, 0, 80]], np.float32)
    assert ctx.surface_points(few, 2) == 0                               # < 6 neighbours
    rng = np.random.Generator(np.random.PCG64(0))
    line = np.zeros((50, 4), np.float32); line[:, 0] = np.linspace(5, 8, 50); line[:, 1] = 3.0; line[:, 3] = 100
    o = orc.surface_points(line, 3.5)
    assert ctx.surface_points(line, 2) == o["mean"].shape[0] == 0        # degenerate covariance -> invalid
    dup = np.tile(np.array([[20, -7, 0, 61]], np.float32), (30, 1)); dup[:, :2] += rng.normal(0, 0.3, (30, 2)).astype(np.float32)
    dup[:, 3] = 60                                                        # all weights 0 -> NaN -> invalid
    assert ctx.surface_points(dup, 2) == orc.surface_points(dup, 3.5)["mean"].shape[0] == 0
    blob = np.zeros((200, 4), np.float32); blob[:, :2] = rng.normal(0, 4.0, (200, 2)) + [30, 30]; blob[:, 3] = rng.uniform(61, 200, 200)
    o = orc.surface_points(blob, 3.5)
    assert ctx.surface_points(blob, 2) == o["mean"].shape[0] > 0
    _assert_cells_close(ctx.cells_download(2), o)


def test_nearest_exact(ctx, orc, imgs):
    im, _ = imgs
    cl, o = helpers.oracle_cells(orc, im[1])
    ctx.cells_upload(3, o)
    rng = np.random.Generator(np.random.PCG64(9))
    q = np.concatenate([o["mean"] + rng.normal(0, 1.0, o["mean"].shape), rng.uniform(-160, 160, (2000, 2)),
                        o["mean"], np.array([[1e4, 1e4], [-1e4, 3.0]])])
    for radius in (2.0, 4.0, 0.5):
        assert np.array_equal(ctx.nearest(3, q, radius), orc.nearest(o["mean"], q, radius))
    # ties: duplicated means -> the smaller index wins
    d = dict(o); d = {k: np.concatenate([v, v]) for k, v in o.items()}
    ctx.cells_upload(3, d)
    assert np.array_equal(ctx.nearest(3, q, 2.0), orc.nearest(d["mean"], q, 2.0))


def _problem(orc, imgs, poses, K, radius):
    sets = [helpers.oracle_cells(orc, imgs[i], radius=radius)[1] for i in range(K + 1)]
    P = poses[:K + 1].copy()
    P[K] = P[K - 1]          # guess = previous pose (about 2.5 m / 1.4 deg off)
    return sets, P


@pytest.mark.parametrize("cost,loss,wopt,K,solver", [
    ("P2L", "Huber", 0, 1, "ceres_lm"), ("P2L", "Huber", 0, 1, "gn_fixed"), ("P2L", "Huber", 0, 4, "ceres_lm"),
    ("P2D", "Huber", 4, 4, "ceres_lm"), ("P2P", "Huber", 4, 4, "ceres_lm"), ("P2D", "Cauchy", 1, 3, "ceres_lm"),
    ("P2L", "None", 2, 2, "ceres_lm"), ("P2P", "SoftLOne", 3, 2, "gn_fixed"), ("P2D", "Combined", 0, 4, "ceres_lm"),
    ("P2L", "Tukey", 0, 2, "ceres_lm")])
def test_register_parity(orc, imgs, cost, loss, wopt, K, solver):
    im, poses = imgs
    radius = 3.0
    sets, P = _problem(orc, im, poses, K, radius)
    reg = 0.1 if cost == "P2D" else 1.0
    c = capi.Context(max_batch=2, max_cellsets=8, max_keyframes=4, cost=cost, loss=loss, weight_opt=wopt,
                     solver_mode=solver, regularization=reg, radius=radius)
    for i, s in enumerate(sets):
        c.cells_upload(i, s)
    slots = np.arange(K + 1, dtype=np.int32)[None]
    gp, gcov, gst, gassoc = c.register_batch(slots, P[None], want_assoc=True)
    ocfg = orc.reg_cfg(cost=cost, loss=loss, weight_opt=wopt, regularization=reg, solver_mode=capi.SOLVER[solver])
    ok, op, ocov, ost, oassoc = orc.register(sets, P, ocfg, want_assoc=True)
    assert bool(gst["success"][0]) == ok
    assert gst["outer_iterations"][0] == ost.outer_iterations
    assert gst["inner_iterations"][0] == ost.inner_iterations
    assert gst["num_residuals"][0] == ost.num_residuals
    d = gp[0, K] - op[K]
    assert np.hypot(d[0], d[1]) < POS_TOL and abs(d[2]) < ROT_TOL, d
    assert np.array_equal(gp[0, :K], P[:K])                       # keyframe blocks are constant
    n_src = sets[-1]["mean"].shape[0]
    assert np.array_equal(gassoc[0, :, :n_src], oassoc)
    np.testing.assert_allclose(gst["final_cost"][0], ost.final_cost, rtol=1e-6)
    np.testing.assert_allclose(gst["score"][0], ost.score, rtol=1e-6)
    np.testing.assert_allclose(gcov[0], ocov, rtol=1e-5, atol=1e-12)
    # the solve actually moved toward the truth
    if loss != "Tukey":      # Tukey(0.1) has zero gradient beyond 0.1 m: it (like the oracle) stays near the guess
        assert np.hypot(*(gp[0, K, :2] - poses[K, :2])) < 0.3
    c.close()


@pytest.mark.parametrize("cost,wopt,K", [("P2D", 4, 3), ("P2L", 1, 1), ("P2P", 2, 2)])
def test_register_soft_prior_and_association_tables(orc, imgs, cost, wopt, K):
    """cfear_register_batch_ex: Register(..., soft_constraints = true) (prior block alpha L (guess - x), n_scan_normal.cpp:373-377)
    and the content of scan_associations_ / weight_associations_ (target index + direction similarity per source cell)."""
    im, poses = imgs
    sets, P = _problem(orc, im, poses, K, 3.0)
    reg = 0.1 if cost == "P2D" else 1.0
    c = capi.Context(max_batch=2, max_cellsets=8, max_keyframes=4, cost=cost, loss="Huber", weight_opt=wopt, regularization=reg, radius=3.0)
    for i, s in enumerate(sets):
        c.cells_upload(i, s)
    slots = np.arange(K + 1, dtype=np.int32)[None]
    ocfg = orc.reg_cfg(cost=cost, loss="Huber", weight_opt=wopt, regularization=reg)
    c6 = np.eye(6); c6[0, 0] = 4.0; c6[1, 1] = 9.0; c6[0, 1] = c6[1, 0] = 1.0; c6[5, 5] = 0.04; c6[0, 5] = c6[5, 0] = 0.1
    n_src = sets[-1]["mean"].shape[0]
    for L in (None, orc.prior_sqrt_information(c6), np.eye(3)):
        gp, gcov, gst, gassoc, gsim = c.register_batch(slots, P[None], want_sim=True, prior_sqrt_info=None if L is None else L[None])
        ok, op, ocov, ost, oassoc, osim = orc.register(sets, P, ocfg, prior_sqrt_info=L, want_sim=True)
        assert bool(gst["success"][0]) == ok and gst["pose_written"][0] == ost.pose_written == 1
        assert gst["outer_iterations"][0] == ost.outer_iterations and gst["inner_iterations"][0] == ost.inner_iterations
        assert gst["num_residuals"][0] == ost.num_residuals and gst["num_blocks"][0] == ost.num_blocks
        d = gp[0, K] - op[K]
        assert np.hypot(d[0], d[1]) < POS_TOL and abs(d[2]) < ROT_TOL, d
        assert np.array_equal(gassoc[0, :, :n_src], oassoc)
        np.testing.assert_allclose(gsim[0, :, :n_src], osim, rtol=0, atol=1e-13)
        assert (gsim[0, :, :n_src][oassoc >= 0] > np.cos(np.pi / 6)).all() and (gsim[0, :, :n_src][oassoc < 0] == 0).all()
        np.testing.assert_allclose(gst["final_cost"][0], ost.final_cost, rtol=1e-6)
        np.testing.assert_allclose(gcov[0], ocov, rtol=1e-5, atol=1e-12)
    nores = c.register_batch(slots, P[None])                        # the plain call is unaffected by the calls before
    ok, op, _, ost, _ = orc.register(sets, P, ocfg)
    assert nores[2]["num_residuals"][0] == ost.num_residuals and np.abs(nores[0][0, K] - op[K]).max() < 1e-9
    c.close()


def test_get_cost_batch_matches_oracle(orc, imgs):
    """cfear_get_cost_batch (n_scan_normal_reg::GetCost): 27 pose samples around a registered pose in one launch vs one
    oracle GetCost per sample; a sample that sees nothing returns ok = 0."""
    K = 3
    c = capi.Context(max_batch=32, max_cellsets=K + 1, max_keyframes=K, cost="P2D", loss="Huber", weight_opt=4, regularization=0.1)
    sets = []
    for i in range(K + 1):
        cl, s = helpers.oracle_cells(orc, imgs[0][i], radius=3.5)
        sets.append(s); c.cells_upload(i, s)
    P = imgs[1][:K + 1].copy(); P[K] = imgs[1][K - 1]
    cfg = orc.reg_cfg(cost="P2D", loss="Huber", weight_opt=4, regularization=0.1)
    ok, p, _, st, _ = orc.register(sets, P, cfg)
    assert ok
    xs = np.linspace(-0.2, 0.2, 3); ts = np.linspace(-0.00218125, 0.00218125, 3)
    samples = [p.copy() for _ in range(28)]
    n = 0
    for t in ts:
        for x in xs:
            for y in xs:
                samples[n][K] = p[K] + [x, y, t]; n += 1
    samples[27][K, :2] += 1e4                                       # nothing in reach: GetCost returns false
    cost, nres, okv = c.get_cost_batch(np.tile(np.arange(K + 1, dtype=np.int32), (28, 1)), np.stack(samples))
    for i in range(28):
        o_ok, o_cost, o_nres = orc.get_cost(sets, samples[i], cfg)
        assert bool(okv[i]) == o_ok and nres[i] == o_nres
        if o_ok:
            np.testing.assert_allclose(cost[i], o_cost, rtol=1e-11)
    assert okv[:27].all() and not okv[27]
    # the pose table is an input only, and a following Register is unaffected by the cost-only launch
    p2, _, st2, _ = c.register_batch(np.arange(K + 1, dtype=np.int32)[None], P[None])
    assert st2["outer_iterations"][0] == st.outer_iterations and np.abs(p2[0, K] - p[K]).max() < 1e-9
    c.close()


def test_register_failure_modes(orc, imgs):
    im, poses = imgs
    sets, P = _problem(orc, im, poses, 1, 3.0)
    c = capi.Context(max_batch=1, max_cellsets=4, max_keyframes=2)
    c.cells_upload(0, sets[0]); c.cells_upload(1, sets[1])
    far = P.copy(); far[1] = [500.0, 500.0, 0.0]                   # no correspondences -> Register() returns false
    gp, gcov, gst = c.register([0, 1], far)
    ok, op, ocov, ost, _ = orc.register(sets[:2], far, orc.reg_cfg())
    assert not ok and gst["success"] == 0 and gst["num_residuals"] == ost.num_residuals == 0
    assert np.array_equal(gp, far)                                 # pose untouched
    assert np.array_equal(gcov, ocov)
    with pytest.raises(capi.CfearError):
        c.register([0, 9], P)                                      # bad slot
    with pytest.raises(capi.CfearError):
        c.register([0], P[:1])                                     # needs >= 2 scans (n_scan_normal.cpp:190)
    c.close()


def test_odometry_step_batch_matches_oracle_pipeline(orc):
    """Whole path on host images: filter -> compensate -> surface points -> register, 6 independent problems."""
    K, radius, nprob = 4, 3.0, 6
    c = capi.Context(max_batch=nprob * (K + 1), max_cellsets=nprob * (K + 1), max_keyframes=K, cost="P2D", loss="Huber",
                     weight_opt=4, regularization=0.1, radius=radius)
    ocfg = orc.reg_cfg(cost="P2D", loss="Huber", weight_opt=4, regularization=0.1)
    polar, kf_sets, kf_ids, poses, mot = [], [], [], [], []
    for b in range(nprob):
        im, tp = helpers.scan_images(20 + b, K)
        ids = []
        for i in range(K):
            ids.append(len(kf_sets))
            kf_sets.append(helpers.oracle_cells(orc, im[i], radius=radius)[1])
        kf_ids.append(ids)
        polar.append(im[K])
        P = tp.copy(); P[K] = tp[K - 1]
        poses.append(P)
        from cfear_radarodometry_code_public_b200.synth import se2_inv, se2_mul
        mot.append(se2_mul(se2_inv(tp[K - 2]), tp[K - 1]))
    polar = np.stack(polar); poses = np.stack(poses); mot = np.stack(mot); kf_ids = np.array(kf_ids, np.int32)
    ref = orc.pipeline_batch(polar, mot, kf_sets, kf_ids, poses, ocfg, radius=radius)
    # keyframe cell sets are built by the GPU path itself from the keyframe images
    for b in range(nprob):
        im, _ = helpers.scan_images(20 + b, K)
        for i in range(K):
            out = c.filter(im[i][None])
            c.surface_points(out["clouds"][0], kf_ids[b, i])
    cur = np.arange(nprob, dtype=np.int32) + nprob * K
    out = c.odometry_step_batch(polar, mot, kf_ids, cur, poses)
    npts, ncells = c.last_counts(cur)
    assert np.array_equal(npts, ref["npts"]) and np.array_equal(ncells, ref["ncells"])
    for b in range(nprob):
        d = out["poses"][b, K] - ref["poses"][b, K]
        assert np.hypot(d[0], d[1]) < POS_TOL and abs(d[2]) < ROT_TOL, (b, d)
        assert out["stats"]["outer_iterations"][b] == ref["stats"][b].outer_iterations
        assert out["stats"]["num_residuals"][b] == ref["stats"][b].num_residuals
    c.close()


def test_submit_wait_pipelining_matches_synchronous_calls(orc):
    """cfear_odometry_step_batch_submit/_wait with two steps in flight (two sub-batches each, another entry point writing
    the image staging area in between) returns bit for bit what the synchronous call returns."""
    K, radius, base, rep = 2, 3.0, 3, 12
    nprob = base * rep                       # 36 problems -> two sub-batches of the host-buffer path
    c = capi.Context(max_batch=nprob, max_cellsets=base * K + nprob, max_keyframes=K, cost="P2D", loss="Huber",
                     weight_opt=4, regularization=0.1, radius=radius)
    polar, poses, mot, kf = [], [], [], []
    for b in range(base):
        im, tp = helpers.scan_images(40 + b, K)
        for i in range(K):
            c.surface_points(c.filter(im[i][None])["clouds"][0], b * K + i)
        polar.append(im[K]); P = tp.copy(); P[K] = tp[K - 1]; poses.append(P); mot.append(np.zeros(3)); kf.append([b * K, b * K + 1])
    polar = np.ascontiguousarray(np.tile(np.stack(polar), (rep, 1, 1)))
    kf = np.ascontiguousarray(np.tile(np.array(kf, np.int32), (rep, 1)))
    mot = np.ascontiguousarray(np.tile(np.stack(mot), (rep, 1)))
    cur = (base * K + np.arange(nprob)).astype(np.int32)
    posesA = np.ascontiguousarray(np.tile(np.stack(poses), (rep, 1, 1)))
    posesB = posesA.copy(); posesB[:, K, 0] += 0.3; posesB[:, K, 2] -= 0.01      # a second step with another guess
    refA = c.odometry_step_batch(polar, mot, kf, cur, posesA)
    refB = c.odometry_step_batch(polar, mot, kf, cur, posesB)
    c.filter(polar[:2])                       # writes the staging area on the compute stream (polar_dirty path)

    def outs(p):
        return dict(poses=p.copy(), cov=np.zeros((nprob, 36)), stats=np.zeros(nprob, capi.STATS_DTYPE), npts=np.zeros(nprob, np.int32))
    oA, oB, oA2 = outs(posesA), outs(posesB), outs(posesA)
    tA = c.odometry_step_batch_submit(polar, mot, kf, cur, oA)
    tB = c.odometry_step_batch_submit(polar, mot, kf, cur, oB)
    c.odometry_step_batch_wait(tA)
    tA2 = c.odometry_step_batch_submit(polar, mot, kf, cur, oA2)
    c.odometry_step_batch_wait(tB)
    c.odometry_step_batch_wait(tA2)
    for o, r in ((oA, refA), (oB, refB), (oA2, refA)):
        assert np.array_equal(o["poses"], r["poses"]) and np.array_equal(o["cov"], r["cov"])
        assert np.array_equal(o["stats"], r["stats"]) and np.array_equal(o["npts"], r["npts"])
    assert not np.array_equal(refA["poses"][:, K], posesA[:, K])
    with pytest.raises(capi.CfearError):
        c.odometry_step_batch_wait(99)
    c.close()


@pytest.mark.parametrize("cost,wopt,submap,wint,radius", [("P2L", 0, 3, True, 3.5), ("P2D", 4, 4, True, 3.0)])
def test_sequences_lockstep_match_oracle_replay(orc, cost, wopt, submap, wint, radius):
    """cfear_seq_*: OdometryKeyframeFuser bookkeeping on the device, 3 sequences advancing together, no host sync per
    step.  Every sequence must reproduce the oracle's sequential replay: keyframe decisions, iteration counts, poses."""
    from cfear_radarodometry_code_public_b200 import synth
    nseq, nsteps, kmax = 3, 10, 4
    seqs = [synth.make_sequence(40 + b, nsteps)[0] for b in range(nseq)]
    reg = 0.1 if cost == "P2D" else 0.0
    c = capi.Context(max_batch=nseq, max_cellsets=nseq * (kmax + 1), max_keyframes=kmax, cost=cost, weight_opt=wopt,
                     regularization=reg, radius=radius, weight_intensity=int(wint))
    S = capi.Sequences(c, nseq, nsteps, submap_scan_size=submap)
    for t in range(nsteps):
        S.step(np.stack([seqs[b][t] for b in range(nseq)]))
    poses, kf, st = S.read(0, nsteps)
    for b in range(nseq):
        ref = orc.odometry_sequence(seqs[b], orc.reg_cfg(cost=cost, weight_opt=wopt, regularization=reg), radius=radius,
                                    weight_intensity=wint, submap_scan_size=submap)
        assert np.array_equal(kf[b], ref["keyframe"])
        assert np.array_equal(st[b]["outer_iterations"], [s.outer_iterations for s in ref["stats"]])
        assert np.array_equal(st[b]["num_residuals"], [s.num_residuals for s in ref["stats"]])
        d = poses[b] - ref["poses"]
        assert np.hypot(d[:, 0], d[:, 1]).max() < POS_TOL and np.abs(d[:, 2]).max() < ROT_TOL, (b, d)
    S.close(); c.close()


def test_large_k_global_fallback_paths(orc):
    """k=40 (the reference's CFEAR-3 preset): 16000-point clouds do not fit shared memory, so K3 runs on its global
    scratch buffers; r=1.0 makes the voxel grid exceed the shared histogram (global histogram fallback)."""
    im = helpers.scan_images(8, 1)[0]
    c = capi.Context(max_batch=2, k_strongest=40, max_cellsets=4, max_keyframes=2, radius=3.0)
    out = c.filter(im)
    sets = []
    for i in range(2):
        oi, oc = orc.kstrongest(im[i], 60, 40)
        assert np.array_equal(out["idx"][i], oi)
        ocl = orc.cloud(im[i], oi, oc)
        assert np.array_equal(out["clouds"][i].view(np.uint32), ocl.view(np.uint32))
        o = orc.surface_points(ocl, 3.0, True)
        assert c.surface_points(ocl, i) == o["mean"].shape[0] > 100
        _assert_cells_close(c.cells_download(i), o)
        sets.append(o)
    P = np.array([[0, 0, 0], [0.5, 0.1, 0.01]], np.float64)
    gp, gcov, gst = c.register([0, 1], P)
    ok, op, ocov, ost, _ = orc.register(sets, P, orc.reg_cfg())
    assert bool(gst["success"]) == ok and gst["outer_iterations"] == ost.outer_iterations
    assert np.hypot(*(gp[1, :2] - op[1, :2])) < POS_TOL and abs(gp[1, 2] - op[1, 2]) < ROT_TOL
    # whole path with the fused kernel chain (mode 0 of K3 on global scratch)
    npts, nc = c.scans_to_cells_batch(im, None, [2, 3])
    assert nc[0] == sets[0]["mean"].shape[0] and nc[1] == sets[1]["mean"].shape[0]
    c.close()
    c = capi.Context(max_batch=1, max_cellsets=2, radius=1.0)
    oi, oc = orc.kstrongest(im[0], 60, 12)
    ocl = orc.cloud(im[0], oi, oc)
    o = orc.surface_points(ocl, 1.0, True)
    assert c.surface_points(ocl, 0) == o["mean"].shape[0]
    _assert_cells_close(c.cells_download(0), o)
    c.close()


def test_small_max_cells_and_capacity_errors(orc):
    im = helpers.scan_images(8, 0)[0]
    oi, oc = orc.kstrongest(im[0], 60, 12)
    ocl = orc.cloud(im[0], oi, oc)
    o = orc.surface_points(ocl, 3.5, True)
    n = o["mean"].shape[0]
    c = capi.Context(max_batch=1, max_cellsets=2, max_cells=100)          # fewer cells than the scan produces: truncated, in order
    assert c.surface_points(ocl, 0) == 100 < n
    g = c.cells_download(0)
    np.testing.assert_allclose(g["mean"], o["mean"][:100], atol=1e-9)
    with pytest.raises(capi.CfearError):
        c.cells_upload(1, o)                                              # exceeds max_cells
    with pytest.raises(capi.CfearError):
        c.filter(np.zeros((2, 400, 3360), np.uint8))                      # exceeds max_batch
    c.close()


@pytest.mark.parametrize("window,guard,far", [(10, 20, 0.01), (40, 5, 0.01), (3, 0, 0.2), (10, 70, 0.001)])
def test_cfar_filter_bit_exact(ctx, orc, imgs, window, guard, far):
    """CA-CFAR (the reference's alternative filter): detections and fp32 cloud identical to the oracle."""
    im, _ = imgs
    both = np.stack([im[0], helpers.adversarial_image(5)])
    got = ctx.cfar_filter(both, window_size=window, nb_guard_cells=guard, false_alarm_rate=far, capacity=400 * 3360 // 4)
    for i in range(2):
        ref = orc.cfar(both[i], window_size=window, nb_guard_cells=guard, false_alarm_rate=far)
        assert got[i].shape == ref.shape and ref.shape[0] > 0
        assert np.array_equal(got[i].view(np.uint32), ref.view(np.uint32))
    with pytest.raises(capi.CfearError):
        ctx.cfar_filter(both[:1], window_size=3, nb_guard_cells=0, false_alarm_rate=0.2, capacity=5)


def test_degenerate_images_through_the_whole_path(orc):
    """Empty, saturated and pure-noise images must neither crash nor diverge from the oracle: empty clouds give empty
    cell sets and a failed registration that leaves the guess untouched (the fuser then keeps the guess)."""
    K, radius = 2, 3.0
    base = helpers.scan_images(9, K)[0]
    rng = np.random.Generator(np.random.PCG64(3))
    cur_imgs = np.stack([np.zeros((400, 3360), np.uint8), np.full((400, 3360), 255, np.uint8),
                         rng.integers(0, 256, (400, 3360), dtype=np.uint8), base[K]])
    n = cur_imgs.shape[0]
    c = capi.Context(max_batch=n, max_cellsets=K + n, max_keyframes=K, radius=radius, cost="P2L")
    kf_sets = []
    for i in range(K):
        c.scans_to_cells_batch(base[i][None], None, [i])
        kf_sets.append(helpers.oracle_cells(orc, base[i], radius=radius)[1])
    kf = np.tile(np.arange(K, dtype=np.int32), (n, 1)); cur = (K + np.arange(n)).astype(np.int32)
    poses = np.tile(np.array([[0, 0, 0], [2.5, 0, 0.02], [5.0, 0.1, 0.04]], np.float64), (n, 1, 1))
    mot = np.tile(np.array([2.5, 0.0, 0.02]), (n, 1))
    out = c.odometry_step_batch(cur_imgs, mot, kf, cur, poses)
    ref = orc.pipeline_batch(cur_imgs, mot, kf_sets, np.tile(np.arange(K, dtype=np.int32), (n, 1)), poses, orc.reg_cfg(cost="P2L"), radius=radius)
    npts, ncells = c.last_counts(cur)
    assert np.array_equal(npts, ref["npts"]) and np.array_equal(ncells, ref["ncells"])
    assert npts[0] == 0 and ncells[0] == 0 and npts[1] == 400 * 12
    for b in range(n):
        assert out["stats"]["success"][b] == ref["stats"][b].success
        assert out["stats"]["num_residuals"][b] == ref["stats"][b].num_residuals
        d = out["poses"][b, K] - ref["poses"][b, K]
        assert np.hypot(d[0], d[1]) < POS_TOL and abs(d[2]) < ROT_TOL
    assert out["stats"]["success"][0] == 0 and np.array_equal(out["poses"][0], poses[0])     # nothing to register: guess untouched
    assert out["stats"]["success"][3] == 1
    c.close()
