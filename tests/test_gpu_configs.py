"""GPU parity on the BASELINE.json configurations the bench and the real data exercise:

  configs[2]  the first 64 problems of the EXACT bench workload (workload.make_batch(.., seed0=0), CFEAR3 parameters)
  configs[1]  scan-to-1-keyframe P2L on ~3000-cell sets: gn_fixed with 10 iterations and the ceres_lm loop
  Oxford      400 x 3768 images decoded from Oxford-format PNGs (11 metadata bytes per row) through the whole path

Bars as in test_gpu_parity.py: index sets / clouds / counts bit-exact, poses within 1e-4 m / 1e-5 rad of the oracle
after the same iteration counts.
"""
import os

import numpy as np
import pytest

from cfear_radarodometry_code_public_b200 import capi, workload
import helpers

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_form")]
POS_TOL, ROT_TOL = 1e-4, 1e-5


def _oracle_kf_sets(orc, kf_polar, radius, k=12, z_min=60):
    nprob, K = kf_polar.shape[:2]
    kf_sets, kf_ids = [], np.zeros((nprob, K), np.int32)
    for p in range(nprob):
        for i in range(K):
            ki, kc = orc.kstrongest(kf_polar[p, i], z_min, k)
            kf_ids[p, i] = len(kf_sets)
            kf_sets.append(orc.surface_points(orc.cloud(kf_polar[p, i], ki, kc), radius, True))
    return kf_sets, kf_ids


def test_bench_workload_first_64_problems_match_oracle(orc):
    """bench.py's workload, its first 64 problems, set up like bench.py sets it up (keyframe sets built by the GPU path)."""
    nprob, K = 64, 4
    b = workload.make_batch(nprob, K, seed0=0)
    ctx = capi.Context(max_batch=nprob, max_cellsets=nprob * (K + 1), max_keyframes=K, **workload.CFEAR3)
    kf_slots = np.arange(nprob * K, dtype=np.int32).reshape(nprob, K)
    cur = (nprob * K + np.arange(nprob)).astype(np.int32)
    for i in range(K):
        ctx.scans_to_cells_batch(b["kf_polar"][:, i], None, kf_slots[:, i])
    out = ctx.odometry_step_batch(b["polar"], b["mot"], kf_slots, cur, b["poses"])
    npts, ncells = ctx.last_counts(cur)
    kf_sets, kf_ids = _oracle_kf_sets(orc, b["kf_polar"], 3.0)
    ref = orc.pipeline_batch(b["polar"], b["mot"], kf_sets, kf_ids, b["poses"],
                             orc.reg_cfg(cost="P2D", loss="Huber", loss_limit=0.1, weight_opt=4, regularization=0.1, cov_scale=1.0),
                             k=12, z_min=60, radius=3.0, weight_intensity=True, compensate=True, nthreads=os.cpu_count() or 1)
    assert np.array_equal(npts, ref["npts"]) and np.array_equal(ncells, ref["ncells"])
    st = out["stats"]
    assert np.array_equal(st["outer_iterations"], [s.outer_iterations for s in ref["stats"]])
    assert np.array_equal(st["inner_iterations"], [s.inner_iterations for s in ref["stats"]])
    assert np.array_equal(st["num_residuals"], [s.num_residuals for s in ref["stats"]])
    assert np.array_equal(st["success"], [s.success for s in ref["stats"]])
    d = out["poses"][:, K] - ref["poses"][:, K]
    assert np.hypot(d[:, 0], d[:, 1]).max() < POS_TOL and np.abs(d[:, 2]).max() < ROT_TOL
    np.testing.assert_allclose(out["cov"].reshape(nprob, 6, 6), ref["cov"], rtol=1e-5, atol=1e-12)
    # k-strongest index sets of the bench images themselves, bit-exact
    idx, cnt = ctx.kstrongest(b["polar"][:8])
    for p in range(8):
        oi, oc = orc.kstrongest(b["polar"][p], 60, 12)
        assert np.array_equal(idx[p], oi) and np.array_equal(cnt[p], oc)
    ctx.close()


@pytest.mark.parametrize("solver", ["gn_fixed", "ceres_lm"])
@pytest.mark.parametrize("cost,wopt", [("P2L", 0), ("P2D", 4)])
def test_config1_registration_micro_3000_cells(orc, solver, cost, wopt):
    """BASELINE configs[1]: one keyframe, ~3000-cell sets, offset (0.5 m, 0.2 m, 2 deg), identity guess."""
    sets, P, delta = workload.make_cellset_pair(3000, seed=0)
    c = capi.Context(max_batch=2, max_cellsets=4, max_keyframes=1, cost=cost, loss="Huber", weight_opt=wopt,
                     solver_mode=solver, gn_iters=10, regularization=0.1 if cost == "P2D" else 1.0)
    c.cells_upload(0, sets[0]); c.cells_upload(1, sets[1])
    gp, gcov, gst, gassoc = c.register_batch(np.array([[0, 1]], np.int32), P[None], want_assoc=True)
    ocfg = orc.reg_cfg(cost=cost, loss="Huber", weight_opt=wopt, solver_mode=capi.SOLVER[solver], gn_iters=10,
                       regularization=0.1 if cost == "P2D" else 1.0)
    ok, op, ocov, ost, oassoc = orc.register(sets, P, ocfg, want_assoc=True)
    assert ok and gst["success"][0] == 1
    assert gst["outer_iterations"][0] == ost.outer_iterations and gst["inner_iterations"][0] == ost.inner_iterations
    assert gst["num_residuals"][0] == ost.num_residuals > 2000
    assert np.array_equal(gassoc[0, :, :3000], oassoc)
    d = gp[0, 1] - op[1]
    assert np.hypot(d[0], d[1]) < POS_TOL and abs(d[2]) < ROT_TOL, d
    e = gp[0, 1] - delta
    assert np.hypot(e[0], e[1]) < 0.02 and abs(e[2]) < 2e-3          # and it found the offset
    if solver == "gn_fixed":
        assert ost.inner_iterations == 10
    np.testing.assert_allclose(gcov[0], ocov, rtol=1e-5, atol=1e-12)
    c.close()


def test_oxford_width_png_through_the_whole_path(orc, tmp_path):
    """Oxford Radar RobotCar format: PNG rows of 11 metadata bytes + 3768 range bins.  load_oxford_png -> pinned batch ->
    filter / surface points / registration on a context configured for R = 3768 (rows are 8- but not 16-byte aligned)."""
    import cv2
    from cfear_radarodometry_code_public_b200 import io as cio, synth
    K, R, radius = 2, 3768, 3.0
    imgs, tp = synth.make_problem_images(77, K, R=R)
    paths = []
    for i in range(K + 1):
        raw = np.zeros((400, 11 + R), np.uint8)
        raw[:, :8] = (np.int64(1547131046353776000) + np.arange(400, dtype=np.int64) * 625000 + i * 250000000).view(np.uint8).reshape(400, 8)
        raw[:, 8:10] = (np.arange(400, dtype=np.uint16) * 14).view(np.uint8).reshape(400, 2)
        raw[:, 10] = 255
        raw[:, 11:] = imgs[i]
        p = str(tmp_path / f"{1547131046353776 + i}.png")
        assert cv2.imwrite(p, raw)
        paths.append(p)
    loaded = [cio.load_oxford_png(p) for p in paths]
    assert all(np.array_equal(l[0], imgs[i]) for i, l in enumerate(loaded)) and loaded[0][3].all()
    assert loaded[1][1][0] - loaded[0][1][0] == 250000000
    batch = cio.to_pinned_batch([l[0] for l in loaded])
    c = capi.Context(max_batch=K + 1, azimuths=400, range_bins=R, max_cellsets=K + 1, max_keyframes=K, radius=radius,
                     cost="P2D", loss="Huber", weight_opt=4, regularization=0.1)
    out = c.filter(batch, peaks=True)
    sets = []
    for i in range(K + 1):
        oi, oc = orc.kstrongest(imgs[i], 60, 12)
        assert np.array_equal(out["idx"][i], oi) and np.array_equal(out["cnt"][i], oc)
        ocl = orc.cloud(imgs[i], oi, oc)
        assert np.array_equal(out["clouds"][i].view(np.uint32), ocl.view(np.uint32))
        pi, pc = orc.peaks(imgs[i], oi, oc)
        assert np.array_equal(out["peaks"][i].view(np.uint32), orc.cloud(imgs[i], pi, pc).view(np.uint32))
        sets.append(orc.surface_points(ocl, radius, True))
    kf = np.arange(K, dtype=np.int32)[None]
    c.scans_to_cells_batch(batch[:K], None, kf[0])
    P = tp.copy(); P[K] = tp[K - 1]
    mot = synth.se2_mul(synth.se2_inv(tp[K - 2]), tp[K - 1])[None]
    got = c.odometry_step_batch(batch[K:K + 1], mot, kf, np.array([K], np.int32), P[None])
    ref = orc.pipeline_batch(imgs[K:K + 1], mot, sets[:K], kf, P[None], orc.reg_cfg(cost="P2D", loss="Huber", weight_opt=4, regularization=0.1), radius=radius)
    npts, ncells = c.last_counts(np.array([K], np.int32))
    assert npts[0] == ref["npts"][0] and ncells[0] == ref["ncells"][0] > 100
    assert got["stats"]["outer_iterations"][0] == ref["stats"][0].outer_iterations
    assert got["stats"]["num_residuals"][0] == ref["stats"][0].num_residuals
    d = got["poses"][0, K] - ref["poses"][0, K]
    assert np.hypot(d[0], d[1]) < POS_TOL and abs(d[2]) < ROT_TOL
    assert np.hypot(*(got["poses"][0, K, :2] - tp[K, :2])) < 0.3
    c.close()


def test_overlapped_device_steps_equal_stream_ordered_steps(orc):
    """cfear_odometry_step_batch_dev_submit: consecutive steps on the library's two internal streams (K1 / K3 of step i+1
    under K5 of step i), rotating two current-slot / result sets -- bit for bit what the stream-ordered call returns,
    whatever entry point is interleaved; tickets can be waited for from the host or from the context stream."""
    import torch
    K, nprob, nsets = 2, 6, 2
    c = capi.Context(max_batch=nprob, max_cellsets=nprob * (K + nsets), max_keyframes=K, cost="P2D", loss="Huber",
                     weight_opt=4, regularization=0.1, radius=3.0)
    b = workload.make_batch(nprob, K, seed0=300, workers=1)
    kf = np.arange(nprob * K, dtype=np.int32).reshape(nprob, K)
    for i in range(K):
        c.scans_to_cells_batch(b["kf_polar"][:, i], None, kf[:, i])
    curs = [(nprob * (K + j) + np.arange(nprob)).astype(np.int32) for j in range(nsets)]
    dev = torch.device("cuda", 0)
    ext = torch.cuda.ExternalStream(c.stream_ptr, device=dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    t_polar, t_mot, t_kf = t(b["polar"]), t(b["mot"]), t(kf)
    guesses = [b["poses"].copy() for _ in range(3)]
    guesses[1][:, K, 0] += 0.2; guesses[2][:, K, 2] -= 0.01
    t_cur = [t(x) for x in curs]
    t_pose = [torch.zeros(nprob, K + 1, 3, dtype=torch.float64, device=dev) for _ in range(nsets)]
    t_cov = [torch.zeros(nprob, 36, dtype=torch.float64, device=dev) for _ in range(nsets)]
    t_st = [torch.zeros(nprob, capi.STATS_DTYPE.itemsize, dtype=torch.uint8, device=dev) for _ in range(nsets)]

    def args(j):
        return (nprob, t_polar.data_ptr(), t_mot.data_ptr(), t_kf.data_ptr(), K, t_cur[j].data_ptr(), t_pose[j].data_ptr(),
                t_cov[j].data_ptr(), t_st[j].data_ptr())
    # stream-ordered reference results for the three guesses
    want = []
    for g in guesses:
        with torch.cuda.stream(ext):
            t_pose[0].copy_(t(g), non_blocking=True)
        torch.cuda.synchronize()
        c.odometry_step_batch_dev(*args(0))
        c.sync()
        want.append((t_pose[0].cpu().numpy().copy(), t_cov[0].cpu().numpy().copy(), t_st[0].cpu().numpy().copy()))
    assert not np.array_equal(want[0][0], want[1][0])
    # overlapped: 9 steps cycling through the guesses and the two sets
    tickets, which, got = [None] * nsets, [None] * nsets, []
    for s in range(9):
        j = s % nsets
        if tickets[j] is not None:
            if s % 3 == 0:
                c.odometry_step_batch_wait(tickets[j])            # host wait
            else:
                c.stream_wait_ticket(tickets[j])                  # device-side wait of the context stream
                torch.cuda.current_stream().wait_stream(ext)
            with torch.cuda.stream(ext):
                got.append((which[j], t_pose[j].clone(), t_cov[j].clone(), t_st[j].clone()))
        g = s % 3
        with torch.cuda.stream(ext):
            t_pose[j].copy_(t(guesses[g]), non_blocking=True)
        tickets[j], which[j] = c.odometry_step_batch_dev_submit(*args(j)), g
        if s == 4:
            c.kstrongest(b["polar"][:2])                          # any other entry point joins the steps in flight first
    c.join()
    with torch.cuda.stream(ext):
        for j in range(nsets):
            got.append((which[j], t_pose[j].clone(), t_cov[j].clone(), t_st[j].clone()))
    c.sync()
    assert len(got) == 9
    for g, p, cv, st in got:
        assert np.array_equal(p.cpu().numpy(), want[g][0]) and np.array_equal(cv.cpu().numpy(), want[g][1])
        assert np.array_equal(st.cpu().numpy(), want[g][2])
    npts, ncells = c.last_counts(curs[0])
    assert (npts > 1000).all() and (ncells > 100).all()
    c.close()
