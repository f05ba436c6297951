"""The surface-point stage written a second time in numpy from the reference source and the published PCL algorithm, with
FLANN itself doing the radius search, against the oracle.  CPU only.

  Compensate                 utils.cpp:96-113, utils.h:28-32
  pcl::VoxelGrid             pointnormal.cpp:277-280 -> pcl/filters/impl/voxel_grid.hpp applyFilter (fp32 keys / centroids)
  radiusSearchT, >= 6        pointnormal.cpp:286-296 -> FLANN KDTreeSingleIndex through OpenCV's copy (tests/test_flann_pin.py)
  cell::cell / ComputeNormal pointnormal.cpp:7-63

Nothing of the oracle is used on the way (test_flann_pin.py still takes the voxel centroids from it): cloud in, cells out,
compared in order -- count, sample counts, means, covariances, normals, planarity, average intensity.
"""
import numpy as np
import pytest

import helpers

cv2 = pytest.importorskip("cv2")


def compensate_np(cloud, mot, ccw=False):
    out = cloud.copy()
    x, y = cloud[:, 0].astype(np.float64), cloud[:, 1].astype(np.float64)
    a = np.arctan2(y, x)                                                            # GetRelTimeStamp  utils.h:28-32
    d = np.where(a > 0.00001, a, 2 * np.pi + a) / (2 * np.pi)
    d = -(d - 0.5) if ccw else (d - 0.5)
    c, s = np.cos(d * mot[2]), np.sin(d * mot[2])                                   # getScaledRotationMatrix  utils.cpp:130-141
    out[:, 0] = (c * x - s * y + d * mot[0]).astype(np.float32)                     # R p + t, stored as float  utils.cpp:104-107
    out[:, 1] = (s * x + c * y + d * mot[1]).astype(np.float32)
    return out


def voxel_grid_np(cloud, leaf):
    """Centroids (x, y, intensity) in ascending voxel-index order; inside a voxel the points are summed in input order (the
    restatement's choice where PCL's unstable sort leaves it open), sequentially, in float32 like CentroidPoint."""
    f32 = np.float32
    inv = f32(1.0) / f32(leaf)
    xyz = cloud[:, :3]
    minb = np.floor(xyz.min(0) * inv).astype(np.int64)
    maxb = np.floor(xyz.max(0) * inv).astype(np.int64)
    div = maxb - minb + 1
    ijk = (np.floor(xyz * inv) - minb.astype(f32)).astype(np.int64)                 # floor(p * inv) - float(min_b), in float
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    cents = []
    i = 0
    while i < order.size:
        j = i
        sx = sy = si = f32(0)
        while j < order.size and idx[order[j]] == idx[order[i]]:
            pt = cloud[order[j]]
            sx = f32(sx + pt[0]); sy = f32(sy + pt[1]); si = f32(si + pt[3])
            j += 1
        n = f32(j - i)
        cents.append((f32(sx / n), f32(sy / n), f32(si / n)))
        i = j
    return np.array(cents, np.float32)


def cells_np(cloud, radius, weight_intensity=True, origin=(0.0, 0.0), downsample_factor=1.0):
    cents = voxel_grid_np(cloud, np.float32(np.float32(radius) / downsample_factor))          # pointnormal.cpp:279
    pts3 = np.ascontiguousarray(cloud[:, :3], np.float32)                           # search::KdTree<PointXYZI>: x, y, z
    index = cv2.flann_Index(pts3, {"algorithm": 4, "leaf_max_size": 15})
    r2 = float(np.float32(radius * radius))
    out = dict(mean=[], cov=[], normal=[], planarity=[], nsamples=[], avg_intensity=[])
    for c in cents:
        q = np.array([[c[0], c[1], 0.0]], np.float32)
        cnt, ind, _ = index.radiusSearch(q, r2, 8192, params={"checks": -1, "eps": 0.0, "sorted": True})
        if cnt < 6:                                                                 # pointnormal.cpp:291
            continue
        nb = ind[0, :cnt]
        P = cloud[nb, :2].astype(np.float64)
        w = np.maximum(cloud[nb, 3].astype(np.float64) - 60.0, 0.0) if weight_intensity else np.ones(cnt)   # :15
        W = w.sum()
        with np.errstate(all="ignore"):
            wn = w / W
            mu = (wn[:, None] * P).sum(0)                                           # :21-25
            X = P - mu
            C = X.T @ (wn[:, None] * X)                                             # :29-33
            if not np.all(np.isfinite(C)):
                continue
            lam, vec = np.linalg.eigh(C)                                            # SelfAdjointEigenSolver  :39-45
            cond = abs(lam[1] / lam[0]) if lam[0] != 0 else np.inf
        if not (cond <= 10000 and lam[0] * lam[1] > 0.00001 and lam[0] > 0 and lam[1] > 0):   # :53-56
            continue
        n = vec[:, 0]
        if n @ (np.asarray(origin) - mu) < 0:                                       # :59-61
            n = -n
        out["mean"].append(mu); out["cov"].append(C); out["normal"].append(n)
        out["planarity"].append(np.log(1 + cond / 2)); out["nsamples"].append(cnt); out["avg_intensity"].append(W / cnt)
    return {k: np.array(v) for k, v in out.items()}


@pytest.mark.parametrize("seed,radius,wint,mot", [(3, 3.5, True, None), (5, 3.0, True, (2.4, 0.1, 0.02)), (7, 3.0, False, None),
                                                  (9, 2.0, True, (-1.0, 0.3, -0.05))])
def test_oracle_surface_points_match_an_independent_numpy_flann_stage(orc, seed, radius, wint, mot):
    im, _ = helpers.scan_images(seed, 0)
    idx, cnt = orc.kstrongest(im[0], 60, 12)
    cloud = orc.cloud(im[0], idx, cnt)                                              # filter rows are pinned to the reference source
    if mot is not None:
        mine = compensate_np(cloud, np.asarray(mot))
        theirs = orc.compensate(cloud, mot)
        ulp = np.abs(mine[:, :2].view(np.int32).astype(np.int64) - theirs[:, :2].view(np.int32).astype(np.int64))
        assert ulp.max() <= 1 and (ulp > 0).mean() < 0.01                           # numpy's vectorised sin / cos vs glibc's
        cloud = theirs
    got = orc.surface_points(cloud, radius, wint)
    exp = cells_np(cloud, radius, wint)
    assert got["mean"].shape[0] == exp["mean"].shape[0] > 100
    assert np.array_equal(got["nsamples"], exp["nsamples"])
    np.testing.assert_allclose(got["mean"], exp["mean"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(got["cov"].reshape(-1, 2, 2), exp["cov"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got["normal"], exp["normal"], atol=1e-7)
    np.testing.assert_allclose(got["planarity"], exp["planarity"], rtol=1e-8)
    np.testing.assert_allclose(got["avg_intensity"], exp["avg_intensity"], rtol=1e-12)


def test_oracle_surface_points_with_a_downsample_factor(orc):
    """MapPointNormal::downsample_factor = 2 (pointnormal.cpp:5, :279): voxel leaf r / 2, four times the centroids, same radius."""
    im, _ = helpers.scan_images(4, 0)
    idx, cnt = orc.kstrongest(im[0], 60, 12)
    cloud = orc.cloud(im[0], idx, cnt)
    got = orc.surface_points(cloud, 3.0, True, downsample_factor=2.0)
    exp = cells_np(cloud, 3.0, True, downsample_factor=2.0)
    assert got["mean"].shape[0] == exp["mean"].shape[0] > 500
    assert np.array_equal(got["nsamples"], exp["nsamples"])
    np.testing.assert_allclose(got["mean"], exp["mean"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(got["cov"].reshape(-1, 2, 2), exp["cov"], rtol=1e-9, atol=1e-12)
