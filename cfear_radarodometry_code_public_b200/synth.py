"""Seeded synthetic Navtech-shaped polar radar images.

World: 2-D line segments (street-canyon building rectangles + random short
segments).  Sensor trajectory: v = 10 m/s, omega = 0.1*sin(t/5) rad/s, 4 Hz.
Per azimuth a (theta = 2*pi*(a+1)/A, matching radar_filters.cpp:317 of the
reference) rays are cast to the first <=3 hits within range_res*R metres; the
image is a noise floor clip(N(30, 8^2)) plus, per hit (attenuated with hit order), a 7-bin
triangular return with peak U[80,200] times clipped Exp(1) speckle, plus Poisson(5)
isolated clutter returns per azimuth; clipped to uint8.

Input generator only -- carries no CFEAR algorithm.  RNG: numpy PCG64(seed).
"""
from __future__ import annotations

import numpy as np

A_DEFAULT, R_DEFAULT, RES_DEFAULT = 400, 3360, 0.0438


def make_world(seed: int, extent: float = 220.0) -> np.ndarray:
    """Returns segments [S,4] = (x0,y0,x1,y1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    segs = []
    # street-canyon blocks: a jittered lattice of rectangles, streets 20-60 m wide
    y = -extent
    while y < extent:
        h = rng.uniform(12, 35)
        x = -extent
        while x < extent:
            w = rng.uniform(12, 40)
            if abs(x + w / 2) > 9 or True:
                x0, y0, x1, y1 = x, y, x + w, y + h
                segs += [(x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)]
            x += w + rng.uniform(20, 60) * 0.5
        y += h + rng.uniform(20, 60) * 0.5
    segs = np.array(segs, np.float64)
    # keep the street the vehicle drives on clear (|y| < 7 m corridor along x)
    mid_y = 0.5 * (segs[:, 1] + segs[:, 3])
    lo = np.minimum(segs[:, 1], segs[:, 3]); hi = np.maximum(segs[:, 1], segs[:, 3])
    keep = ~((lo < 7) & (hi > -7))
    segs = segs[keep]
    del mid_y
    # random short segments (poles, cars, fences)
    n_short = 400
    c = rng.uniform(-extent, extent, (n_short, 2))
    c = c[np.abs(c[:, 1]) > 4]
    ang = rng.uniform(0, np.pi, c.shape[0]); ln = rng.uniform(1.0, 6.0, c.shape[0])
    d = np.stack([np.cos(ang), np.sin(ang)], 1) * ln[:, None] * 0.5
    short = np.concatenate([c - d, c + d], 1)
    return np.concatenate([segs, short], 0)


def trajectory(n: int, dt: float = 0.25, v: float = 10.0, t0: float = 0.0) -> np.ndarray:
    """Poses [n,3] (x,y,yaw), integrating v=10 m/s, omega=0.1*sin(t/5)."""
    poses = np.zeros((n, 3))
    x = y = th = 0.0
    sub = 20
    t = t0
    for i in range(n):
        poses[i] = (x, y, th)
        for _ in range(sub):
            h = dt / sub
            th += 0.1 * np.sin(t / 5.0) * h
            x += v * np.cos(th) * h; y += v * np.sin(th) * h
            t += h
    return poses


def render_polar(world: np.ndarray, pose, seed: int, A: int = A_DEFAULT, R: int = R_DEFAULT,
                 res: float = RES_DEFAULT, max_hits: int = 5) -> np.ndarray:
    """uint8 polar image [A,R]: rows azimuth, cols range bins."""
    rng = np.random.Generator(np.random.PCG64(seed))
    img = rng.standard_normal((A, R), dtype=np.float32) * 8.0 + 30.0
    theta = 2.0 * np.pi * (np.arange(A) + 1.0) / A + pose[2]
    d = np.stack([np.cos(theta), np.sin(theta)], 1)            # [A,2]
    o = np.array([pose[0], pose[1]])
    p0 = world[:, 0:2] - o; e = world[:, 2:4] - world[:, 0:2]  # [S,2]
    # ray o + t d hits segment p0 + u e :  t = cross(p0, e)/cross(d, e), u = cross(p0, d)/cross(d, e)
    den = d[:, None, 0] * e[None, :, 1] - d[:, None, 1] * e[None, :, 0]          # [A,S]
    den = np.where(np.abs(den) < 1e-12, np.nan, den)
    t = (p0[None, :, 0] * e[None, :, 1] - p0[None, :, 1] * e[None, :, 0]) / den
    u = (p0[None, :, 0] * d[:, None, 1] - p0[None, :, 1] * d[:, None, 0]) / den
    rmax = res * R
    ok = (t > 1.0) & (t < rmax - 1.0) & (u >= 0) & (u <= 1)
    t = np.where(ok, t, np.inf)
    t.sort(axis=1)
    hits = t[:, :max_hits]                                     # [A,max_hits]
    peak = rng.uniform(80, 200, hits.shape) * np.minimum(rng.exponential(1.0, hits.shape) + 0.3, 3.0)
    atten = np.array([1.0, 0.8, 0.65, 0.5, 0.4, 0.3, 0.25, 0.2])[:max_hits]
    tri = 1.0 - np.abs(np.arange(-3, 4)) / 4.0
    for h in range(hits.shape[1]):
        valid = np.isfinite(hits[:, h])
        rows = np.nonzero(valid)[0]
        if rows.size == 0:
            continue
        bins = np.rint(hits[rows, h] / res).astype(np.int64)
        cols = bins[:, None] + np.arange(-3, 4)[None, :]
        cols = np.clip(cols, 0, R - 1)
        val = (peak[rows, h] * atten[h])[:, None] * tri[None, :]
        np.maximum.at(img, (rows[:, None].repeat(7, 1), cols), val.astype(np.float32))
    # speckle / multipath clutter: Poisson(5) isolated returns per azimuth at U[55,95]
    ncl = rng.poisson(5.0, A)
    rows = np.repeat(np.arange(A), ncl)
    cols = rng.integers(40, R, rows.size)
    np.maximum.at(img, (rows, cols), rng.uniform(55, 95, rows.size).astype(np.float32))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def se2_mul(a, b):
    ca, sa = np.cos(a[2]), np.sin(a[2])
    return np.array([a[0] + ca * b[0] - sa * b[1], a[1] + sa * b[0] + ca * b[1], a[2] + b[2]])


def se2_inv(a):
    ca, sa = np.cos(a[2]), np.sin(a[2])
    return np.array([-(ca * a[0] + sa * a[1]), -(-sa * a[0] + ca * a[1]), -a[2]])


def make_problem_images(seed: int, n_keyframes: int = 4, A: int = A_DEFAULT, R: int = R_DEFAULT, res: float = RES_DEFAULT):
    """One independent (scan, K keyframes) problem: returns (images [K+1,A,R] u8, poses_true [K+1,3])."""
    world = make_world(seed)
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    poses = trajectory(n_keyframes + 1, t0=float(rng.uniform(0, 30)))
    imgs = np.stack([render_polar(world, poses[i], seed * 1000 + i, A, R, res) for i in range(n_keyframes + 1)])
    return imgs, poses


def make_sequence(seed: int, n_scans: int, A: int = A_DEFAULT, R: int = R_DEFAULT, res: float = RES_DEFAULT):
    """One synthetic Oxford-shaped (sub)sequence: n_scans images of one world along the 4 Hz trajectory.
    Returns (images [n,A,R] u8, poses_true [n,3] relative to the first pose)."""
    world = make_world(seed)
    rng = np.random.Generator(np.random.PCG64(seed + 104729))
    poses = trajectory(n_scans, t0=float(rng.uniform(0, 30)))
    imgs = np.stack([render_polar(world, poses[i], seed * 100003 + i, A, R, res) for i in range(n_scans)])
    return imgs, poses
