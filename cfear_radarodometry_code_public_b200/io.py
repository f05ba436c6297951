"""Polar radar image ingestion without ROS: the image conventions radarDriver expects.

  * CallbackOxford (radar_driver.cpp:99-111): the image is used as is -- 8UC1, rows = azimuths, cols = range bins.
    Oxford Radar RobotCar PNGs carry 11 metadata bytes per azimuth row (timestamp int64, sweep counter uint16,
    valid flag uint8) in front of 3768 range bins; `load_oxford_png` strips them.
  * Callback (radar_driver.cpp:74-90), every other dataset (e.g. MulRan polar PNGs, range x azimuth = 3360 x 400):
    MONO8, then cv::rotate(ROTATE_90_COUNTERCLOCKWISE) so that rows become azimuths; `load_range_azimuth_png` does that.

Images land in page-locked host buffers (capi.pinned_array) ready for cfear_odometry_step_batch / cfear_seq_step.
"""
from __future__ import annotations

import numpy as np

OXFORD_META_COLS = 11


def _imread_gray(path: str) -> np.ndarray:
    import cv2
    img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise FileNotFoundError(path)      # the reference exits on a NULL image (radar_driver.cpp:75-78)
    return img


def load_oxford_png(path: str):
    """Returns (polar uint8 [azimuths, 3768], timestamps int64 [azimuths], azimuth_rad float64 [azimuths], valid bool)."""
    raw = _imread_gray(path)
    ts = raw[:, :8].copy().view(np.int64).reshape(-1)
    sweep = raw[:, 8:10].copy().view(np.uint16).reshape(-1)
    valid = raw[:, 10] == 255
    az = sweep.astype(np.float64) * (2.0 * np.pi / 5600.0)      # encoder ticks -> radians
    return np.ascontiguousarray(raw[:, OXFORD_META_COLS:]), ts, az, valid


def load_range_azimuth_png(path: str) -> np.ndarray:
    """range x azimuth image -> azimuth x range, exactly cv::ROTATE_90_COUNTERCLOCKWISE (radar_driver.cpp:84)."""
    return np.ascontiguousarray(np.rot90(_imread_gray(path), 1))


def to_pinned_batch(images) -> np.ndarray:
    """Stack equally shaped polar images into one page-locked uint8 array [n, A, R]."""
    from . import capi
    images = list(images)
    out = capi.pinned_array((len(images),) + images[0].shape, np.uint8)
    for i, im in enumerate(images):
        out[i] = im
    return out


def write_frames(path: str, imgs, stamps_ns=None) -> None:
    """Raw frame file read by examples/offline_odometry.cpp (stands in for the reference's rosbag of sensor_msgs::Image):
    "CFRS" | int32 n, azimuths, range_bins | n x { uint64 stamp_ns | azimuths*range_bins uint8 }.  imgs: (n, A, R) uint8,
    rows = azimuths (what load_oxford_png / load_range_azimuth_png return)."""
    imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
    n, A, R = imgs.shape
    if stamps_ns is None:
        stamps_ns = np.arange(n, dtype=np.uint64) * np.uint64(250_000_000)           # 4 Hz
    with open(path, "wb") as f:
        f.write(b"CFRS")
        f.write(np.array([n, A, R], np.int32).tobytes())
        for i in range(n):
            f.write(np.uint64(stamps_ns[i]).tobytes())
            f.write(imgs[i].tobytes())
