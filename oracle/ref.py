"""ctypes binding of oracle/_ref/libcfear_ref.so: the reference's OWN radar_filters.cpp / cfar.cpp, compiled unmodified
from /root/reference against interface stubs (oracle/ref_stubs, recipe: oracle/Makefile target `ref`).

TEST INFRASTRUCTURE ONLY (tests/ use it to pin the oracle restatement and the CUDA path to the reference source for
the k-strongest filter, its cloud, the peaks cloud and CA-CFAR).  The .so is built in the build container -- the only
place /root/reference exists -- and travels to the GPU box as a prebuilt file; nothing here reads /root/reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libcfear_ref.so")
_lib = None


def build() -> bool:
    """(Re)builds the library when the reference tree is present; returns availability."""
    if os.path.isdir("/root/reference/src/cfear_radarodometry"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return available()


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def kstrongest(img, z_min=60.0, k=12, min_distance=2.5, range_res=0.0438):
    """radarDriver::Process's k-strongest branch (radar_driver.cpp:57-61) on one image.
    Returns dict(idx [A,k], cnt [A], pidx, pcnt (peaks), cloud [n,4] f32, peaks [m,4] f32)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    A, R = img.shape
    idx = np.full((A, k), -1, np.int32); cnt = np.zeros(A, np.int32)
    pidx = np.full((A, k), -1, np.int32); pcnt = np.zeros(A, np.int32)
    cap = A * k
    cloud = np.zeros((cap, 4), np.float32); peaks = np.zeros((cap, 4), np.float32)
    nc = C.c_int32(0); npk = C.c_int32(0)
    rc = lib().ref_kstrongest(_p(img), A, R, C.c_float(z_min), int(k), C.c_float(min_distance), C.c_float(range_res),
                              _p(idx), _p(cnt), _p(pidx), _p(pcnt), _p(cloud), C.byref(nc), _p(peaks), C.byref(npk), cap)
    assert rc == 0
    return dict(idx=idx, cnt=cnt, pidx=pidx, pcnt=pcnt, cloud=cloud[:nc.value].copy(), peaks=peaks[:npk.value].copy())


def cfar(img, window_size=10, false_alarm_rate=0.01, nb_guard_cells=20, range_res=0.0438, z_min=60.0, min_distance=2.5,
         max_distance=400.0):
    """radarDriver::Process's CA-CFAR branch (radar_driver.cpp:52-56) on one image -> cloud [n,4] f32."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    A, R = img.shape
    out = np.zeros((A * R, 4), np.float32)
    n = lib().ref_cfar(_p(img), A, R, int(window_size), C.c_float(false_alarm_rate), int(nb_guard_cells), C.c_float(range_res),
                       C.c_float(z_min), C.c_float(min_distance), C.c_double(max_distance), _p(out), out.shape[0])
    return out[:n].copy()
