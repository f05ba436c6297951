"""ctypes binding of libcfear_b200.so (include/cfear_b200.h) -- plumbing for tests/, bench.py and replay scripts.

The product is the C-ABI library and the C++ host mirror (include/cfear_b200.hpp); this module only marshals
numpy buffers into it.  There is no CPU fallback: if the library is missing or no CUDA device is present the
calls raise.  This module never imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import weakref
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CFEAR_B200_LIB: experiment builds (profiles/ab); the shipped library otherwise
LIB_PATH = os.environ.get("CFEAR_B200_LIB") or os.path.join(_HERE, "libcfear_b200.so")

COST = {"P2P": 0, "P2L": 1, "P2D": 2}
LOSS = {"None": 0, "Huber": 1, "Cauchy": 2, "SoftLOne": 3, "Combined": 4, "Tukey": 5}
SOLVER = {"ceres_lm": 0, "gn_fixed": 1}


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_batch", C.c_int32), ("azimuths", C.c_int32), ("range_bins", C.c_int32),
                ("k_strongest", C.c_int32), ("z_min", C.c_float), ("range_res", C.c_float), ("min_distance", C.c_float),
                ("radius", C.c_float), ("downsample_factor", C.c_double), ("weight_intensity", C.c_int32),
                ("compensate", C.c_int32), ("radar_ccw", C.c_int32), ("cost", C.c_int32), ("loss", C.c_int32),
                ("weight_opt", C.c_int32), ("solver_mode", C.c_int32), ("loss_limit", C.c_double),
                ("cov_scale", C.c_double), ("regularization", C.c_double), ("reg_radius", C.c_double),
                ("max_outer", C.c_int32), ("min_outer", C.c_int32), ("max_inner", C.c_int32), ("gn_iters", C.c_int32),
                ("max_keyframes", C.c_int32), ("max_cellsets", C.c_int32), ("max_cells", C.c_int32),
                ("steps_in_flight", C.c_int32)]


class RegStats(C.Structure):
    _fields_ = [("success", C.c_int32), ("outer_iterations", C.c_int32), ("inner_iterations", C.c_int32),
                ("num_residuals", C.c_int32), ("num_blocks", C.c_int32), ("usable", C.c_int32),
                ("final_cost", C.c_double), ("score", C.c_double), ("pose_written", C.c_int32), ("reserved", C.c_int32)]


class CfarParams(C.Structure):
    _fields_ = [("window_size", C.c_int32), ("nb_guard_cells", C.c_int32), ("false_alarm_rate", C.c_double), ("max_distance", C.c_double)]


class SeqParams(C.Structure):
    _fields_ = [("submap_scan_size", C.c_int32), ("use_guess", C.c_int32), ("use_keyframe", C.c_int32), ("reserved", C.c_int32),
                ("min_keyframe_dist", C.c_double), ("min_keyframe_rot_deg", C.c_double)]


CELL_DTYPE = np.dtype([("mean", np.float64, 2), ("normal", np.float64, 2), ("cov", np.float64, 4),
                       ("planarity", np.float64), ("avg_intensity", np.float64), ("nsamples", np.int32),
                       ("pad", np.int32)])
STATS_DTYPE = np.dtype([("success", np.int32), ("outer_iterations", np.int32), ("inner_iterations", np.int32),
                        ("num_residuals", np.int32), ("num_blocks", np.int32), ("usable", np.int32),
                        ("final_cost", np.float64), ("score", np.float64), ("pose_written", np.int32), ("reserved", np.int32)])

# every symbol include/cfear_b200.h declares
SYMBOLS = ["cfear_default_config", "cfear_create", "cfear_destroy", "cfear_update_config", "cfear_last_error", "cfear_version",
           "cfear_launch_count", "cfear_kstrongest", "cfear_filter", "cfear_compensate", "cfear_surface_points",
           "cfear_scans_to_cells_batch", "cfear_cells_count", "cfear_cells_download", "cfear_cells_upload", "cfear_nearest", "cfear_register",
           "cfear_register_batch", "cfear_register_batch_ex", "cfear_get_cost_batch", "cfear_odometry_step_batch", "cfear_odometry_step_batch_submit", "cfear_odometry_step_batch_wait", "cfear_odometry_step_batch_dev", "cfear_odometry_step_batch_dev_submit", "cfear_stream_wait_ticket", "cfear_join", "cfear_sync",
           "cfear_stream", "cfear_stage_timing", "cfear_last_counts", "cfear_alloc_pinned", "cfear_free_pinned",
           "cfear_alloc_device", "cfear_free_device", "cfear_memcpy_h2d", "cfear_memcpy_d2h",
           "cfear_cfar_filter", "cfear_seq_create", "cfear_seq_destroy", "cfear_seq_step", "cfear_seq_step_dev", "cfear_seq_read"]

_lib = None


class CfearError(RuntimeError):
    pass


def load():
    """Load the C-ABI library.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CfearError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.cfear_last_error.restype = C.c_char_p
        lib.cfear_version.restype = C.c_char_p
        lib.cfear_launch_count.restype = C.c_int64
        lib.cfear_launch_count.argtypes = [C.c_void_p]
        lib.cfear_stream.restype = C.c_void_p
        lib.cfear_stream.argtypes = [C.c_void_p]
        lib.cfear_alloc_pinned.restype = C.c_void_p
        lib.cfear_alloc_pinned.argtypes = [C.c_size_t]
        lib.cfear_free_pinned.argtypes = [C.c_void_p]
        lib.cfear_alloc_device.restype = C.c_void_p
        lib.cfear_alloc_device.argtypes = [C.c_void_p, C.c_size_t]
        lib.cfear_free_device.argtypes = [C.c_void_p, C.c_void_p]
        lib.cfear_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.cfear_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.cfear_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        lib.cfear_destroy.argtypes = [C.c_void_p]
        lib.cfear_update_config.argtypes = [C.c_void_p, C.POINTER(Config)]
        lib.cfear_sync.argtypes = [C.c_void_p]
        vp, i32 = C.c_void_p, C.c_int
        lib.cfear_kstrongest.argtypes = [vp, vp, i32, vp, vp]
        lib.cfear_filter.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp]
        lib.cfear_compensate.argtypes = [vp, vp, i32, vp, i32]
        lib.cfear_surface_points.argtypes = [vp, vp, i32, i32, vp]
        lib.cfear_scans_to_cells_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        lib.cfear_cells_count.argtypes = [vp, i32, vp]
        lib.cfear_cells_download.argtypes = [vp, i32, vp, i32, vp]
        lib.cfear_cells_upload.argtypes = [vp, i32, vp, i32]
        lib.cfear_nearest.argtypes = [vp, i32, vp, i32, C.c_double, vp]
        lib.cfear_register.argtypes = [vp, vp, i32, vp, vp, vp]
        lib.cfear_register_batch.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp]
        lib.cfear_register_batch_ex.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp]
        lib.cfear_get_cost_batch.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp]
        lib.cfear_odometry_step_batch.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]
        lib.cfear_odometry_step_batch_submit.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]
        lib.cfear_odometry_step_batch_wait.argtypes = [vp, i32]
        lib.cfear_odometry_step_batch_dev.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, vp]
        lib.cfear_odometry_step_batch_dev_submit.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp]
        lib.cfear_stream_wait_ticket.argtypes = [vp, i32]
        lib.cfear_join.argtypes = [vp]
        lib.cfear_stage_timing.argtypes = [vp, i32, vp]
        lib.cfear_last_counts.argtypes = [vp, i32, vp, vp, vp]
        lib.cfear_cfar_filter.argtypes = [vp, vp, i32, C.POINTER(CfarParams), vp, i32, vp]
        lib.cfear_seq_create.argtypes = [vp, i32, i32, i32, C.POINTER(SeqParams), C.POINTER(C.c_void_p)]
        lib.cfear_seq_destroy.argtypes = [vp]
        lib.cfear_seq_step.argtypes = [vp, vp]
        lib.cfear_seq_step_dev.argtypes = [vp, vp]
        lib.cfear_seq_read.argtypes = [vp, i32, i32, vp, vp, vp]
        _lib = lib
    return _lib


def default_config(**kw) -> Config:
    cfg = Config()
    load().cfear_default_config(C.byref(cfg))
    for k, v in kw.items():
        if k == "cost" and isinstance(v, str):
            v = COST[v]
        if k == "loss" and isinstance(v, str):
            v = LOSS[v]
        if k == "solver_mode" and isinstance(v, str):
            v = SOLVER[v]
        if not hasattr(cfg, k):
            raise KeyError(k)
        setattr(cfg, k, v)
    return cfg


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def bind_to_device_numa(device: int) -> dict:
    """Pin the calling process to the CPU cores of the NUMA node the GPU hangs off, so that page-locked host buffers
    allocated afterwards (first touch) sit next to the GPU's PCIe root port.  Plumbing for multi-GPU hosts where eight
    ranks stream images host->device at once; a no-op (returns the reason) when the topology cannot be read."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return {"numa_node": None, "reason": "single node"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"numa_node": node, "reason": "no allowed cpu on that node"}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"numa_node": None, "reason": repr(e)[:80]}


def pinned_array(shape, dtype):
    """numpy array over page-locked host memory.  The memory is released (cfear_free_pinned) when the last numpy view of
    it is garbage collected: the finalizer hangs on the ctypes buffer every view keeps alive through `.base`."""
    lib = load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.cfear_alloc_pinned(max(n, 1))
    if not p:
        raise CfearError(lib.cfear_last_error().decode())
    buf = (C.c_char * max(n, 1)).from_address(p)
    weakref.finalize(buf, lib.cfear_free_pinned, p).atexit = False      # at interpreter exit the driver reclaims it
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class Context:
    """Owns one cfear_ctx.  Thin, 1:1 with the C entry points."""

    def __init__(self, cfg: Config | None = None, **kw):
        self.lib = load()
        self.cfg = cfg if cfg is not None else default_config(**kw)
        h = C.c_void_p()
        rc = self.lib.cfear_create(C.byref(self.cfg), C.byref(h))
        if rc != 0:
            raise CfearError(f"cfear_create failed ({rc}): {self.lib.cfear_last_error().decode()}")
        self.h = h
        self.A, self.R, self.k = self.cfg.azimuths, self.cfg.range_bins, self.cfg.k_strongest
        self.cap_pts = self.A * self.k
        self.max_cells = self.cfg.max_cells if self.cfg.max_cells > 0 else self.cap_pts

    def update_config(self, **kw):
        for k, v in kw.items():
            if k == "cost" and isinstance(v, str):
                v = COST[v]
            if k == "loss" and isinstance(v, str):
                v = LOSS[v]
            if k == "solver_mode" and isinstance(v, str):
                v = SOLVER[v]
            setattr(self.cfg, k, v)
        self._ck(self.lib.cfear_update_config(self.h, C.byref(self.cfg)), "cfear_update_config")

    def close(self):
        if getattr(self, "h", None):
            self.lib.cfear_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise CfearError(f"{what} failed ({rc}): {self.lib.cfear_last_error().decode()}")

    @property
    def launches(self) -> int:
        return int(self.lib.cfear_launch_count(self.h))

    # ---- filter ----
    def kstrongest(self, polar):
        polar = np.ascontiguousarray(polar, dtype=np.uint8).reshape(-1, self.A, self.R)
        n = polar.shape[0]
        idx = np.empty((n, self.A, self.k), np.int32)
        cnt = np.empty((n, self.A), np.int32)
        self._ck(self.lib.cfear_kstrongest(self.h, _ptr(polar), n, _ptr(idx), _ptr(cnt)), "cfear_kstrongest")
        return idx, cnt

    def filter(self, polar, peaks=False):
        polar = np.ascontiguousarray(polar, dtype=np.uint8).reshape(-1, self.A, self.R)
        n = polar.shape[0]
        idx = np.empty((n, self.A, self.k), np.int32)
        cnt = np.empty((n, self.A), np.int32)
        cloud = np.zeros((n, self.cap_pts, 4), np.float32)
        npts = np.zeros(n, np.int32)
        pk = np.zeros((n, self.cap_pts, 4), np.float32) if peaks else None
        npk = np.zeros(n, np.int32) if peaks else None
        self._ck(self.lib.cfear_filter(self.h, _ptr(polar), n, _ptr(idx), _ptr(cnt), _ptr(cloud), _ptr(npts),
                                       _ptr(pk), _ptr(npk)), "cfear_filter")
        out = dict(idx=idx, cnt=cnt, clouds=[cloud[i, :npts[i]].copy() for i in range(n)], npts=npts)
        if peaks:
            out["peaks"] = [pk[i, :npk[i]].copy() for i in range(n)]
        return out

    def cfar_filter(self, polar, window_size=10, nb_guard_cells=20, false_alarm_rate=0.01, max_distance=400.0, capacity=None):
        polar = np.ascontiguousarray(polar, dtype=np.uint8).reshape(-1, self.A, self.R)
        n = polar.shape[0]
        cap = capacity if capacity is not None else 8 * self.cap_pts
        cp = CfarParams(window_size, nb_guard_cells, false_alarm_rate, max_distance)
        cloud = np.zeros((n, cap, 4), np.float32); npts = np.zeros(n, np.int32)
        self._ck(self.lib.cfear_cfar_filter(self.h, _ptr(polar), n, C.byref(cp), _ptr(cloud), cap, _ptr(npts)), "cfear_cfar_filter")
        return [cloud[i, :npts[i]].copy() for i in range(n)]

    def compensate(self, cloud, mot, ccw=False):
        out = np.ascontiguousarray(cloud, dtype=np.float32).copy()
        m = np.ascontiguousarray(mot, dtype=np.float64)
        self._ck(self.lib.cfear_compensate(self.h, _ptr(out), out.shape[0], _ptr(m), int(bool(ccw))), "cfear_compensate")
        return out

    # ---- surface points ----
    def surface_points(self, cloud, slot):
        cloud = np.ascontiguousarray(cloud, dtype=np.float32)
        nc = C.c_int32(0)
        self._ck(self.lib.cfear_surface_points(self.h, _ptr(cloud), cloud.shape[0], int(slot), C.byref(nc)),
                 "cfear_surface_points")
        return nc.value

    def scans_to_cells_batch(self, polar, mot, slots):
        polar = np.ascontiguousarray(polar, dtype=np.uint8).reshape(-1, self.A, self.R)
        n = polar.shape[0]
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        m = None if mot is None else np.ascontiguousarray(mot, dtype=np.float64)
        npts = np.zeros(n, np.int32); nc = np.zeros(n, np.int32)
        self._ck(self.lib.cfear_scans_to_cells_batch(self.h, n, _ptr(polar), _ptr(m), _ptr(slots), _ptr(npts), _ptr(nc)),
                 "cfear_scans_to_cells_batch")
        return npts, nc

    def cells_count(self, slot):
        nc = C.c_int32(0)
        self._ck(self.lib.cfear_cells_count(self.h, int(slot), C.byref(nc)), "cfear_cells_count")
        return nc.value

    def cells_download(self, slot):
        n = self.cells_count(slot)
        out = np.zeros(max(n, 1), CELL_DTYPE)
        nc = C.c_int32(0)
        self._ck(self.lib.cfear_cells_download(self.h, int(slot), _ptr(out), out.shape[0], C.byref(nc)), "cfear_cells_download")
        out = out[:nc.value]
        return dict(mean=out["mean"].copy(), normal=out["normal"].copy(), cov=out["cov"].reshape(-1, 2, 2).copy(),
                    planarity=out["planarity"].copy(), nsamples=out["nsamples"].copy(),
                    avg_intensity=out["avg_intensity"].copy())

    def cells_upload(self, slot, cells: dict):
        n = cells["mean"].shape[0]
        a = np.zeros(max(n, 1), CELL_DTYPE)
        if n:
            a["mean"][:n] = cells["mean"]; a["normal"][:n] = cells["normal"]
            a["cov"][:n] = np.asarray(cells["cov"]).reshape(n, 4)
            a["planarity"][:n] = cells["planarity"]; a["nsamples"][:n] = cells["nsamples"]
            a["avg_intensity"][:n] = cells.get("avg_intensity", np.zeros(n))
        self._ck(self.lib.cfear_cells_upload(self.h, int(slot), _ptr(a), n), "cfear_cells_upload")

    def nearest(self, slot, queries, radius):
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 2)
        out = np.full(q.shape[0], -2, np.int32)
        self._ck(self.lib.cfear_nearest(self.h, int(slot), _ptr(q), q.shape[0], float(radius), _ptr(out)), "cfear_nearest")
        return out

    # ---- registration ----
    def register_batch(self, slots, poses, want_assoc=False, want_sim=False, prior_sqrt_info=None):
        """cfear_register_batch[_ex].  Returns (poses, cov [n,6,6], stats, assoc) -- plus the similarity table when want_sim.
        prior_sqrt_info: (nprob, 3, 3) lower-triangular L of Register(..., soft_constraints=true)."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        nprob, ns = slots.shape
        p = np.ascontiguousarray(poses, dtype=np.float64).reshape(nprob, ns, 3).copy()
        cov = np.zeros((nprob, 36))
        st = np.zeros(nprob, STATS_DTYPE)
        assoc = np.full((nprob, ns - 1, self.max_cells), -1, np.int32) if (want_assoc or want_sim) else None
        sim = np.zeros((nprob, ns - 1, self.max_cells)) if want_sim else None
        L = None if prior_sqrt_info is None else np.ascontiguousarray(prior_sqrt_info, dtype=np.float64).reshape(nprob, 9)
        if sim is None and L is None:
            self._ck(self.lib.cfear_register_batch(self.h, nprob, _ptr(slots), ns, _ptr(p), _ptr(cov), _ptr(st), _ptr(assoc)),
                     "cfear_register_batch")
        else:
            self._ck(self.lib.cfear_register_batch_ex(self.h, nprob, _ptr(slots), ns, _ptr(p), _ptr(cov), _ptr(st), _ptr(assoc),
                                                      _ptr(sim), _ptr(L)), "cfear_register_batch_ex")
        if want_sim:
            return p, cov.reshape(nprob, 6, 6), st, assoc, sim
        return p, cov.reshape(nprob, 6, 6), st, assoc

    def get_cost_batch(self, slots, poses):
        """n_scan_normal_reg::GetCost for (nprob, nscans) slot tables and (nprob, nscans, 3) poses.
        Returns (cost [nprob], num_residuals [nprob], ok [nprob])."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        n, ns = slots.shape
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(n, ns, 3)
        cost = np.zeros(n); nres = np.zeros(n, np.int32); ok = np.zeros(n, np.int32)
        self._ck(self.lib.cfear_get_cost_batch(self.h, n, _ptr(slots), ns, _ptr(poses), _ptr(cost), _ptr(nres), _ptr(ok)),
                 "cfear_get_cost_batch")
        return cost, nres, ok

    def register(self, slots, poses):
        slots = np.ascontiguousarray(slots, dtype=np.int32).reshape(1, -1)
        p, cov, st, _ = self.register_batch(slots, np.asarray(poses)[None])
        return p[0], cov[0], st[0]

    # ---- whole path ----
    def odometry_step_batch(self, polar, mot, kf_slots, cur_slots, poses, out=None):
        """polar (n,A,R) u8 HOST; mot (n,3) or None; kf_slots (n,K); cur_slots (n); poses (n,K+1,3).
        `out`: optional dict of preallocated arrays (poses/cov/stats) reused across calls."""
        kf_slots = np.ascontiguousarray(kf_slots, dtype=np.int32)
        n, K = kf_slots.shape
        cur_slots = np.ascontiguousarray(cur_slots, dtype=np.int32)
        if polar.dtype != np.uint8 or not polar.flags["C_CONTIGUOUS"]:
            polar = np.ascontiguousarray(polar, dtype=np.uint8)
        if out is None:
            out = dict(poses=np.empty((n, K + 1, 3)), cov=np.zeros((n, 36)), stats=np.zeros(n, STATS_DTYPE),
                       npts=np.zeros(n, np.int32))
        np.copyto(out["poses"], np.asarray(poses, dtype=np.float64).reshape(n, K + 1, 3))
        m = None if mot is None else np.ascontiguousarray(mot, dtype=np.float64)
        self._ck(self.lib.cfear_odometry_step_batch(self.h, n, _ptr(polar), _ptr(m), _ptr(kf_slots), K, _ptr(cur_slots),
                                                    _ptr(out["poses"]), _ptr(out["cov"]), _ptr(out["stats"]),
                                                    _ptr(out.get("npts")), None), "cfear_odometry_step_batch")
        return out

    def odometry_step_batch_submit(self, polar, mot, kf_slots, cur_slots, out):
        """Asynchronous half of odometry_step_batch (cfear_odometry_step_batch_submit): every array must be C-contiguous of
        the right dtype already and stay alive and untouched until odometry_step_batch_wait(ticket).  `out` as in
        odometry_step_batch, with out["poses"] holding the input poses (last = guess).  Returns the ticket."""
        n, K = kf_slots.shape
        for a, dt in ((polar, np.uint8), (kf_slots, np.int32), (cur_slots, np.int32), (out["poses"], np.float64)):
            if a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                raise CfearError("odometry_step_batch_submit needs C-contiguous arrays of the ABI dtypes")
        if mot is not None and (mot.dtype != np.float64 or not mot.flags["C_CONTIGUOUS"]):
            raise CfearError("odometry_step_batch_submit needs C-contiguous arrays of the ABI dtypes")
        t = C.c_int32(-1)
        self._ck(self.lib.cfear_odometry_step_batch_submit(self.h, n, _ptr(polar), _ptr(mot), _ptr(kf_slots), K, _ptr(cur_slots),
                                                           _ptr(out["poses"]), _ptr(out.get("cov")), _ptr(out.get("stats")),
                                                           _ptr(out.get("npts")), C.byref(t)), "cfear_odometry_step_batch_submit")
        return int(t.value)

    def odometry_step_batch_wait(self, ticket):
        self._ck(self.lib.cfear_odometry_step_batch_wait(self.h, int(ticket)), "cfear_odometry_step_batch_wait")

    def last_counts(self, cur_slots):
        cur_slots = np.ascontiguousarray(cur_slots, dtype=np.int32)
        n = cur_slots.shape[0]
        npts = np.zeros(n, np.int32); nc = np.zeros(n, np.int32)
        self._ck(self.lib.cfear_last_counts(self.h, n, _ptr(cur_slots), _ptr(npts), _ptr(nc)), "cfear_last_counts")
        return npts, nc

    def stage_timing(self, enable=True):
        """Returns (nsteps, [ms_kstrongest, ms_surface, ms_register]) summed since the previous call."""
        ms = (C.c_float * 3)()
        rc = self.lib.cfear_stage_timing(self.h, int(bool(enable)), ms)
        if rc < 0:
            self._ck(rc, "cfear_stage_timing")
        return rc, [ms[0], ms[1], ms[2]]

    @property
    def stream_ptr(self) -> int:
        return int(self.lib.cfear_stream(self.h))

    def sync(self):
        self._ck(self.lib.cfear_sync(self.h), "cfear_sync")

    # ---- device-resident variant ----
    def dev_alloc(self, nbytes):
        p = self.lib.cfear_alloc_device(self.h, int(nbytes))
        if not p:
            raise CfearError(self.lib.cfear_last_error().decode())
        return p

    def dev_free(self, p):
        self.lib.cfear_free_device(self.h, p)

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._ck(self.lib.cfear_memcpy_h2d(self.h, dptr, _ptr(arr), arr.nbytes), "cfear_memcpy_h2d")

    def d2h(self, arr, dptr):
        self._ck(self.lib.cfear_memcpy_d2h(self.h, _ptr(arr), dptr, arr.nbytes), "cfear_memcpy_d2h")

    def odometry_step_batch_dev(self, n, d_polar, d_mot, d_kf_slots, K, d_cur_slots, d_poses, d_cov36, d_stats):
        self._ck(self.lib.cfear_odometry_step_batch_dev(self.h, n, d_polar, d_mot, d_kf_slots, K, d_cur_slots, d_poses,
                                                        d_cov36, d_stats), "cfear_odometry_step_batch_dev")

    def odometry_step_batch_dev_submit(self, n, d_polar, d_mot, d_kf_slots, K, d_cur_slots, d_poses, d_cov36, d_stats):
        """Overlapped device-resident step (cfear_odometry_step_batch_dev_submit); returns the ticket."""
        t = C.c_int32(-1)
        self._ck(self.lib.cfear_odometry_step_batch_dev_submit(self.h, n, d_polar, d_mot, d_kf_slots, K, d_cur_slots, d_poses,
                                                               d_cov36, d_stats, C.byref(t)), "cfear_odometry_step_batch_dev_submit")
        return int(t.value)

    def stream_wait_ticket(self, ticket):
        self._ck(self.lib.cfear_stream_wait_ticket(self.h, int(ticket)), "cfear_stream_wait_ticket")

    def join(self):
        self._ck(self.lib.cfear_join(self.h), "cfear_join")


class Sequences:
    """nseq independent sequences replayed in lock-step on the device (cfear_seq_*)."""

    def __init__(self, ctx: Context, nseq: int, max_steps: int, slot_base: int = 0, submap_scan_size: int = 3,
                 use_guess: bool = True, use_keyframe: bool = True, min_keyframe_dist: float = 1.5, min_keyframe_rot_deg: float = 5.0):
        self.ctx, self.nseq, self.max_steps = ctx, nseq, max_steps
        sp = SeqParams(submap_scan_size, int(use_guess), int(use_keyframe), 0, min_keyframe_dist, min_keyframe_rot_deg)
        h = C.c_void_p()
        ctx._ck(ctx.lib.cfear_seq_create(ctx.h, nseq, slot_base, max_steps, C.byref(sp), C.byref(h)), "cfear_seq_create")
        self.h = h

    def step(self, polar):
        """polar: host (nseq, A, R) uint8."""
        if polar.dtype != np.uint8 or not polar.flags["C_CONTIGUOUS"]:
            polar = np.ascontiguousarray(polar, dtype=np.uint8)
        self._keep = polar          # the copy is asynchronous: keep the buffer alive until the next call
        self.ctx._ck(self.ctx.lib.cfear_seq_step(self.h, _ptr(polar)), "cfear_seq_step")

    def step_dev(self, d_polar_ptr):
        self.ctx._ck(self.ctx.lib.cfear_seq_step_dev(self.h, d_polar_ptr), "cfear_seq_step_dev")

    def read(self, step_from, nsteps):
        poses = np.zeros((self.nseq, nsteps, 3)); kf = np.zeros((self.nseq, nsteps), np.int32)
        st = np.zeros((self.nseq, nsteps), STATS_DTYPE)
        self.ctx._ck(self.ctx.lib.cfear_seq_read(self.h, step_from, nsteps, _ptr(poses), _ptr(kf), _ptr(st)), "cfear_seq_read")
        return poses, kf, st

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.cfear_seq_destroy(self.h)
            self.h = None
