"""Pins rows a1-a3 (k-strongest index sets, filtered cloud, peaks cloud) and the CA-CFAR filter to the REFERENCE SOURCE.

oracle/_ref/libcfear_ref.so is the reference's own radar_filters.cpp + cfar.cpp, compiled unmodified from
/root/reference against interface stubs (oracle/ref_stubs; recipe oracle/Makefile `ref`; built by __graft_entry__.build()
in the build container and shipped to the GPU box as a prebuilt file).  Calls go through oracle/ref_shim.cc exactly like
radarDriver::Process makes them (radar_driver.cpp:48-61).

  not gpu:  the oracle restatement (oracle/cfear_oracle.cc) == the reference, bit for bit
  gpu:      the CUDA path through the C ABI == the reference, bit for bit (no oracle in the loop)
"""
import numpy as np
import pytest

import helpers
from oracle import ref

pytestmark = pytest.mark.skipif(not (ref.available() or ref.build()), reason="oracle/_ref/libcfear_ref.so not built (needs /root/reference)")

FILTER_CASES = [(12, 60.0), (1, 60.0), (40, 0.0), (5, 255.0), (12, 65.7), (64, 128.0), (40, 60.0)]
CFAR_CASES = [dict(), dict(window_size=40, false_alarm_rate=0.001, nb_guard_cells=5),
              dict(window_size=3, nb_guard_cells=0, z_min=20.0), dict(window_size=10, nb_guard_cells=20, z_min=80.5, min_distance=1.0)]


def _images():
    synth = helpers.scan_images(3, 0)[0][0]
    wide = np.ascontiguousarray(np.pad(helpers.scan_images(4, 0)[0][0], ((0, 0), (0, 408)), mode="wrap"))   # Oxford width 3768
    return [("adversarial", helpers.adversarial_image(0)), ("synthetic", synth), ("synthetic3768", wide)]


@pytest.fixture(scope="module")
def images():
    return _images()


# ---- CPU: the oracle restatement against the reference source ------------------------------------------------------
@pytest.mark.parametrize("k,zmin", FILTER_CASES)
def test_oracle_filter_equals_reference_source(orc, images, k, zmin):
    for name, img in images:
        r = ref.kstrongest(img, zmin, k)
        oi, oc = orc.kstrongest(img, int(zmin), k)          # z_min: float -> int (ctor argument) -> uchar, radar_filters.cpp:198,212
        assert np.array_equal(oc, r["cnt"]) and np.array_equal(oi, r["idx"]), name
        cl = orc.cloud(img, oi, oc)
        assert cl.shape == r["cloud"].shape and np.array_equal(cl.view(np.uint32), r["cloud"].view(np.uint32)), name
        pi, pc = orc.peaks(img, oi, oc)
        assert np.array_equal(pc, r["pcnt"]) and np.array_equal(pi, r["pidx"]), name
        pk = orc.cloud(img, pi, pc)
        assert pk.shape == r["peaks"].shape and np.array_equal(pk.view(np.uint32), r["peaks"].view(np.uint32)), name


def test_oracle_filter_equals_reference_source_other_geometry(orc):
    rng = np.random.Generator(np.random.PCG64(11))
    for A, R, k in [(37, 1001, 12), (5, 17, 3), (1, 7, 12), (400, 64, 64)]:
        img = rng.integers(0, 256, (A, R), dtype=np.uint8)
        r = ref.kstrongest(img, 60.0, k, min_distance=0.5, range_res=0.0595238)
        oi, oc = orc.kstrongest(img, 60, k)
        assert np.array_equal(oc, r["cnt"]) and np.array_equal(oi, r["idx"])
        cl = orc.cloud(img, oi, oc, min_distance=0.5, range_res=0.0595238)
        assert np.array_equal(cl.view(np.uint32), r["cloud"].view(np.uint32))
        pi, pc = orc.peaks(img, oi, oc)
        assert np.array_equal(pc, r["pcnt"]) and np.array_equal(pi, r["pidx"])


@pytest.mark.parametrize("pars", CFAR_CASES)
def test_oracle_cfar_equals_reference_source(orc, images, pars):
    for name, img in images:
        a, b = ref.cfar(img, **pars), orc.cfar(img, **pars)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), name


# ---- GPU: the CUDA path against the reference source ---------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("k,zmin", FILTER_CASES)
def test_cuda_filter_equals_reference_source(images, k, zmin):
    from cfear_radarodometry_code_public_b200 import capi
    for name, img in images:
        A, R = img.shape
        c = capi.Context(max_batch=1, azimuths=A, range_bins=R, k_strongest=k, z_min=zmin, max_cellsets=2, max_cells=4096)
        out = c.filter(img[None], peaks=True)
        r = ref.kstrongest(img, zmin, k)
        assert np.array_equal(out["cnt"][0], r["cnt"]) and np.array_equal(out["idx"][0], r["idx"]), name
        assert out["npts"][0] == r["cloud"].shape[0], name
        assert np.array_equal(out["clouds"][0].view(np.uint32), r["cloud"].view(np.uint32)), name
        assert out["peaks"][0].shape == r["peaks"].shape, name
        assert np.array_equal(out["peaks"][0].view(np.uint32), r["peaks"].view(np.uint32)), name
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pars", CFAR_CASES)
def test_cuda_cfar_equals_reference_source(images, pars):
    from cfear_radarodometry_code_public_b200 import capi
    for name, img in images:
        A, R = img.shape
        kw = {k: pars[k] for k in ("z_min", "min_distance") if k in pars}
        c = capi.Context(max_batch=1, azimuths=A, range_bins=R, max_cellsets=2, max_cells=4096, **kw)
        cp = {k: v for k, v in pars.items() if k in ("window_size", "false_alarm_rate", "nb_guard_cells")}
        g = c.cfar_filter(img[None], **cp)[0]
        a = ref.cfar(img, **pars)
        assert g.shape == a.shape, name
        assert np.array_equal(g.view(np.uint32), a.view(np.uint32)), name
        c.close()
