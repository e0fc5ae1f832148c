"""GPU parity: the device detector (FeatureDetector::DetectWithSubPix, detector.cu, through rvio_tracker_detect) against
the CPU restatement oracle/detector.c -- bit for bit (corner order, integer selection, float32 sub-pixel coordinates) --
and, through the oracle's own pin, against cv2.  Then the fused pipeline with the detector on the device against the
CPU oracle fed with the very same corners."""
import numpy as np
import pytest

import rvio_b200  # noqa: F401
from rvio_b200 import synth, host, capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("idx,min_dist", [(1, 15.0), (0, 15.0), (2, 15.0), (1, 8.0)])
def test_device_detector_matches_restatement(idx, min_dist):
    cfg = synth.baseline_config(idx)
    cfg.min_dist = min_dist
    st = synth.Stream(cfg, 5, 20260930 + idx, t_static=0.1)
    trk = host.Tracker(cfg, 0, detector="device")
    consumed = 0
    checked = 0
    for i in range(5):
        imu, consumed = st.imu_for_frame(i, consumed)
        if len(imu) < 2:
            continue
        im = np.ascontiguousarray(st.frames[i])
        imu_c = np.ascontiguousarray(imu, np.float64).reshape(-1, 8)
        rc = capi.check(trk.L.rvio_tracker_track(trk.h, im.reshape(-1), im.shape[1], im.shape[0], im.strides[0], 1,
                                                 imu_c.ctypes.data, len(imu_c)))
        eq = trk.equalized_image()
        assert np.array_equal(eq, orc.clahe(im))
        for s in (1, 2):
            got = trk.detect(s)
            want = orc.detect_restated(eq, cfg.n_features, s, cfg)
            assert len(got) == len(want), (i, s, len(got), len(want))
            assert np.array_equal(_bits(got), _bits(want)), (i, s, int((_bits(got) != _bits(want)).any(1).sum()))
            checked += 1
        # finish the frame the ordinary way so that the next one tracks
        if rc == capi.FIRST_IMAGE:
            pts = trk.detect(1)
            capi.check(trk.L.rvio_tracker_seed(trk.h, pts, len(pts)))
        capi.check(trk.L.rvio_tracker_commit(trk.h))
    assert checked >= 6


def test_device_detector_close_to_cv2():
    cfg = synth.baseline_config(1)
    st = synth.Stream(cfg, 3, 20261001, t_static=0.1)
    trk = host.Tracker(cfg, 0, detector="device")
    imu, _ = st.imu_for_frame(1, 0)
    im = np.ascontiguousarray(st.frames[1])
    imu_c = np.ascontiguousarray(imu, np.float64).reshape(-1, 8)
    capi.check(trk.L.rvio_tracker_track(trk.h, im.reshape(-1), im.shape[1], im.shape[0], im.strides[0], 1, imu_c.ctypes.data, len(imu_c)))
    eq = trk.equalized_image()
    for s in (1, 2):
        got = trk.detect(s)
        want = orc.detect_with_subpix(eq, cfg.n_features, s, cfg)              # real OpenCV
        assert len(got) == len(want)
        d = np.linalg.norm(got - want, axis=1)
        assert (d < 1e-4).mean() >= 0.98 and d.max() < 5e-2, (s, float(d.max()))


def test_vio_with_device_detector_matches_oracle():
    """Whole Tracker::track on the device: the fused pipeline run with n_cand = -1 (device detector) must produce the same
    poses as the CPU oracle pipeline whose detector is the restatement (same corners by the test above)."""
    cfg = synth.baseline_config(1)
    n = 60
    st = synth.Stream(cfg, n, 20261002, t_static=1.0)
    vio = host.Vio(cfg, 0)
    det = lambda img, k, s: orc.detect_restated(img, k, s, cfg)
    o = orc.VioOracle(cfg, det) if "detector" in orc.VioOracle.__init__.__code__.co_varnames else None
    if o is None:
        pytest.skip("oracle pipeline has no pluggable detector")
    try:
        consumed = 0
        worst = 0.0
        poses = 0
        for i in range(n):
            imu, consumed = st.imu_for_frame(i, consumed)
            if len(imu) < 2:
                continue
            pg = vio.step(st.frames[i], imu, device_detector=True)
            po = o.step(st.frames[i], imu)
            assert (pg is None) == (po is None), i
            if pg is not None:
                poses += 1
                worst = max(worst, float(np.abs(pg[:3] - po[:3]).max()))
        assert poses >= 25 and worst < 1e-7, (poses, worst)
    finally:
        pass
