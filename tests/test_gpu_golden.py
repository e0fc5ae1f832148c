"""GPU: CUDA tracker stages against the committed golden fixture produced by the real OpenCV (cv2 4.13.0):
CLAHE, pyramid, pyramidal LK (status + float32 coordinates bit-exact, lost points included), undistortPoints, cvtColor."""
import ctypes as C
import os

import numpy as np
import pytest

import rvio_b200  # noqa: F401
from rvio_b200 import synth, host, capi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "tracker_golden.npz")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _cfg(w, h, eq, F=128):
    K = np.load(GOLD)["K4"]
    return synth.Config(width=w, height=h, fx=float(K[0]), fy=float(K[1]), cx=float(K[2]), cy=float(K[3]), n_features=F,
                        enable_equalizer=eq)


def _imu():
    imu = np.zeros((10, 8)); imu[:, 7] = 0.005; imu[:, 5] = 9.8
    return imu


def test_clahe_and_pyramid_vs_opencv():
    g = np.load(GOLD)
    H, W = g["raw"][0].shape
    for k in range(3):
        t = host.Tracker(_cfg(W, H, 1), 0)
        rc = t.track(g["raw"][k], _imu(), detections=np.zeros((0, 2), np.float32))
        assert rc == capi.FIRST_IMAGE
        # nothing seeded -> still "first image"; the equalised image is the current pyramid level 0
        assert np.array_equal(t.equalized_image(), g["clahe"][k])
        if k == 0:
            for l in range(1, 4):
                assert np.array_equal(t.pyramid(0, l), g[f"pyr{l}"])
        t.close()


@pytest.mark.parametrize("name,idx", [("b", 1), ("c", 2)])
def test_lk_undistort_vs_opencv(name, idx):
    g = np.load(GOLD)
    H, W = g["clahe"][0].shape
    t = host.Tracker(_cfg(W, H, 0), 0)
    assert t.track(g["clahe"][0], _imu(), detections=g["pts"]) == capi.FIRST_IMAGE
    assert t.track(g["clahe"][idx], _imu(), detections=np.zeros((0, 2), np.float32)) == capi.OK
    d = t.debug()
    assert d["n"] == len(g["pts"])
    assert np.array_equal(d["status"], g[f"lk_{name}_status"])
    assert np.array_equal(_bits(d["lk"]), _bits(g[f"lk_{name}_px"]))
    assert np.array_equal(_bits(d["un"]), _bits(g[f"un_{name}"]))
    t.close()


@pytest.mark.parametrize("is_rgb,key", [(0, "bgr2gray"), (1, "rgb2gray")])
def test_color_conversion_vs_opencv(is_rgb, key):
    g = np.load(GOLD)
    img = np.ascontiguousarray(g["bgr"])
    H, W, _ = img.shape
    cfg = _cfg(W, H, 0, F=16)
    L = capi.lib()
    tc = capi.tracker_cfg(cfg); tc.is_rgb = is_rgb
    h = C.c_void_p()
    capi.check(L.rvio_tracker_create(C.byref(tc), 0, C.byref(h)))
    imu = _imu()
    rc = capi.check(L.rvio_tracker_track(h, img.reshape(-1), W, H, img.strides[0], 3, imu.ctypes.data, len(imu)))
    assert rc == capi.FIRST_IMAGE
    out = np.empty((H, W), np.uint8)
    capi.check(L.rvio_tracker_get_image(h, out, W))
    assert np.array_equal(out, g[key])
    L.rvio_tracker_destroy(h)


UGOLD = os.path.join(os.path.dirname(__file__), "golden", "update_golden.npz")


def test_update_golden_vectors():
    """Updater::update on the device against the frozen vectors of oracle/make_golden_update.py (reference rule; shapes of
    configs 1, 2, 3, 5; includes the frames where the reference's first-small-row cut discards rows and where only the
    replay of its Givens sweep can decide).  float64: x 1e-9, P 1e-9 relative, kept-row count and per-feature status exact."""
    g = np.load(UGOLD)
    seen_cut = seen_sweep = 0
    for name in g["names"]:
        cfg = synth.baseline_config(int(g[f"{name}/cfg"]))
        upd = host.Updater(cfg)
        x, P = g[f"{name}/x"], g[f"{name}/P"]
        xg, Pg = upd.update(x, P, g[f"{name}/types"], (g[f"{name}/off"], g[f"{name}/xy"]))
        gi = upd.info
        gd = upd.debug(len(g[f"{name}/types"]))
        assert np.array_equal(gd["status"], g[f"{name}/status"]), name
        ok = g[f"{name}/status"] == 0
        np.testing.assert_allclose(gd["gamma"][ok], g[f"{name}/gamma"][ok], rtol=1e-8, err_msg=name)
        assert gi.n_good == int(g[f"{name}/n_good"]) and gi.rows_stacked == int(g[f"{name}/rows"]), name
        if gi.rank_flags & 16:
            assert gi.rank <= int(g[f"{name}/rank"]) == int(g[f"{name}/rank_full"]), name
        else:
            assert gi.rank == int(g[f"{name}/rank"]), (name, gi.rank, int(g[f"{name}/rank"]), gi.rank_flags)
        cut = int(g[f"{name}/rank"]) < int(g[f"{name}/rank_full"])
        assert bool(gi.rank_flags & 1) == cut, (name, gi.rank_flags)
        seen_cut += int(cut); seen_sweep += int(bool(gi.rank_flags & 2))
        np.testing.assert_allclose(xg, g[f"{name}/x_out"], rtol=0, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(Pg, g[f"{name}/P_out"], rtol=0, atol=1e-9 * np.abs(g[f"{name}/P_out"]).max(), err_msg=name)
        upd.close()
    assert seen_cut >= 2
