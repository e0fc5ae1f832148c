"""CPU: the oracle's restatement of the OpenCV stages is pinned bit-exactly
 (a) against the committed golden fixture produced by the real OpenCV (oracle/make_golden.py), and
 (b) against cv2 live when it is importable (it is in this image), on more sizes / border cases."""
import os

import numpy as np
import pytest

import rvio_b200  # noqa: F401
from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tracker_golden.npz")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_golden_clahe(gold):
    for raw, ref in zip(gold["raw"], gold["clahe"]):
        assert np.array_equal(orc.clahe(np.ascontiguousarray(raw)), ref)


def test_golden_pyramid(gold):
    lv = np.ascontiguousarray(gold["clahe"][0])
    for l in range(1, 4):
        lv = orc.pyr_down(lv)
        assert np.array_equal(lv, gold[f"pyr{l}"])


def test_golden_scharr(gold):
    img = np.ascontiguousarray(gold["clahe"][0])
    h, w = img.shape
    d = np.empty((h, w, 2), np.int16)
    orc.lib().orc_scharr(img, w, h, w, d)
    assert np.array_equal(d[..., 0], gold["scharr_dx"]) and np.array_equal(d[..., 1], gold["scharr_dy"])


@pytest.mark.parametrize("name,idx", [("b", 1), ("c", 2)])
def test_golden_lk_and_undistort(gold, name, idx):
    prev = np.ascontiguousarray(gold["clahe"][0]); nxt = np.ascontiguousarray(gold["clahe"][idx])
    px, st = orc.lk(prev, nxt, gold["pts"])
    assert np.array_equal(st, gold[f"lk_{name}_status"])
    assert np.array_equal(_bits(px), _bits(gold[f"lk_{name}_px"]))            # bit-exact float32, lost points included
    un = np.empty_like(px)
    orc.lib().orc_undistort(np.ascontiguousarray(px), len(px), np.ascontiguousarray(gold["K4"]), np.ascontiguousarray(gold["D5"]), un)
    assert np.array_equal(_bits(un), _bits(gold[f"un_{name}"]))


def test_live_cv2_various_sizes():
    cv2 = pytest.importorskip("cv2")
    for (w, h, seed) in [(752, 480, 1), (641, 479, 2), (100, 75, 3), (37, 52, 4)]:
        r = np.random.default_rng(seed)
        a = cv2.GaussianBlur(r.standard_normal((h + 32, w + 32)).astype(np.float32), (0, 0), 2.0)
        a = ((a - a.min()) / (a.max() - a.min()) * 255).astype(np.uint8)
        a[:, : w // 3] = (a[:, : w // 3] * 0.3).astype(np.uint8)
        im0 = np.ascontiguousarray(a[5:5 + h, 5:5 + w])
        M = cv2.getRotationMatrix2D((w / 2, h / 2), 0.9, 1.006); M[:, 2] += (2.3, -1.6)
        im1 = cv2.warpAffine(im0, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        assert np.array_equal(orc.clahe(im0), cv2.createCLAHE(3.0, (5, 5)).apply(im0)), (w, h)
        assert np.array_equal(orc.pyr_down(im0), cv2.pyrDown(im0)), (w, h)
        e0, e1 = orc.clahe(im0), orc.clahe(im1)
        pts = np.stack([r.uniform(-6, w + 6, 150), r.uniform(-6, h + 6, 150)], 1).astype(np.float32)
        nxt, st, _ = cv2.calcOpticalFlowPyrLK(e0, e1, pts, None, winSize=(15, 15), maxLevel=3,
                                              criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 1e-2), flags=0, minEigThreshold=1e-3)
        px, s2 = orc.lk(e0, e1, pts)
        assert np.array_equal(s2, st.ravel()), (w, h)
        assert np.array_equal(_bits(px), _bits(nxt.reshape(-1, 2))), (w, h)


def test_gray_conversion_formula(gold):
    # Tracker.cc:183-196 -> cvtColor: OpenCV 4.x 15-bit fixed point; the CUDA k_gray kernel uses the same integers
    bgr = gold["bgr"].astype(np.int64)
    f = lambda b, g, r: ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)
    assert np.array_equal(f(bgr[..., 0], bgr[..., 1], bgr[..., 2]), gold["bgr2gray"])
    assert np.array_equal(f(bgr[..., 2], bgr[..., 1], bgr[..., 0]), gold["rgb2gray"])


def test_live_cv2_fisheye_undistort():
    """cv::fisheye::undistortPoints (Tracker.cc:119, Camera.Fisheye: 1) restated in oracle/img.c, bit for bit against cv2:
    points all over the image (incl. far outside: the non-converged / clipped branch) and two coefficient sets."""
    cv2 = pytest.importorskip("cv2")
    r = np.random.default_rng(11)
    for D in ([-0.0127, 0.0154, -0.0201, 0.0072], [0.21, -0.35, 0.6, -0.4]):
        K4 = np.array([458.654, 457.296, 367.215, 248.375], np.float32)
        D4 = np.array(D, np.float32)
        px = np.concatenate([np.stack([r.uniform(-50, 800, 3000), r.uniform(-50, 530, 3000)], 1),
                             np.stack([r.uniform(-4000, 5000, 500), r.uniform(-4000, 5000, 500)], 1),
                             [[367.215, 248.375], [367.2150001, 248.375]]]).astype(np.float32)
        K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
        want = cv2.fisheye.undistortPoints(px.reshape(-1, 1, 2), K, D4.reshape(4, 1)).reshape(-1, 2)
        got = np.empty_like(px)
        orc.lib().orc_undistort_fisheye(px, len(px), K4, D4, got)
        bad = (_bits(got) != _bits(want)).any(1)
        assert not bad.any(), (D, int(bad.sum()), px[bad][:4], got[bad][:4], want[bad][:4])

