"""Generates tests/golden/tracker_golden.npz with the REAL OpenCV (cv2) as the source of truth for the OpenCV stages of
Tracker::track (Tracker.cc:198-202 CLAHE, :237-244 calcOpticalFlowPyrLK, :100-132 undistortPoints).

Run here (cv2 %s is in this image); the fixture is committed so that the GPU box / CI never needs /root/reference.
The reference repository itself has no golden vectors (SURVEY 4); OpenCV is its non-vendored dependency.
""" % "4.13.0"
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvio_b200  # noqa: E402,F401
from rvio_b200 import synth  # noqa: E402


def main():
    out = {}
    out["cv2_version"] = np.array(cv2.__version__)
    W, H = 320, 240
    cfg = synth.Config(width=W, height=H, fx=458.654 * W / 752, fy=457.296 * H / 480, cx=367.215 * W / 752, cy=248.375 * H / 480,
                       n_features=128)
    st = synth.Stream(cfg, 3, 99, t_static=0.0)
    # a second pair with large motion / low contrast so that tracks are lost
    a = st.frames[0].copy(); b = st.frames[2].copy()
    M = cv2.getRotationMatrix2D((W / 2, H / 2), 3.0, 1.03); M[:, 2] += (6.5, -4.25)
    c = cv2.warpAffine(a, M, (W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    c[:, :W // 4] = (c[:, :W // 4].astype(np.float32) * 0.15 + 100).astype(np.uint8)
    raws = [a, b, c]
    clahe = [cv2.createCLAHE(3.0, (5, 5)).apply(r) for r in raws]
    out["raw"] = np.stack(raws); out["clahe"] = np.stack(clahe)
    pyr = [clahe[0]]
    for _ in range(3):
        pyr.append(cv2.pyrDown(pyr[-1]))
    for l in range(1, 4):
        out[f"pyr{l}"] = pyr[l]
    out["scharr_dx"] = cv2.Scharr(clahe[0], cv2.CV_16S, 1, 0); out["scharr_dy"] = cv2.Scharr(clahe[0], cv2.CV_16S, 0, 1)
    pts = cv2.goodFeaturesToTrack(clahe[0], 100, 0.01, 8).reshape(-1, 2).astype(np.float32)
    r = np.random.default_rng(5)
    extra = np.stack([r.uniform(-4, W + 4, 24), r.uniform(-4, H + 4, 24)], 1).astype(np.float32)
    pts = np.ascontiguousarray(np.concatenate([pts, extra])[:cfg.n_features])
    out["pts"] = pts
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 1e-2)
    for name, nxt_img in (("b", clahe[1]), ("c", clahe[2])):
        nxt, stt, _ = cv2.calcOpticalFlowPyrLK(clahe[0], nxt_img, pts, None, winSize=(15, 15), maxLevel=3, criteria=crit, flags=0,
                                               minEigThreshold=1e-3)
        out[f"lk_{name}_px"] = nxt.reshape(-1, 2); out[f"lk_{name}_status"] = stt.reshape(-1)
        K = np.array([[cfg.fx, 0, cfg.cx], [0, cfg.fy, cfg.cy], [0, 0, 1]], np.float32)
        D = np.array([cfg.k1, cfg.k2, cfg.p1, cfg.p2], np.float32)
        out[f"un_{name}"] = cv2.undistortPoints(nxt.reshape(-1, 1, 2), K, D).reshape(-1, 2)
    out["K4"] = np.array([cfg.fx, cfg.fy, cfg.cx, cfg.cy], np.float32)
    out["D5"] = np.array([cfg.k1, cfg.k2, cfg.p1, cfg.p2, 0], np.float32)
    # colour conversion (Tracker.cc:183-196)
    rgb = r.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    out["bgr"] = rgb; out["bgr2gray"] = cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY); out["rgb2gray"] = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", "tracker_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; lost b/c:", int((out["lk_b_status"] == 0).sum()), int((out["lk_c_status"] == 0).sum()))


if __name__ == "__main__":
    main()
