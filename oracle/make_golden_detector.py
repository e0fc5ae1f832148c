"""Generates tests/golden/detector_golden.npz with the REAL OpenCV (cv2 4.13.0) as the source of truth for the detector the
reference calls (FeatureDetector.cc:55-75: cv::goodFeaturesToTrack + cv::cornerSubPix), plus cv::fisheye::undistortPoints
(Tracker.cc:119).  The images are the equalised frames already stored in tests/golden/tracker_golden.npz.
Run here; the fixture is committed so that nothing at test time depends on how cv2 was built on that machine."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "tracker_golden.npz"))
    out = {"cv2_version": np.array(cv2.__version__)}
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 1e-2)
    for k, img in enumerate(g["clahe"]):
        img = np.ascontiguousarray(img)
        if k == 0:
            out["eig0"] = cv2.cornerMinEigenVal(img, 3, ksize=3)
        for s, md in ((1, 15.0), (2, 15.0), (1, 8.0)):
            c = cv2.goodFeaturesToTrack(img, 128, float(np.float32(0.01)), s * md)
            c = np.zeros((0, 1, 2), np.float32) if c is None else np.ascontiguousarray(c, np.float32)
            out[f"gftt{k}_s{s}_d{int(md)}"] = c.reshape(-1, 2).copy()
            if len(c):
                hw = int(np.floor(.5 * md))
                cv2.cornerSubPix(img, c, (hw, hw), (-1, -1), crit)
            out[f"subpix{k}_s{s}_d{int(md)}"] = c.reshape(-1, 2)
    r = np.random.default_rng(5)
    px = np.stack([r.uniform(-40, 360, 400), r.uniform(-40, 280, 400)], 1).astype(np.float32)
    K4 = np.array([195.17, 228.65, 156.26, 124.19], np.float32)
    D4 = np.array([-0.0127, 0.0154, -0.0201, 0.0072], np.float32)
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
    out["fisheye_px"] = px; out["fisheye_K4"] = K4; out["fisheye_D4"] = D4
    out["fisheye_un"] = cv2.fisheye.undistortPoints(px.reshape(-1, 1, 2), K, D4.reshape(4, 1)).reshape(-1, 2)
    path = os.path.join(ROOT, "tests", "golden", "detector_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
