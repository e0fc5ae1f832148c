#!/usr/bin/env python
"""Runs the device pipeline over an EuRoC ASL directory and writes stamped_pose_ests.dat / time_cost.dat
(what `roslaunch rvio euroc.launch` + `rosbag play` produce with the reference, System.cc:369-380).

    python tools/run_asl.py <asl_dir> <out_dir> [--config config.yaml] [--device 0]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rvio_b200  # noqa: E402,F401
from rvio_b200 import synth, host, io_formats  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("asl_dir"); ap.add_argument("out_dir")
    ap.add_argument("--config", default=None, help="R-VIO yaml (keys of config/rvio_euroc.yaml); default: EuRoC values")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--detector", default="device", choices=["device", "host"],
                    help="device: FeatureDetector::DetectWithSubPix on the GPU inside the step; host: cv2 on the equalised frame")
    args = ap.parse_args()
    cfg = synth.Config.from_yaml(args.config) if args.config else synth.Config()
    host_detect = None
    if args.detector == "host":                 # the caller's own detector: real OpenCV, as FeatureDetector.cc:55-75 calls it
        import math
        import cv2
        clahe = cv2.createCLAHE(3.0, (5, 5))
        q, md = float(np.float32(cfg.qual_lvl)), float(np.float32(cfg.min_dist))
        hw = int(math.floor(.5 * md))

        def host_detect(eq, s):
            c = cv2.goodFeaturesToTrack(eq, cfg.n_features, q, s * md)
            if c is None or len(c) == 0:
                return np.zeros((0, 2), np.float32)
            c = np.ascontiguousarray(c.reshape(-1, 1, 2), np.float32)
            cv2.cornerSubPix(eq, c, (hw, hw), (-1, -1), (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 1e-2))
            return c.reshape(-1, 2)

    vio = host.Vio(cfg, args.device)
    out = io_formats.PoseWriter(args.out_dir)
    k = 0
    started = False
    for t, im, imu in io_formats.EurocAslReader(args.asl_dir, cfg.time_offset):
        t0 = time.perf_counter()
        cand = None
        if host_detect is not None:
            eq = clahe.apply(im) if cfg.enable_equalizer else im     # the image the reference's detector sees (Tracker.cc:198-207)
            cand = host_detect(eq, 2 if started else 1)
        t1 = time.perf_counter()
        pose = vio.step(im, imu, cand, device_detector=args.detector == "device")
        t2 = time.perf_counter()
        if pose is not None:
            started = True
            k += 1
            out.write(t, pose, k, 1e3 * (t1 - t0), 1e3 * (t2 - t1))   # host detector ms, fused device step ms
    out.close()
    print(f"{k} poses -> {args.out_dir}/stamped_pose_ests.dat")


if __name__ == "__main__":
    main()
