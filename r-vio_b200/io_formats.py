"""Data formats on either side of the hot path.

Input side: the EuRoC MAV "ASL" layout the reference is run on (README: V1_01_easy, converted from the rosbag the ROS node
subscribes to, rvio_mono.cc:60-113): `mav0/cam0/data.csv` + `mav0/cam0/data/<t>.png`, `mav0/imu0/data.csv`.  The reader
replays InputBuffer semantics (InputBuffer.cc:53-81, rvio_mono.cc:92-110): every image is paired with all not-yet-consumed
IMU rows whose stamp is <= t_image + Camera.nTimeOffset, each row = [w(3), a(3), t, dt] with dt = stamp difference to the
previous IMU message (0 for the very first), exactly the (n, 8) array the C ABI takes.

Output side: `stamped_pose_ests.dat` ("t px py pz qx qy qz qw", System.cc:371-373) and `time_cost.dat`
("k tracker_ms filter_ms", System.cc:376-378), numbers printed like std::setprecision(19) does (%.19g).
"""
import csv
import os

import numpy as np


def _g19(v) -> str:
    return "%.19g" % float(v)


class PoseWriter:
    """stamped_pose_ests.dat / time_cost.dat writers (System.cc:84-88,369-380)."""

    def __init__(self, directory: str):
        os.makedirs(directory, exist_ok=True)
        self.fp = open(os.path.join(directory, "stamped_pose_ests.dat"), "w")
        self.ft = open(os.path.join(directory, "time_cost.dat"), "w")

    def write(self, stamp: float, pose7, n_image_after_init: int, tracker_ms: float, filter_ms: float):
        p = np.asarray(pose7, np.float64)           # [pGk(3), qkG(4)]  (rvio_vio_step pose layout)
        self.fp.write(" ".join([_g19(stamp)] + [_g19(v) for v in p]) + "\n")
        self.fp.flush()
        self.ft.write(f"{int(n_image_after_init)} {_g19(tracker_ms)} {_g19(filter_ms)}\n")
        self.ft.flush()

    def close(self):
        self.fp.close(); self.ft.close()


def read_pose_file(path: str) -> np.ndarray:
    """(k, 8) array [t, px, py, pz, qx, qy, qz, qw] from a stamped_pose_ests.dat."""
    rows = [[float(v) for v in line.split()] for line in open(path) if line.strip()]
    return np.array(rows, np.float64).reshape(-1, 8)


def write_asl(directory: str, frame_t, frames, imu_rows, imu_rate_hint: float = 200.0):
    """Writes a stream in the EuRoC ASL layout.  frame_t: seconds; frames: list of (H, W) uint8; imu_rows: (n, >=7)
    [w(3), a(3), t, ...].  Stamps are integer nanoseconds (the dataset's convention); images are PNG (cv2)."""
    import cv2
    cam = os.path.join(directory, "mav0", "cam0"); imu = os.path.join(directory, "mav0", "imu0")
    os.makedirs(os.path.join(cam, "data"), exist_ok=True); os.makedirs(imu, exist_ok=True)
    with open(os.path.join(cam, "data.csv"), "w") as f:
        f.write("#timestamp [ns],filename\n")
        for t, im in zip(frame_t, frames):
            ns = int(round(float(t) * 1e9))
            name = f"{ns}.png"
            cv2.imwrite(os.path.join(cam, "data", name), im)
            f.write(f"{ns},{name}\n")
    with open(os.path.join(imu, "data.csv"), "w") as f:
        f.write("#timestamp [ns],w_RS_S_x [rad s^-1],w_RS_S_y [rad s^-1],w_RS_S_z [rad s^-1],"
                "a_RS_S_x [m s^-2],a_RS_S_y [m s^-2],a_RS_S_z [m s^-2]\n")
        for r in np.asarray(imu_rows, np.float64):
            f.write(f"{int(round(r[6] * 1e9))}," + ",".join(repr(float(v)) for v in r[:6]) + "\n")


class EurocAslReader:
    """Iterates (stamp_seconds, image_u8, imu_rows(n, 8)) over an ASL directory, InputBuffer.cc:53-81 pairing."""

    def __init__(self, directory: str, time_offset: float = 0.0, grayscale: bool = True):
        self.dir = directory
        self.time_offset = float(time_offset)
        self.grayscale = grayscale
        cam = os.path.join(directory, "mav0", "cam0")
        self.images = []
        with open(os.path.join(cam, "data.csv")) as f:
            for row in csv.reader(f):
                if not row or row[0].startswith("#"):
                    continue
                self.images.append((int(row[0]) * 1e-9, os.path.join(cam, "data", row[1].strip())))
        rows = []
        with open(os.path.join(directory, "mav0", "imu0", "data.csv")) as f:
            for row in csv.reader(f):
                if not row or row[0].startswith("#"):
                    continue
                rows.append([float(v) for v in row[1:7]] + [int(row[0]) * 1e-9])
        imu = np.array(rows, np.float64).reshape(-1, 7)
        dt = np.zeros(len(imu))
        dt[1:] = np.diff(imu[:, 6])                       # rvio_mono.cc:103-108 (0 for the first message)
        self.imu = np.concatenate([imu, dt[:, None]], 1)  # [w, a, t, dt]

    def __len__(self):
        return len(self.images)

    def __iter__(self):
        import cv2
        consumed = 0
        for t, path in self.images:
            lim = t + self.time_offset
            if len(self.imu) == 0 or self.imu[-1, 6] < lim:   # InputBuffer.cc:59-60: wait for enough IMU
                break
            j = consumed
            while j < len(self.imu) and self.imu[j, 6] <= lim:
                j += 1
            rows = self.imu[consumed:j]
            consumed = j
            if len(rows) < 2:                                  # InputBuffer.cc:76-77: image dropped
                continue
            im = cv2.imread(path, cv2.IMREAD_GRAYSCALE if self.grayscale else cv2.IMREAD_COLOR)
            if im is None:
                raise IOError(f"cannot read {path}")
            yield t, im, np.ascontiguousarray(rows)
