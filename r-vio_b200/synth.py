"""Seeded synthetic EuRoC-shaped mono+IMU streams and the config surface.

Not part of the hot path: this only manufactures the bytes that are fed identically to the CUDA
path, the CPU oracle and the reference arm.  Config keys are the reference's YAML keys
(config/rvio_euroc.yaml of the reference); the default values are EuRoC V1_01's.

Scene: two textured planes (a wall 4 m in front of the start pose and a floor 1.5 m below it) rendered
through the radtan camera model per pose; IMU = analytic body rates / specific force of a smooth
sum-of-sinusoids trajectory (2 s static prefix) + white noise + bias random walk.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

try:  # cv2 is only used to blur the texture and to resample it (data generation, not compute under test)
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


@dataclasses.dataclass
class Config:
    # IMU.*  (rvio_euroc.yaml:8-20)
    imu_rate: float = 200.0
    sigma_g: float = 1.6968e-04
    sigma_wg: float = 1.9393e-05
    sigma_a: float = 2.0000e-3
    sigma_wa: float = 3.0000e-3
    gravity: float = 9.8082
    small_angle: float = 0.001745329
    # Camera.*  (rvio_euroc.yaml:27-65)
    fps: float = 20.0
    width: int = 752
    height: int = 480
    fx: float = 458.654
    fy: float = 457.296
    cx: float = 367.215
    cy: float = 248.375
    k1: float = -0.28340811
    k2: float = 0.07395907
    p1: float = 0.00019359
    p2: float = 1.76187114e-05
    k3: float = 0.0
    fisheye: int = 0              # Camera.Fisheye (k1, k2, p1, p2 are then the equidistant k1..k4, Tracker.cc:119)
    sigma_px: float = 0.002180293
    sigma_py: float = 0.002186767
    T_BC0: tuple = (0.0148655429818, -0.999880929698, 0.00414029679422, -0.0216401454975,
                    0.999557249008, 0.0149672133247, 0.025715529948, -0.064676986768,
                    -0.0257744366974, 0.00375618835797, 0.999660727178, 0.00981073058949,
                    0.0, 0.0, 0.0, 1.0)
    time_offset: float = 0.0
    # Tracker.*  (rvio_euroc.yaml:72-97)
    n_features: int = 200
    max_track_len: int = 15
    min_track_len: int = 3
    min_dist: float = 15.0
    qual_lvl: float = 0.01
    block_x: int = 150
    block_y: int = 120
    enable_equalizer: int = 1
    use_sampson: int = 1
    inlier_thr: float = 1e-5
    # INI.*  (rvio_euroc.yaml:104-111)
    thr_angle: float = 0.005
    thr_displ: float = 0.01
    enable_alignment: int = 1

    # reference YAML key -> field (config/rvio_euroc.yaml; System.cc:53-83, Tracker.cc:37-90, Updater.cc:38-53)
    YAML_KEYS = {
        "IMU.dps": "imu_rate", "IMU.sigma_g": "sigma_g", "IMU.sigma_wg": "sigma_wg", "IMU.sigma_a": "sigma_a",
        "IMU.sigma_wa": "sigma_wa", "IMU.nG": "gravity", "IMU.nSmallAngle": "small_angle",
        "Camera.fps": "fps", "Camera.Fisheye": "fisheye", "Camera.width": "width", "Camera.height": "height",
        "Camera.fx": "fx", "Camera.fy": "fy", "Camera.cx": "cx", "Camera.cy": "cy", "Camera.k1": "k1", "Camera.k2": "k2",
        "Camera.p1": "p1", "Camera.p2": "p2", "Camera.k3": "k3", "Camera.sigma_px": "sigma_px", "Camera.sigma_py": "sigma_py",
        "Camera.T_BC0": "T_BC0", "Camera.nTimeOffset": "time_offset",
        "Tracker.nFeatures": "n_features", "Tracker.nMaxTrackingLength": "max_track_len",
        "Tracker.nMinTrackingLength": "min_track_len", "Tracker.nMinDist": "min_dist", "Tracker.nQualLvl": "qual_lvl",
        "Tracker.nBlockSizeX": "block_x", "Tracker.nBlockSizeY": "block_y", "Tracker.EnableEqualizer": "enable_equalizer",
        "Tracker.UseSampson": "use_sampson", "Tracker.nInlierThrd": "inlier_thr",
        "INI.nThresholdAngle": "thr_angle", "INI.nThresholdDispl": "thr_displ", "INI.EnableAlignment": "enable_alignment",
    }

    @classmethod
    def from_yaml(cls, path: str) -> "Config":
        """Reads the reference's OpenCV-FileStorage YAML (the `%YAML:1.0` header and `!!opencv-matrix` tags are
        stripped; a matrix keeps its `data` list)."""
        import yaml
        text = open(path).read()
        text = "\n".join(l for l in text.splitlines() if not l.startswith("%YAML")).replace("!!opencv-matrix", "")
        doc = yaml.safe_load(text) or {}
        c = cls()
        types = {f.name: f.type for f in dataclasses.fields(cls)}
        for key, field in cls.YAML_KEYS.items():
            if key not in doc:
                continue
            v = doc[key]
            if isinstance(v, dict):
                v = tuple(float(e) for e in v["data"])
            elif types.get(field) in (int, "int"):
                v = int(v)
            else:
                v = float(v)
            setattr(c, field, v)
        return c

    @property
    def window(self) -> int:          # System.cc:71-72
        return self.max_track_len - 1

    @property
    def min_clones(self) -> int:      # System.cc:74-75
        return self.min_track_len - 1


def baseline_config(idx: int) -> Config:
    """BASELINE.json configs[idx] -> Config ("N-clone window" == nMaxTrackingLength N+1)."""
    if idx == 0:
        return Config(n_features=150, max_track_len=11)
    if idx in (1, 3):
        return Config(n_features=200, max_track_len=12)
    if idx == 2:
        sx, sy = 1280 / 752, 720 / 480
        c = Config(n_features=600, max_track_len=26, width=1280, height=720)
        c.fx *= sx; c.cx *= sx; c.fy *= sy; c.cy *= sy
        c.block_x = int(150 * sx); c.block_y = int(120 * sy)
        return c
    if idx == 4:
        return Config(n_features=2048, max_track_len=31, min_dist=5.0)
    raise ValueError(idx)


# ----------------------------------------------------------------------------- trajectory
_BASE_R_WI = np.array([[0.0, 0.0, 1.0], [0.0, -1.0, 0.0], [1.0, 0.0, 0.0]])  # body x up, y right, z forward(+x world)


def _smoothstep(t, t0, t1):
    s = np.clip((t - t0) / (t1 - t0), 0.0, 1.0)
    return s * s * s * (s * (6 * s - 15) + 10)


def _rot_xyz(a):
    cx, sx, cy, sy, cz, sz = math.cos(a[0]), math.sin(a[0]), math.cos(a[1]), math.sin(a[1]), math.cos(a[2]), math.sin(a[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


class Trajectory:
    """Smooth 6-DoF motion: static for t < t_static, then sum of sinusoids (<= ~1 m/s, <= ~0.5 rad/s)."""

    def __init__(self, seed: int, t_static: float = 2.0):
        r = np.random.default_rng(seed)
        self.t_static = t_static
        self.pa = r.uniform(0.15, 0.35, (3, 2))          # position amplitudes [m]
        self.pf = r.uniform(0.15, 0.45, (3, 2))          # frequencies [Hz]
        self.pp = r.uniform(0, 2 * np.pi, (3, 2))
        self.aa = r.uniform(0.05, 0.12, (3, 2))          # angle amplitudes [rad]
        self.af = r.uniform(0.1, 0.35, (3, 2))
        self.ap = r.uniform(0, 2 * np.pi, (3, 2))

    def _pos_ang(self, t):
        ramp = _smoothstep(t, self.t_static, self.t_static + 2.0)
        tt = max(t - self.t_static, 0.0)
        p = (self.pa * (np.sin(2 * np.pi * self.pf * tt + self.pp) - np.sin(self.pp))).sum(1) * ramp
        a = (self.aa * (np.sin(2 * np.pi * self.af * tt + self.ap) - np.sin(self.ap))).sum(1) * ramp
        return p, a

    def pose(self, t):
        p, a = self._pos_ang(t)
        return _rot_xyz(a) @ _BASE_R_WI, p

    def imu(self, t, grav):
        """Ideal (omega_body, specific force in body) by central differences of the analytic pose."""
        h = 1e-4
        R0, _ = self.pose(t)
        Rp, pp = self.pose(t + h)
        Rm, pm = self.pose(t - h)
        _, p0 = self.pose(t)
        acc = (pp - 2 * p0 + pm) / (h * h)
        dR = R0.T @ (Rp - Rm) / (2 * h)
        w = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) * 0.5
        f = R0.T @ (acc + np.array([0.0, 0.0, grav]))
        return w, f


# ----------------------------------------------------------------------------- renderer
def _texture(seed: int, size: int = 2048):
    r = np.random.default_rng(seed)
    acc = np.zeros((size, size), np.float32)
    for sig, wgt in ((1.6, 1.0), (4.0, 0.8), (11.0, 0.7)):
        n = r.standard_normal((size, size)).astype(np.float32)
        n = cv2.GaussianBlur(n, (0, 0), sig)
        acc += wgt * n / n.std()
    acc = (acc - acc.mean()) / acc.std()
    return np.clip(127.0 + 52.0 * acc, 0, 255).astype(np.float32)


class Stream:
    """frames[i] (u8 HxW), frame_t[i]; imu rows [w(3), a(3), t, dt] (dt = t - t_prev, 0 for the first: rvio_mono.cc:102-107)."""

    def __init__(self, cfg: Config, n_frames: int, seed: int, tex_px_per_m: float = 150.0, t_static: float = 2.0):
        if cv2 is None:
            raise RuntimeError("cv2 is required to render synthetic frames")
        self.cfg, self.seed, self.n_frames = cfg, seed, n_frames
        self.traj = Trajectory(seed, t_static)
        T = np.array(cfg.T_BC0, np.float64).reshape(4, 4)
        self.R_IC, self.p_IC = T[:3, :3], T[:3, 3]
        self.tex_wall = _texture(seed * 7 + 1)
        self.tex_floor = _texture(seed * 7 + 2)
        self.ppm = tex_px_per_m
        W, H = cfg.width, cfg.height
        K = np.array([[cfg.fx, 0, cfg.cx], [0, cfg.fy, cfg.cy], [0, 0, 1]], np.float64)
        D = np.array([cfg.k1, cfg.k2, cfg.p1, cfg.p2, cfg.k3], np.float64)
        uu, vv = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
        pts = np.stack([uu.ravel(), vv.ravel()], 1).reshape(-1, 1, 2)
        un = cv2.undistortPointsIter(pts, K, D, np.eye(3), np.eye(3), (cv2.TERM_CRITERIA_COUNT, 40, 0)).reshape(H, W, 2)
        self.rays = np.concatenate([un, np.ones((H, W, 1))], 2)           # camera-frame rays
        rng = np.random.default_rng(seed + 1000)
        # IMU
        rate = cfg.imu_rate
        t_end = (n_frames - 1) / cfg.fps + 1e-9
        n_imu = int(math.floor(t_end * rate)) + 1
        imu = np.zeros((n_imu, 8))
        bg = np.zeros(3); ba = np.zeros(3)
        for k in range(n_imu):
            t = k / rate
            w, f = self.traj.imu(t, cfg.gravity)
            bg = bg + rng.standard_normal(3) * cfg.sigma_wg / math.sqrt(rate)
            ba = ba + rng.standard_normal(3) * cfg.sigma_wa / math.sqrt(rate)
            imu[k, 0:3] = w + bg + rng.standard_normal(3) * cfg.sigma_g * math.sqrt(rate)
            imu[k, 3:6] = f + ba + rng.standard_normal(3) * cfg.sigma_a * math.sqrt(rate)
            imu[k, 6] = t
            imu[k, 7] = 0.0 if k == 0 else 1.0 / rate
        self.imu = imu
        self.frame_t = np.arange(n_frames) / cfg.fps
        self._noise_rng = np.random.default_rng(seed + 2000)
        self.frames = [self._render(t) for t in self.frame_t]

    def cam_pose(self, t):
        R_WI, p_WI = self.traj.pose(t)
        return R_WI @ self.R_IC, p_WI + R_WI @ self.p_IC

    def _render(self, t):
        cfg = self.cfg
        R_WC, p_WC = self.cam_pose(t)
        d = self.rays @ R_WC.T                                           # world rays HxWx3
        size = self.tex_wall.shape[0]
        half = size / 2.0
        with np.errstate(divide="ignore", invalid="ignore"):
            tw = (4.0 - p_WC[0]) / d[..., 0]                             # wall x = 4
            tf = (-1.5 - p_WC[2]) / d[..., 2]                            # floor z = -1.5
        tw = np.where(tw > 0, tw, np.inf)
        tf = np.where(tf > 0, tf, np.inf)
        use_wall = tw <= tf
        pw = p_WC + d * np.where(np.isfinite(tw), tw, 0)[..., None]
        pf = p_WC + d * np.where(np.isfinite(tf), tf, 0)[..., None]
        mx_w = (pw[..., 1] * self.ppm + half).astype(np.float32)
        my_w = (pw[..., 2] * self.ppm + half).astype(np.float32)
        mx_f = (pf[..., 1] * self.ppm + half).astype(np.float32)
        my_f = (pf[..., 0] * self.ppm * 0.6 + half * 0.3).astype(np.float32)
        iw = cv2.remap(self.tex_wall, mx_w, my_w, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        fl = cv2.remap(self.tex_floor, mx_f, my_f, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        img = np.where(use_wall, iw, fl)
        img = img + self._noise_rng.standard_normal(img.shape).astype(np.float32) * 1.0
        return np.clip(np.rint(img), 0, 255).astype(np.uint8)

    def imu_for_frame(self, i, consumed):
        """InputBuffer::GetMeasurements (InputBuffer.cc:53-81): all not-yet-consumed IMU with t <= t_img + offset."""
        t = self.frame_t[i] + self.cfg.time_offset
        j = consumed
        while j < len(self.imu) and self.imu[j, 6] <= t + 1e-12:
            j += 1
        return self.imu[consumed:j], j

    def gt_pose(self, i):
        return self.traj.pose(self.frame_t[i])


# ----------------------------------------------------------------------------- updater micro-benchmark inputs
def make_update_case(cfg, n_feat, seed, track_len=None, mix_types=True, n_clones=None):
    """Worst-case-shaped Updater::update inputs: a smooth window of N relative poses, P = A A^T + eps I scaled
    to diag ~ [1e-6 .. 1e-2], `n_feat` tracks of length `track_len` (default: maximum) from random 3-D points at 2..10 m
    projected through the true relative poses plus N(0, sigma_px^2) noise.  Returns (x, P, types, offsets, xy)."""
    r = np.random.default_rng(seed)
    N = cfg.window if n_clones is None else n_clones
    L = (N + 1) if track_len is None else track_len
    T = np.array(cfg.T_BC0, np.float64).reshape(4, 4)
    Ric, tic = T[:3, :3], T[:3, 3]
    # relative poses between consecutive frames (JPL): q_i rotates frame i-1 into frame i, p_i = position of frame i in i-1
    x = np.zeros(26 + 7 * N)
    x[3] = 1.0; x[7:10] = [0, 0, 1.0]; x[13] = 1.0
    Rs, ps = [], []
    for c in range(N):
        w = r.normal(0, 0.02, 3); ang = np.linalg.norm(w); k = w / ang
        q = np.concatenate([k * np.sin(ang / 2), [np.cos(ang / 2)]])
        p = r.normal(0, 0.04, 3) + np.array([0.0, 0.0, 0.03])
        x[26 + 7 * c:30 + 7 * c] = q
        x[30 + 7 * c:33 + 7 * c] = p
        qx = np.array([[0, -q[2], q[1]], [q[2], 0, -q[0]], [-q[1], q[0], 0]])
        Rs.append(np.eye(3) - 2 * q[3] * qx + 2 * qx @ qx); ps.append(p)
    d = 24 + 6 * N
    A = r.standard_normal((d, d)) * 0.05
    scale = 10 ** r.uniform(-3.0, -1.0, d)
    P = (A @ A.T + np.eye(d)) * np.outer(scale, scale)
    P = .5 * (P + P.T)
    sig = float(max(np.float32(cfg.sigma_px), np.float32(cfg.sigma_py)))
    types, offsets, xy = [], [0], []
    for f in range(n_feat):
        t2 = mix_types and (f % 2 == 1)
        Lf = (N + 1) if t2 else L
        first = 0 if t2 else N - (Lf - 1)                     # first clone used (type '1': last L-1 clones)
        # point in the first camera frame
        pc = np.array([r.uniform(-1.5, 1.5), r.uniform(-1.0, 1.0), r.uniform(2.0, 10.0)])
        pi = Ric @ pc + tic                                    # in IMU frame of the first image
        meas = []
        Racc, tacc = np.eye(3), np.zeros(3)
        for i in range(Lf):
            if i > 0:
                c = first + i - 1
                Racc = Rs[c] @ Racc
                tacc = Rs[c] @ (tacc - ps[c])
            pcam = Ric.T @ (Racc @ pi + tacc - tic)
            meas.append([pcam[0] / pcam[2] + r.normal(0, sig), pcam[1] / pcam[2] + r.normal(0, sig)])
        types.append(ord('2') if t2 else ord('1'))
        xy.extend(meas)
        offsets.append(len(xy))
    return (x, P, np.array(types, np.uint8), np.array(offsets, np.int32), np.array(xy, np.float32).reshape(-1, 2))
