"""ctypes front end of the CPU ORACLE (test infrastructure only -- see oracle/rvio_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
`VioOracle` restates System::MonoVIO's frame loop (src/rvio/System.cc:173-365, init logic :183-249) on top
of the C restatements; the corner detector is cv2 (real OpenCV: FeatureDetector.cc:55-75), outside the C code.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "librvio_oracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


class TrackerCfg(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("k1", C.c_float), ("k2", C.c_float), ("p1", C.c_float), ("p2", C.c_float), ("k3", C.c_float),
                ("n_features", C.c_int), ("max_track_len", C.c_int), ("min_track_len", C.c_int),
                ("enable_equalizer", C.c_int), ("use_sampson", C.c_int),
                ("inlier_thr", C.c_double), ("small_angle", C.c_double), ("T_BC0", C.c_double * 16),
                ("img_w", C.c_int), ("img_h", C.c_int), ("min_dist", C.c_double), ("block_x", C.c_int), ("block_y", C.c_int),
                ("is_fisheye", C.c_int)]


class UpdaterCfg(C.Structure):
    _fields_ = [("sigma", C.c_double), ("Ric", C.c_double * 9), ("tic", C.c_double * 3)]


class ImuCfg(C.Structure):
    _fields_ = [("gravity", C.c_double), ("small_angle", C.c_double), ("sigma_g", C.c_double),
                ("sigma_wg", C.c_double), ("sigma_a", C.c_double), ("sigma_wa", C.c_double)]


class UpdateInfo(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("n_feat", "n_good", "rows_stacked", "rank", "compressed", "updated",
                                        "n_reject_init", "n_reject_lm", "n_reject_gate", "rank_full")]


class RandState(C.Structure):
    _fields_ = [("r", C.c_int32 * 34), ("f", C.c_int), ("b", C.c_int)]


class RansacState(C.Structure):
    _fields_ = [("use_sampson", C.c_int), ("inlier_thr", C.c_double), ("small_angle", C.c_double),
                ("Ric", C.c_double * 9), ("rng", RandState), ("two_points", C.c_int * 32),
                ("n_inliers", C.c_int * 16), ("winner", C.c_int), ("hyp", C.c_double * 144), ("R", C.c_double * 9)]


def _p(dtype):
    return np.ctypeslib.ndpointer(dtype, flags="C_CONTIGUOUS")


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    u8, f32, f64, i32, i16 = _p(np.uint8), _p(np.float32), _p(np.float64), _p(np.int32), _p(np.int16)
    ci, cd, vp = C.c_int, C.c_double, C.c_void_p
    L.orc_clahe.argtypes = [u8, ci, ci, ci, u8, ci]
    L.orc_pyr_down.argtypes = [u8, ci, ci, ci, u8, ci]
    L.orc_scharr.argtypes = [u8, ci, ci, ci, i16]
    L.orc_lk.argtypes = [u8, u8, ci, ci, ci, f32, ci, f32, u8, ci, ci, ci, cd, cd]
    L.orc_lk.restype = ci
    L.orc_min_eig_map.argtypes = [u8, ci, ci, ci, f32]
    L.orc_good_features.argtypes = [u8, ci, ci, ci, ci, cd, cd, f32]
    L.orc_corner_subpix.argtypes = [u8, ci, ci, ci, f32, ci, ci, ci, cd]
    L.orc_detect_with_subpix.argtypes = [u8, ci, ci, ci, ci, ci, cd, cd, f32]
    L.orc_undistort.argtypes = [f32, ci, f32, f32, f32]
    L.orc_undistort_fisheye.argtypes = [f32, ci, f32, f32, f32]
    L.orc_rand_seed.argtypes = [C.POINTER(RandState), C.c_uint]
    L.orc_rand_next.argtypes = [C.POINTER(RandState)]
    L.orc_rand_next.restype = ci
    L.orc_ransac_init.argtypes = [C.POINTER(RansacState), ci, cd, cd, f64]
    L.orc_ransac_find_inliers.argtypes = [C.POINTER(RansacState), f64, f64, ci, f64, ci, u8]
    L.orc_ransac_find_inliers.restype = ci
    L.orc_tracker_create.argtypes = [C.POINTER(TrackerCfg)]
    L.orc_tracker_create.restype = vp
    L.orc_tracker_destroy.argtypes = [vp]
    L.orc_tracker_track.argtypes = [vp, u8, ci, f64, ci]
    L.orc_tracker_track.restype = ci
    L.orc_tracker_track_ext.argtypes = [vp, u8, f32, u8, f64, ci]
    L.orc_tracker_track_ext.restype = ci
    L.orc_tracker_feats.argtypes = [vp]
    L.orc_tracker_feats.restype = C.POINTER(C.c_float)
    L.orc_tracker_n_feats.argtypes = [vp]
    L.orc_tracker_n_feats.restype = ci
    L.orc_tracker_last_image.argtypes = [vp]
    L.orc_tracker_last_image.restype = C.POINTER(C.c_uint8)
    for name, rt in (("image", C.POINTER(C.c_uint8)), ("tracked_px", C.POINTER(C.c_float)),
                     ("update_types", C.POINTER(C.c_uint8)), ("update_offsets", C.POINTER(C.c_int32)),
                     ("update_xy", C.POINTER(C.c_float)), ("last_status", C.POINTER(C.c_uint8)),
                     ("last_flags", C.POINTER(C.c_uint8)), ("last_lk", C.POINTER(C.c_float)),
                     ("last_un", C.POINTER(C.c_float)), ("slots", C.POINTER(C.c_int32)),
                     ("ransac", C.POINTER(RansacState))):
        f = getattr(L, "orc_tracker_" + name)
        f.argtypes = [vp]
        f.restype = rt
    for name in ("n_free", "n_tracked", "n_update", "last_n"):
        f = getattr(L, "orc_tracker_" + name)
        f.argtypes = [vp]
        f.restype = ci
    L.orc_tracker_seed.argtypes = [vp, f32, ci]
    L.orc_tracker_refill.argtypes = [vp, f32, ci]
    L.orc_tracker_refill.restype = ci
    L.orc_tracker_commit.argtypes = [vp]
    L.orc_find_newer.argtypes = [C.POINTER(TrackerCfg), f32, ci, f32, ci, f32]
    L.orc_find_newer.restype = ci
    L.orc_updater_cfg_init.argtypes = [C.POINTER(UpdaterCfg), C.c_float, C.c_float, f64]
    L.orc_updater_update.argtypes = [C.POINTER(UpdaterCfg), f64, ci, f64, u8, i32, f32, ci, f64, f64,
                                     C.POINTER(UpdateInfo), vp, vp, vp, vp, vp]
    L.orc_updater_set_rank_rule.argtypes = [ci]
    L.orc_propagate.argtypes = [C.POINTER(ImuCfg), f64, ci, f64, f64, ci, f64, f64]
    L.orc_augment_compose.argtypes = [f64, f64, C.POINTER(ci), ci, ci, f64]
    L.orc_initialize.argtypes = [C.POINTER(ImuCfg), cd, f64, f64, ci, ci, f64, f64]
    L.orc_quat_mul.argtypes = [f64, f64, f64]
    L.orc_quat_to_rot.argtypes = [f64, f64]
    L.orc_rot_to_quat.argtypes = [f64, f64]
    L.orc_chi2_95.argtypes = [ci]
    L.orc_chi2_95.restype = cd
    _lib = L
    return L


# ----------------------------------------------------------------------------- config helpers
def tracker_cfg(cfg) -> TrackerCfg:
    t = TrackerCfg()
    for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3"):
        setattr(t, k, np.float32(getattr(cfg, k)))          # read into float: Tracker.cc:39-61
    t.n_features, t.max_track_len, t.min_track_len = cfg.n_features, cfg.max_track_len, cfg.min_track_len
    t.enable_equalizer, t.use_sampson = cfg.enable_equalizer, cfg.use_sampson
    t.inlier_thr, t.small_angle = cfg.inlier_thr, cfg.small_angle
    t.T_BC0 = (C.c_double * 16)(*cfg.T_BC0)
    t.img_w, t.img_h, t.min_dist, t.block_x, t.block_y = cfg.width, cfg.height, cfg.min_dist, cfg.block_x, cfg.block_y
    t.is_fisheye = int(getattr(cfg, "fisheye", 0))
    return t


def updater_cfg(cfg) -> UpdaterCfg:
    u = UpdaterCfg()
    lib().orc_updater_cfg_init(C.byref(u), np.float32(cfg.sigma_px), np.float32(cfg.sigma_py),
                               np.array(cfg.T_BC0, np.float64))
    return u


def imu_cfg(cfg) -> ImuCfg:
    return ImuCfg(cfg.gravity, cfg.small_angle, cfg.sigma_g, cfg.sigma_wg, cfg.sigma_a, cfg.sigma_wa)


# ----------------------------------------------------------------------------- thin functional wrappers
def clahe(img):
    out = np.empty_like(img)
    lib().orc_clahe(img, img.shape[1], img.shape[0], img.strides[0], out, out.strides[0])
    return out


def pyr_down(img):
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyr_down(img, w, h, img.strides[0], out, out.strides[0])
    return out


def lk(prev, nxt, pts, win=15, max_level=3, max_iter=30, eps=1e-2, min_eig=1e-3):
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.zeros_like(pts)
    st = np.zeros(len(pts), np.uint8)
    lib().orc_lk(prev, nxt, prev.shape[1], prev.shape[0], prev.strides[0], pts, len(pts), out, st,
                 win, max_level, max_iter, eps, min_eig)
    return out, st


def undistort(px, cfg):
    px = np.ascontiguousarray(px, np.float32)
    K = np.array([cfg.fx, cfg.fy, cfg.cx, cfg.cy], np.float32)
    D = np.array([cfg.k1, cfg.k2, cfg.p1, cfg.p2, cfg.k3], np.float32)
    out = np.empty_like(px)
    if getattr(cfg, "fisheye", 0):
        lib().orc_undistort_fisheye(px, len(px), K, np.ascontiguousarray(D[:4]), out)
    else:
        lib().orc_undistort(px, len(px), K, D, out)
    return out


def updater_update(cfg, x, P, types, offsets, xy, debug=False):
    """Returns (x_out, P_out, info[, dbg]).  P is a (d, d) array (symmetric layout irrelevant: passed column-major)."""
    L = lib()
    u = updater_cfg(cfg)
    x = np.ascontiguousarray(x, np.float64)
    d = P.shape[0]
    Pf = np.ascontiguousarray(np.asarray(P, np.float64).T)      # column-major bytes
    types = np.ascontiguousarray(types, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int32)
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1)
    if xy.size == 0:
        xy = np.zeros(2, np.float32)
    nf = len(types)
    if nf == 0:
        types = np.zeros(1, np.uint8)
    x_out = np.empty_like(x)
    P_out = np.empty_like(Pf)
    info = UpdateInfo()
    n = d - 24
    rows = int(2 * (offsets[nf] - offsets[0])) if nf else 0
    st = np.zeros(max(nf, 1), np.uint8); pf = np.zeros(3 * max(nf, 1)); gm = np.zeros(max(nf, 1))
    Hs = np.zeros(max(rows, 1) * max(n, 1)); rs = np.zeros(max(rows, 1))
    L.orc_updater_update(C.byref(u), x, len(x), Pf, types, offsets, xy, nf, x_out, P_out, C.byref(info),
                         st.ctypes.data, pf.ctypes.data, gm.ctypes.data, Hs.ctypes.data, rs.ctypes.data)
    P_out = P_out.T.copy()
    if debug:
        R = info.rows_stacked
        dbg = dict(status=st[:nf], pfinv=pf[:3 * nf].reshape(-1, 3), gamma=gm[:nf],
                   H=Hs[:R * n].reshape(R, n).copy(), r=rs[:R].copy())
        return x_out, P_out, info, dbg
    return x_out, P_out, info


def propagate(cfg, x, P, imu):
    x = np.ascontiguousarray(x, np.float64)
    Pf = np.ascontiguousarray(np.asarray(P, np.float64).T)
    imu = np.ascontiguousarray(imu, np.float64)
    xo = np.empty_like(x); Po = np.empty_like(Pf)
    c = imu_cfg(cfg)
    lib().orc_propagate(C.byref(c), x, len(x), Pf, imu, len(imu), xo, Po)
    return xo, Po.T.copy()


def detect_with_subpix(img, n_corners, s, cfg):
    """FeatureDetector::DetectWithSubPix (FeatureDetector.cc:55-75) through real OpenCV."""
    import cv2
    c = cv2.goodFeaturesToTrack(img, n_corners, float(np.float32(cfg.qual_lvl)), s * float(np.float32(cfg.min_dist)))
    if c is None or len(c) == 0:
        return np.zeros((0, 2), np.float32)
    c = np.ascontiguousarray(c.reshape(-1, 1, 2), np.float32)
    hw = int(math.floor(.5 * float(np.float32(cfg.min_dist))))
    cv2.cornerSubPix(img, c, (hw, hw), (-1, -1), (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 1e-2))
    return c.reshape(-1, 2)


def detect_restated(img, n_corners, s, cfg):
    """FeatureDetector::DetectWithSubPix through the C restatement (oracle/detector.c) instead of cv2."""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros((max(n_corners, 1), 2), np.float32)
    n = lib().orc_detect_with_subpix(img, img.shape[1], img.shape[0], img.strides[0], n_corners, s,
                                     float(np.float32(cfg.qual_lvl)), float(np.float32(cfg.min_dist)), out)
    return out[:n].copy()


def good_features(img, max_corners, quality, min_dist):
    """cv::goodFeaturesToTrack through the C restatement: integer-pixel corners, strongest first."""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros((max(max_corners, 1) if max_corners > 0 else img.size // 4, 2), np.float32)
    n = lib().orc_good_features(img, img.shape[1], img.shape[0], img.strides[0], max_corners, float(quality), float(min_dist), out)
    return out[:n].copy()


def min_eig_map(img):
    img = np.ascontiguousarray(img, np.uint8)
    e = np.empty(img.shape, np.float32)
    lib().orc_min_eig_map(img, img.shape[1], img.shape[0], img.strides[0], e)
    return e


def find_newer(cfg, corners, ref):
    corners = np.ascontiguousarray(corners, np.float32).reshape(-1, 2)
    ref = np.ascontiguousarray(ref, np.float32).reshape(-1, 2)
    out = np.zeros((max(len(corners), 1), 2), np.float32)
    t = tracker_cfg(cfg)
    n = lib().orc_find_newer(C.byref(t), corners if len(corners) else np.zeros((1, 2), np.float32), len(corners),
                             ref if len(ref) else np.zeros((1, 2), np.float32), len(ref), out)
    return out[:n].copy()


# ----------------------------------------------------------------------------- Tracker / VIO loop
class Tracker:
    def __init__(self, cfg, detector=None):
        self.cfg = cfg
        self._tc = tracker_cfg(cfg)
        self.h = lib().orc_tracker_create(C.byref(self._tc))
        self.detector = detector or (lambda img, n, s: detect_with_subpix(img, n, s, cfg))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tracker_destroy(self.h)
            self.h = None

    def image(self):
        p = lib().orc_tracker_image(self.h)
        return np.ctypeslib.as_array(p, (self.cfg.height, self.cfg.width)).copy()

    def track(self, img, imu):
        """Whole Tracker::track (Tracker.cc:179-396).  Returns status code of orc_tracker_track."""
        L = lib()
        img = np.ascontiguousarray(img, np.uint8)
        imu = np.ascontiguousarray(imu, np.float64).reshape(-1, 8)
        rc = L.orc_tracker_track(self.h, img, img.strides[0], imu, len(imu))
        if rc == 2:
            return rc
        eq = self.image()
        if rc == 1:
            pts = self.detector(eq, self.cfg.n_features, 1)
            if len(pts) == 0:
                return 3                                    # Tracker.cc:209-213 (stays "first image")
            L.orc_tracker_seed(self.h, np.ascontiguousarray(pts, np.float32), len(pts))
        elif L.orc_tracker_n_free(self.h) > 0:
            cand = self.detector(eq, self.cfg.n_features, 2)
            nt = L.orc_tracker_n_tracked(self.h)
            ref = np.ctypeslib.as_array(L.orc_tracker_tracked_px(self.h), (max(nt, 1), 2))[:nt].copy()
            newer = find_newer(self.cfg, cand, ref)
            if len(newer):
                L.orc_tracker_refill(self.h, newer, len(newer))
        L.orc_tracker_commit(self.h)
        return rc

    def update_lists(self):
        L = lib()
        n = L.orc_tracker_n_update(self.h)
        if n == 0:
            return np.zeros(0, np.uint8), np.zeros(1, np.int32), np.zeros((0, 2), np.float32)
        types = np.ctypeslib.as_array(L.orc_tracker_update_types(self.h), (n,)).copy()
        off = np.ctypeslib.as_array(L.orc_tracker_update_offsets(self.h), (n + 1,)).copy()
        xy = np.ctypeslib.as_array(L.orc_tracker_update_xy(self.h), (int(off[-1]), 2)).copy()
        return types, off, xy

    def debug(self):
        L = lib()
        n = L.orc_tracker_last_n(self.h)
        g = lambda f, shape: np.ctypeslib.as_array(f(self.h), shape).copy() if n else np.zeros((0,) + shape[1:])
        return dict(n=n, status=g(L.orc_tracker_last_status, (n,)), flags=g(L.orc_tracker_last_flags, (n,)),
                    lk=g(L.orc_tracker_last_lk, (n, 2)), un=g(L.orc_tracker_last_un, (n, 2)))


class VioOracle:
    """System::MonoVIO frame loop (System.cc:173-365) with per-instance state."""

    def __init__(self, cfg, detector=None):
        self.cfg = cfg
        self.tracker = Tracker(cfg, detector)
        self.moving = False
        self.ready = False
        self.wm = np.zeros(3); self.am = np.zeros(3); self.n_imu_count = 0
        self.n_clones = 0
        self.n_img_after_init = 0
        W = cfg.window
        self.x = np.zeros(26 + 7 * W)
        self.Pbuf = np.zeros((24 + 6 * W) ** 2)
        self.last_info = None
        self.timing = []

    @property
    def xdim(self):
        return 26 + 7 * self.n_clones

    @property
    def d(self):
        return 24 + 6 * self.n_clones

    def state(self):
        d = self.d
        return self.x[:self.xdim].copy(), self.Pbuf[:d * d].reshape(d, d).T.copy()

    def _init(self, imu):
        cfg = self.cfg
        imu = [r for r in imu]
        if not self.moving:            # System.cc:191-217
            ang = np.zeros(3); vel = np.zeros(3); displ = np.zeros(3)
            for r in imu:
                w, a, dt = r[0:3], r[3:6].copy(), r[7]
                a = a - cfg.gravity * a / np.linalg.norm(a)
                ang += dt * w
                vel += dt * a
                displ += dt * vel + .5 * dt * dt * a
            if np.linalg.norm(ang) > cfg.thr_angle or np.linalg.norm(displ) > cfg.thr_displ:
                self.moving = True
        while imu:                     # System.cc:219-245
            if not self.moving:
                r = imu.pop(0)
                self.wm += r[0:3]; self.am += r[3:6]
                self.n_imu_count += 1
            else:
                if self.n_imu_count == 0:
                    self.wm = imu[0][0:3].copy(); self.am = imu[0][3:6].copy()
                    self.n_imu_count = 1
                else:
                    self.wm = self.wm / self.n_imu_count
                    self.am = self.am / self.n_imu_count
                x = np.zeros(26); P = np.zeros(24 * 24)
                c = imu_cfg(cfg)
                lib().orc_initialize(C.byref(c), cfg.imu_rate, np.ascontiguousarray(self.wm), np.ascontiguousarray(self.am),
                                     self.n_imu_count, cfg.enable_alignment, x, P)
                self.x[:26] = x
                self.Pbuf[:576] = P
                self.ready = True
                break
        return self.ready, np.array(imu).reshape(-1, 8)

    def step(self, img, imu, timer=None):
        """One MonoVIO call.  imu: (n,8).  Returns pose [pGk(3), qkG(4)] or None while initialising."""
        import time
        if len(imu) < 2:
            return None                    # InputBuffer.cc:76-77
        if not self.ready:
            ok, imu = self._init(np.asarray(imu, np.float64))
            if not ok:
                return None
        self.n_img_after_init += 1
        L = lib()
        t1 = time.perf_counter()
        self.tracker.track(img, imu)
        t2 = time.perf_counter()
        imu = np.ascontiguousarray(imu, np.float64).reshape(-1, 8)
        d = self.d
        x = self.x[:self.xdim].copy()
        P = self.Pbuf[:d * d].copy()
        xo = np.empty_like(x); Po = np.empty_like(P)
        c = imu_cfg(self.cfg)
        L.orc_propagate(C.byref(c), x, len(x), P, imu, len(imu), xo, Po)
        if self.n_clones > self.cfg.min_clones:          # System.cc:266-277
            types, off, xy = self.tracker.update_lists()
            u = updater_cfg(self.cfg)
            info = UpdateInfo()
            x2 = np.empty_like(xo); P2 = np.empty_like(Po)
            nf = len(types)
            L.orc_updater_update(C.byref(u), xo, len(xo), Po, types if nf else np.zeros(1, np.uint8), off,
                                 xy.reshape(-1) if nf else np.zeros(2, np.float32), nf, x2, P2, C.byref(info),
                                 None, None, None, None, None)
            self.last_info = info
            self.last_update_in = (xo.copy(), Po.copy(), types, off, xy)
            xo, Po = x2, P2
        self.x[:len(xo)] = xo
        self.Pbuf[:len(Po)] = Po
        n = C.c_int(self.n_clones)
        pose = np.zeros(7)
        L.orc_augment_compose(self.x, self.Pbuf, C.byref(n), self.cfg.window, 1 if self.n_img_after_init > 1 else 0, pose)
        self.n_clones = n.value
        t3 = time.perf_counter()
        self.timing.append((1e3 * (t2 - t1), 1e3 * (t3 - t2)))
        return pose
