"""Host-side mirror of the reference's Tracker / Updater call surface on top of the C ABI.

Same names, argument meaning and result members as src/rvio/Tracker.h:43-127 and src/rvio/Updater.h:36-71:
    Tracker.track(im, lImuData)         -> mvFeatTypesForUpdate, mvlFeatMeasForUpdate (CSR here)
    Updater.update(xk1k, Pk1k, types, meas) -> xk1k1, Pk1k1
The corner detector stays on the host exactly as in the reference (FeatureDetector is not on the hot path):
a `detector(img, n_corners, s) -> (k,2) float32` callable plays FeatureDetector::DetectWithSubPix and
`find_newer` plays FeatureDetector::FindNewer (FeatureDetector.cc:97-150).
The C++ twin of this file (what a reference maintainer would compile in) is r-vio_b200/host/rvio_host.hpp.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi


def find_newer(cfg, corners, ref):
    """FeatureDetector::ChessGrid + FindNewer (FeatureDetector.cc:78-150), host logic (float32/int semantics)."""
    f32 = np.float32
    min_dist = f32(cfg.min_dist)
    bx, by = f32(cfg.block_x), f32(cfg.block_y)
    W, H = cfg.width, cfg.height
    gc, gr = int(math.floor(W / float(bx))), int(math.floor(H / float(by)))
    nb = gc * gr
    offx = int(.5 * (W - gc * float(bx)))
    offy = int(.5 * (H - gr * float(by)))
    max_per_block = int(f32(cfg.n_features) / f32(nb))
    grid = [[] for _ in range(nb)]

    def inside(p):
        return not (p[0] <= offx or p[1] <= offy or p[0] >= (W - offx) or p[1] >= (H - offy))

    for p in np.asarray(ref, f32).reshape(-1, 2):
        if not inside(p):
            continue
        col = int(math.floor(f32(p[0] - f32(offx)) / bx))
        row = int(math.floor(f32(p[1] - f32(offy)) / by))
        grid[row * gc + col].append(p)
    out = []
    for p in np.asarray(corners, f32).reshape(-1, 2):
        if not inside(p):
            continue
        col = int(math.floor(f32(p[0] - f32(offx)) / bx))
        row = int(math.floor(f32(p[1] - f32(offy)) / by))
        xl = f32(f32(col) * bx + f32(offx)); xr = f32(xl + bx)
        yt = f32(f32(row) * by + f32(offy)); yb = f32(yt + by)
        if (abs(float(f32(p[0] - xl))) < min_dist or abs(float(f32(p[0] - xr))) < min_dist or
                abs(float(f32(p[1] - yt))) < min_dist or abs(float(f32(p[1] - yb))) < min_dist):
            continue
        cell = grid[row * gc + col]
        if float(len(cell)) < .75 * max_per_block:
            ok = True
            for q in cell:
                dx, dy = float(f32(p[0] - q[0])), float(f32(p[1] - q[1]))
                if not math.sqrt(dx * dx + dy * dy) > float(min_dist):
                    ok = False
                    break
            if ok:
                out.append(p)
                cell.append(p)
    return np.array(out, f32).reshape(-1, 2)


class Tracker:
    """RVIO::Tracker drop-in (device side behind rvio_tracker_*)."""

    def __init__(self, cfg, device: int = 0, detector=None):
        self.cfg = cfg
        self.L = capi.lib()
        self._cfg_c = capi.tracker_cfg(cfg)
        h = C.c_void_p()
        capi.check(self.L.rvio_tracker_create(C.byref(self._cfg_c), device, C.byref(h)), "rvio_tracker_create")
        self.h = h
        self.detector = detector
        self.mvFeatTypesForUpdate = np.zeros(0, np.uint8)
        self.mvlFeatMeasForUpdate = (np.zeros(1, np.int32), np.zeros((0, 2), np.float32))   # (offsets, xy)
        self._eq = np.empty((cfg.height, cfg.width), np.uint8)
        self.fetch_lists = True

    def close(self):
        if getattr(self, "h", None):
            self.L.rvio_tracker_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- pieces -----------------------------------------------------------------------------
    def equalized_image(self):
        capi.check(self.L.rvio_tracker_get_image(self.h, self._eq, self._eq.strides[0]), "get_image")
        return self._eq

    def _detect(self, s):
        if isinstance(self.detector, str) and self.detector == "device":
            return self.detect(s)
        return self.detector(self.equalized_image(), self.cfg.n_features, s)

    def detect(self, s):
        """FeatureDetector::DetectWithSubPix(equalised current image, nFeatures, s) on the device (rvio_tracker_detect)."""
        out = np.zeros((self.cfg.n_features, 2), np.float32)
        n = C.c_int()
        capi.check(self.L.rvio_tracker_detect(self.h, int(s), float(np.float32(self.cfg.min_dist)), float(np.float32(self.cfg.qual_lvl)),
                                              out, C.byref(n)), "rvio_tracker_detect")
        return out[:n.value].copy()

    def n_free(self):
        n = C.c_int()
        capi.check(self.L.rvio_tracker_n_free(self.h, C.byref(n)))
        return n.value

    def tracked_px(self):
        buf = np.empty((self.cfg.n_features, 2), np.float32)
        n = C.c_int()
        capi.check(self.L.rvio_tracker_get_tracked_px(self.h, buf, C.byref(n)))
        return buf[:n.value].copy()

    def update_lists(self):
        nf, nm = C.c_int(), C.c_int()
        capi.check(self.L.rvio_tracker_get_update_count(self.h, C.byref(nf), C.byref(nm)))
        nf, nm = nf.value, nm.value
        types = np.zeros(max(nf, 1), np.uint8)
        off = np.zeros(nf + 1, np.int32)
        xy = np.zeros((max(nm, 1), 2), np.float32)
        capi.check(self.L.rvio_tracker_get_update_lists(self.h, types, off, xy))
        return types[:nf], off, xy[:nm]

    def debug(self):
        F = self.cfg.n_features
        n = C.c_int()
        st = np.zeros(F, np.uint8); fl = np.zeros(F, np.uint8)
        lk = np.zeros((F, 2), np.float32); un = np.zeros((F, 2), np.float32); sl = np.zeros(F, np.int32)
        capi.check(self.L.rvio_tracker_get_debug(self.h, C.byref(n), st, fl, lk, un, sl))
        m = n.value
        return dict(n=m, status=st[:m], flags=fl[:m], lk=lk[:m], un=un[:m], slots=sl[:m])

    def ransac_debug(self):
        tp = np.zeros(32, np.int32); ni = np.zeros(16, np.int32); hy = np.zeros(144)
        w, nc = C.c_int(), C.c_int()
        capi.check(self.L.rvio_tracker_get_ransac_debug(self.h, tp, ni, C.byref(w), C.byref(nc), hy))
        return dict(two_points=tp, n_inliers=ni, winner=w.value, n_cand=nc.value, hyp=hy.reshape(16, 3, 3))

    def pyramid(self, which, level):
        lw, lh = C.c_int(), C.c_int()
        capi.check(self.L.rvio_tracker_get_pyramid(self.h, which, level, None, C.byref(lw), C.byref(lh)))
        out = np.empty((lh.value, lw.value), np.uint8)
        capi.check(self.L.rvio_tracker_get_pyramid(self.h, which, level, out.ctypes.data, C.byref(lw), C.byref(lh)))
        return out

    # -- Tracker::track ---------------------------------------------------------------------
    def track(self, im, lImuData, detections=None, shard=None):
        """Tracker::track (Tracker.cc:179-396).  `detections`: optional pre-computed corner list that replaces the
        detector call for this frame (seed list on the first image, FindNewer output afterwards).
        `shard` = (rank, world, exchange): feature-sharded form -- LK runs for this rank's share of the feature
        indices only, `exchange(lk_ptr, un_ptr, status_ptr, shard_size)` all-gathers the three per-feature arrays in
        place (device pointers), RANSAC + bookkeeping then run replicated."""
        im = np.ascontiguousarray(im, np.uint8)
        ch = 1 if im.ndim == 2 else im.shape[2]
        imu = np.ascontiguousarray(lImuData, np.float64).reshape(-1, 8)
        if shard is None:
            rc = capi.check(self.L.rvio_tracker_track(self.h, im.reshape(-1), im.shape[1], im.shape[0], im.strides[0], ch,
                                                      imu.ctypes.data, len(imu)), "rvio_tracker_track")
        else:
            rank, world, exchange = shard
            rc = capi.check(self.L.rvio_tracker_track_begin(self.h, im.reshape(-1), im.shape[1], im.shape[0], im.strides[0], ch,
                                                            imu.ctypes.data, len(imu), rank, world), "rvio_tracker_track_begin")
            if rc == capi.OK:
                lk, un, st, S = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
                capi.check(self.L.rvio_tracker_lk_results(self.h, world, C.byref(lk), C.byref(un), C.byref(st), C.byref(S)))
                exchange(lk.value, un.value, st.value, S.value)
                capi.check(self.L.rvio_tracker_track_finish(self.h), "rvio_tracker_track_finish")
        if rc == capi.NO_FEATURES:
            return rc
        if rc == capi.FIRST_IMAGE:
            pts = detections if detections is not None else self._detect(1)
            pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
            capi.check(self.L.rvio_tracker_seed(self.h, pts if len(pts) else np.zeros((1, 2), np.float32), len(pts)))
        else:
            if self.fetch_lists:
                t, o, xy = self.update_lists()
                self.mvFeatTypesForUpdate, self.mvlFeatMeasForUpdate = t, (o, xy)
            if self.n_free() > 0:
                if detections is not None:
                    newer = np.ascontiguousarray(detections, np.float32).reshape(-1, 2)
                else:
                    cand = self._detect(2)
                    newer = find_newer(self.cfg, cand, self.tracked_px())
                if len(newer):
                    used = C.c_int()
                    capi.check(self.L.rvio_tracker_refill(self.h, np.ascontiguousarray(newer), len(newer), C.byref(used)))
        capi.check(self.L.rvio_tracker_commit(self.h), "rvio_tracker_commit")
        return rc


def shard_range(n_features: int, rank: int, world: int):
    """Feature indices [lo, hi) whose LK rank `rank` of `world` computes (rvio_tracker_track_begin): equal shards of
    ceil(nFeatures / world) indices, so that the exchange is one fixed-size all-gather."""
    S = -(-n_features // world)
    return rank * S, min((rank + 1) * S, n_features), S


class Updater:
    """RVIO::Updater drop-in (device side behind rvio_updater_*)."""

    def __init__(self, cfg, device: int = 0):
        self.cfg = cfg
        self.L = capi.lib()
        self._cfg_c = capi.updater_cfg(cfg)
        h = C.c_void_p()
        capi.check(self.L.rvio_updater_create(C.byref(self._cfg_c), device, C.byref(h)), "rvio_updater_create")
        self.h = h
        self.xk1k1 = np.zeros(26)
        self.Pk1k1 = np.zeros((24, 24))
        self.info = capi.UpdateInfo()

    def close(self):
        if getattr(self, "h", None):
            self.L.rvio_updater_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update(self, xk1k, Pk1k, vFeatTypesForUpdate, vlFeatMeasForUpdate):
        """Updater::update (Updater.cc:72-628).  Pk1k: (d,d) array; measurements as (offsets, xy) CSR."""
        off, xy = vlFeatMeasForUpdate
        x = np.ascontiguousarray(xk1k, np.float64)
        d = Pk1k.shape[0]
        Pc = np.ascontiguousarray(np.asarray(Pk1k, np.float64).T)           # column-major bytes
        types = np.ascontiguousarray(vFeatTypesForUpdate, np.uint8)
        nf = len(types)
        off = np.ascontiguousarray(off, np.int32)
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1)
        xo = np.empty_like(x); Po = np.empty_like(Pc)
        capi.check(self.L.rvio_updater_update(self.h, x, len(x), Pc, d, types if nf else np.zeros(1, np.uint8), off,
                                              xy if xy.size else np.zeros(2, np.float32), nf, xo, Po, C.byref(self.info)),
                   "rvio_updater_update")
        self.xk1k1, self.Pk1k1 = xo, Po.T.copy()
        return self.xk1k1, self.Pk1k1

    def update_from_tracker(self, xk1k, Pk1k, tracker: Tracker):
        x = np.ascontiguousarray(xk1k, np.float64)
        d = Pk1k.shape[0]
        Pc = np.ascontiguousarray(np.asarray(Pk1k, np.float64).T)
        xo = np.empty_like(x); Po = np.empty_like(Pc)
        capi.check(self.L.rvio_updater_update_from_tracker(self.h, tracker.h, x, len(x), Pc, d, xo, Po, C.byref(self.info)),
                   "rvio_updater_update_from_tracker")
        self.xk1k1, self.Pk1k1 = xo, Po.T.copy()
        return self.xk1k1, self.Pk1k1

    def set_rank_rule(self, full_information: bool):
        """False (default): the reference's first-small-row cut (Updater.cc:515-524); True: keep every row."""
        capi.check(self.L.rvio_updater_set_rank_rule(self.h, 1 if full_information else 0))

    def debug(self, n_feat):
        st = np.zeros(max(n_feat, 1), np.uint8); pf = np.zeros(3 * max(n_feat, 1)); gm = np.zeros(max(n_feat, 1))
        dof = np.zeros(max(n_feat, 1), np.int32)
        capi.check(self.L.rvio_updater_get_debug(self.h, n_feat, st, pf, gm, dof))
        return dict(status=st[:n_feat], pfinv=pf[:3 * n_feat].reshape(-1, 3), gamma=gm[:n_feat], dof=dof[:n_feat])

    def normal_terms(self, n):
        G = np.zeros((n, n)); z = np.zeros(n)
        capi.check(self.L.rvio_updater_get_normal_terms(self.h, G, z, n))
        return G, z


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    capi.check(capi.lib().rvio_b200_nccl_unique_id(buf), "rvio_b200_nccl_unique_id")
    return buf.raw


class Vio:
    """Fused per-frame pipeline (rvio_vio_*): System::MonoVIO with x, P, pyramids and feature lists on the device."""

    def __init__(self, cfg, device: int = 0):
        self.cfg = cfg
        self.L = capi.lib()
        self._cfg_c = capi.vio_cfg(cfg)
        h = C.c_void_p()
        capi.check(self.L.rvio_vio_create(C.byref(self._cfg_c), device, C.byref(h)), "rvio_vio_create")
        self.h = h
        self._pose = np.zeros(7)
        self._valid = C.c_int()

    def close(self):
        if getattr(self, "h", None):
            self.L.rvio_vio_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, im, imu, cand=None, cand_filtered=False, device_detector=False):
        """One frame.  cand: detector corners computed by the caller; device_detector=True: FeatureDetector::DetectWithSubPix
        runs on the GPU inside the step instead (n_cand = -1 at the C ABI)."""
        im = np.ascontiguousarray(im, np.uint8)
        ch = 1 if im.ndim == 2 else im.shape[2]
        imu = np.ascontiguousarray(imu, np.float64).reshape(-1, 8)
        nc = 0 if cand is None else len(cand)
        cp = np.ascontiguousarray(cand, np.float32).reshape(-1) if nc else None
        if device_detector:
            nc, cp = -1, None
        self.last_rc = capi.check(self.L.rvio_vio_step(self.h, im.reshape(-1), im.shape[1], im.shape[0], im.strides[0], ch,
                                                       imu.ctypes.data, len(imu), cp.ctypes.data if nc > 0 else None, nc,
                                                       1 if cand_filtered else 0, self._pose, C.byref(self._valid)), "rvio_vio_step")
        return self._pose.copy() if self._valid.value else None

    def prefetch(self, im):
        """Announce a frame ahead of its step (System::PushImageData time): a pinned single-channel frame is uploaded beside the
        frame being processed; `step` with the same buffer then skips its own upload (rvio_vio_prefetch)."""
        ch = 1 if im.ndim == 2 else im.shape[2]
        capi.check(self.L.rvio_vio_prefetch(self.h, im.reshape(-1), im.shape[1], im.shape[0], im.strides[0], ch), "rvio_vio_prefetch")

    def prefetch_fence(self) -> int:
        n = C.c_uint64()
        capi.check(self.L.rvio_vio_prefetch_fence(self.h, C.byref(n)), "rvio_vio_prefetch_fence")
        return int(n.value)

    def step_dev(self, img_dev_ptr, pitch, imu, cand_dev_ptr=None, n_cand=0, cand_filtered=False):
        imu = np.ascontiguousarray(imu, np.float64).reshape(-1, 8)
        capi.check(self.L.rvio_vio_step_dev(self.h, img_dev_ptr, pitch, imu.ctypes.data, len(imu), cand_dev_ptr, n_cand,
                                            1 if cand_filtered else 0, self._pose, C.byref(self._valid)), "rvio_vio_step_dev")
        return self._pose.copy() if self._valid.value else None

    def state(self):
        xd, d = C.c_int(), C.c_int()
        capi.check(self.L.rvio_vio_get_state(self.h, None, C.byref(xd), None, C.byref(d)))
        x = np.zeros(xd.value); P = np.zeros(d.value * d.value)
        capi.check(self.L.rvio_vio_get_state(self.h, x.ctypes.data, C.byref(xd), P.ctypes.data, C.byref(d)))
        return x, P.reshape(d.value, d.value).T.copy()

    def set_rank_rule(self, full_information: bool):
        capi.check(self.L.rvio_updater_set_rank_rule(self.L.rvio_vio_updater(self.h), 1 if full_information else 0))

    def graph_launches(self) -> int:
        n = C.c_uint64()
        capi.check(self.L.rvio_vio_graphs(self.h, -1, C.byref(n)))
        return int(n.value)

    def shard_init(self, rank: int, world: int, unique_id: bytes):
        """Feature-sharded single stream (rvio_vio_shard_init): call on every rank before the first frame with the 128-byte id
        rank 0 got from nccl_unique_id()."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        capi.check(self.L.rvio_vio_shard_init(self.h, rank, world, buf), "rvio_vio_shard_init")

    def shard_probe(self, iters: int = 50):
        """us per frame spent in the two collectives of the sharded frame (all-gather of the LK results + all-reduce of the
        normal terms), measured on the pipeline's own stream with nothing else in flight."""
        us = (C.c_float * 2)()
        capi.check(self.L.rvio_vio_shard_probe(self.h, iters, us), "rvio_vio_shard_probe")
        return float(us[0]), float(us[1])

    def update_info(self):
        inf = capi.UpdateInfo()
        capi.check(self.L.rvio_vio_get_update_info(self.h, C.byref(inf)))
        return inf
