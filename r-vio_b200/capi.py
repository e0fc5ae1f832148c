"""ctypes binding of include/rvio_b200.h (librvio_b200.so).  Used by tests/ and bench.py.

There is no fallback: if the shared library is missing, or no sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RVIO_B200_LIB") or os.path.join(_HERE, "librvio_b200.so")      # (override: the profiling variant, make PHASES=1)

OK, FIRST_IMAGE, NO_FEATURES = 0, 1, 2


class TrackerCfg(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("k1", C.c_float), ("k2", C.c_float), ("p1", C.c_float), ("p2", C.c_float), ("k3", C.c_float),
                ("is_rgb", C.c_int32), ("is_fisheye", C.c_int32), ("enable_equalizer", C.c_int32),
                ("n_features", C.c_int32), ("max_track_len", C.c_int32), ("min_track_len", C.c_int32),
                ("use_sampson", C.c_int32), ("inlier_thr", C.c_double), ("small_angle", C.c_double),
                ("T_BC0", C.c_double * 16)]


class UpdaterCfg(C.Structure):
    _fields_ = [("sigma_px", C.c_float), ("sigma_py", C.c_float), ("T_BC0", C.c_double * 16),
                ("max_clones", C.c_int32), ("max_features", C.c_int32), ("max_track_len", C.c_int32)]


class VioCfg(C.Structure):
    _fields_ = [("tracker", TrackerCfg), ("updater", UpdaterCfg),
                ("imu_rate", C.c_double), ("sigma_g", C.c_double), ("sigma_wg", C.c_double), ("sigma_a", C.c_double),
                ("sigma_wa", C.c_double), ("gravity", C.c_double), ("thr_angle", C.c_double), ("thr_displ", C.c_double),
                ("enable_alignment", C.c_int32), ("min_dist", C.c_float), ("block_x", C.c_int32), ("block_y", C.c_int32),
                ("qual_lvl", C.c_float)]


class UpdateInfo(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n_feat", "n_good", "rows_stacked", "updated",
                                          "n_reject_init", "n_reject_lm", "n_reject_gate", "rank", "rank_flags")]


# every symbol include/rvio_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "rvio_tracker_create", "rvio_tracker_destroy", "rvio_tracker_track", "rvio_tracker_track_dev", "rvio_tracker_track_begin", "rvio_tracker_lk_results", "rvio_tracker_track_finish", "rvio_tracker_detect",
    "rvio_tracker_get_image", "rvio_tracker_n_free", "rvio_tracker_get_tracked_px", "rvio_tracker_seed",
    "rvio_tracker_refill", "rvio_tracker_commit", "rvio_tracker_get_update_count", "rvio_tracker_get_update_lists",
    "rvio_tracker_get_debug", "rvio_tracker_get_ransac_debug", "rvio_tracker_get_pyramid",
    "rvio_updater_create", "rvio_updater_destroy", "rvio_updater_update", "rvio_updater_update_from_tracker",
    "rvio_updater_get_debug", "rvio_updater_get_normal_terms", "rvio_updater_update_begin",
    "rvio_updater_reduce_buffer", "rvio_updater_update_finish", "rvio_updater_set_rank_rule",
    "rvio_vio_create", "rvio_vio_destroy", "rvio_vio_step", "rvio_vio_step_dev", "rvio_vio_prefetch", "rvio_vio_prefetch_fence", "rvio_vio_get_state",
    "rvio_vio_get_update_info", "rvio_vio_shard_init", "rvio_vio_shard_probe", "rvio_b200_nccl_unique_id", "rvio_vio_tracker", "rvio_vio_updater", "rvio_vio_timeline", "rvio_vio_graphs",
    "rvio_b200_version", "rvio_b200_last_error", "rvio_b200_kernel_launches", "rvio_b200_pdl",
    "rvio_tracker_stream", "rvio_updater_stream", "rvio_b200_profile", "rvio_b200_profile_report",
]

_lib = None


def _p(dtype):
    return np.ctypeslib.ndpointer(dtype, flags="C_CONTIGUOUS")


def lib():
    """Loads librvio_b200.so (raises OSError when it has not been built: there is no other path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(f"{LIB_PATH} not built -- run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp, ci, pi = C.c_void_p, C.c_int, C.POINTER(C.c_int)
    u8, f32, f64, i32 = _p(np.uint8), _p(np.float32), _p(np.float64), _p(np.int32)
    L.rvio_tracker_create.argtypes = [C.POINTER(TrackerCfg), ci, C.POINTER(vp)]
    L.rvio_tracker_destroy.argtypes = [vp]
    L.rvio_tracker_destroy.restype = None
    L.rvio_tracker_track.argtypes = [vp, u8, ci, ci, ci, ci, vp, ci]
    L.rvio_tracker_track_dev.argtypes = [vp, vp, ci, vp, ci]
    L.rvio_tracker_track_begin.argtypes = [vp, u8, ci, ci, ci, ci, vp, ci, ci, ci]
    L.rvio_tracker_lk_results.argtypes = [vp, ci, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), pi]
    L.rvio_tracker_track_finish.argtypes = [vp]
    L.rvio_tracker_detect.argtypes = [vp, ci, C.c_float, C.c_float, f32, pi]
    L.rvio_tracker_get_image.argtypes = [vp, u8, ci]
    L.rvio_tracker_n_free.argtypes = [vp, pi]
    L.rvio_tracker_get_tracked_px.argtypes = [vp, f32, pi]
    L.rvio_tracker_seed.argtypes = [vp, f32, ci]
    L.rvio_tracker_refill.argtypes = [vp, f32, ci, pi]
    L.rvio_tracker_commit.argtypes = [vp]
    L.rvio_tracker_get_update_count.argtypes = [vp, pi, pi]
    L.rvio_tracker_get_update_lists.argtypes = [vp, u8, i32, f32]
    L.rvio_tracker_get_debug.argtypes = [vp, pi, u8, u8, f32, f32, i32]
    L.rvio_tracker_get_ransac_debug.argtypes = [vp, i32, i32, pi, pi, f64]
    L.rvio_tracker_get_pyramid.argtypes = [vp, ci, ci, vp, pi, pi]
    L.rvio_updater_create.argtypes = [C.POINTER(UpdaterCfg), ci, C.POINTER(vp)]
    L.rvio_updater_destroy.argtypes = [vp]
    L.rvio_updater_destroy.restype = None
    L.rvio_updater_update.argtypes = [vp, f64, ci, f64, ci, u8, i32, f32, ci, f64, f64, C.POINTER(UpdateInfo)]
    L.rvio_updater_update_from_tracker.argtypes = [vp, vp, f64, ci, f64, ci, f64, f64, C.POINTER(UpdateInfo)]
    L.rvio_updater_get_debug.argtypes = [vp, ci, u8, f64, f64, i32]
    L.rvio_updater_get_normal_terms.argtypes = [vp, f64, f64, ci]
    L.rvio_updater_update_begin.argtypes = [vp, f64, ci, f64, ci, u8, i32, f32, ci, ci, ci]
    L.rvio_updater_reduce_buffer.argtypes = [vp, C.POINTER(vp), pi]
    L.rvio_updater_update_finish.argtypes = [vp, f64, f64, C.POINTER(UpdateInfo)]
    L.rvio_updater_set_rank_rule.argtypes = [vp, ci]
    L.rvio_vio_create.argtypes = [C.POINTER(VioCfg), ci, C.POINTER(vp)]
    L.rvio_vio_destroy.argtypes = [vp]
    L.rvio_vio_destroy.restype = None
    L.rvio_vio_step.argtypes = [vp, u8, ci, ci, ci, ci, vp, ci, vp, ci, ci, f64, pi]
    L.rvio_vio_step_dev.argtypes = [vp, vp, ci, vp, ci, vp, ci, ci, f64, pi]
    L.rvio_vio_prefetch.argtypes = [vp, u8, ci, ci, ci, ci]
    L.rvio_vio_prefetch_fence.argtypes = [vp, vp]
    L.rvio_vio_get_state.argtypes = [vp, vp, pi, vp, pi]
    L.rvio_vio_get_update_info.argtypes = [vp, C.POINTER(UpdateInfo)]
    L.rvio_vio_shard_init.argtypes = [vp, ci, ci, vp]
    L.rvio_b200_nccl_unique_id.argtypes = [vp]
    L.rvio_vio_shard_probe.argtypes = [vp, ci, vp]
    L.rvio_vio_timeline.argtypes = [vp, ci, vp]
    L.rvio_vio_graphs.argtypes = [vp, ci, vp]
    L.rvio_vio_tracker.argtypes = [vp]
    L.rvio_vio_tracker.restype = vp
    L.rvio_vio_updater.argtypes = [vp]
    L.rvio_vio_updater.restype = vp
    L.rvio_b200_version.restype = C.c_char_p
    L.rvio_b200_last_error.restype = C.c_char_p
    L.rvio_b200_kernel_launches.restype = C.c_uint64
    L.rvio_b200_pdl.argtypes = [ci]
    L.rvio_b200_profile.argtypes = [ci]
    L.rvio_b200_profile.restype = None
    L.rvio_b200_profile_report.argtypes = [C.c_char_p, ci]
    L.rvio_tracker_stream.argtypes = [vp]
    L.rvio_tracker_stream.restype = vp
    L.rvio_updater_stream.argtypes = [vp]
    L.rvio_updater_stream.restype = vp
    _lib = L
    return L


class RvioError(RuntimeError):
    pass


def check(rc, what=""):
    if rc < 0:
        raise RvioError(f"{what} failed rc={rc}: {lib().rvio_b200_last_error().decode()}")
    return rc


def tracker_cfg(cfg) -> TrackerCfg:
    t = TrackerCfg()
    t.width, t.height = cfg.width, cfg.height
    for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3"):
        setattr(t, k, float(np.float32(getattr(cfg, k))))
    t.is_rgb, t.is_fisheye, t.enable_equalizer = 0, int(getattr(cfg, "fisheye", 0)), cfg.enable_equalizer
    t.n_features, t.max_track_len, t.min_track_len = cfg.n_features, cfg.max_track_len, cfg.min_track_len
    t.use_sampson, t.inlier_thr, t.small_angle = cfg.use_sampson, cfg.inlier_thr, cfg.small_angle
    t.T_BC0 = (C.c_double * 16)(*cfg.T_BC0)
    return t


def updater_cfg(cfg) -> UpdaterCfg:
    u = UpdaterCfg()
    u.sigma_px, u.sigma_py = float(np.float32(cfg.sigma_px)), float(np.float32(cfg.sigma_py))
    u.T_BC0 = (C.c_double * 16)(*cfg.T_BC0)
    u.max_clones = cfg.max_track_len - 1
    u.max_features = (cfg.n_features + 1) // 2
    u.max_track_len = cfg.max_track_len
    return u


def vio_cfg(cfg) -> VioCfg:
    v = VioCfg()
    v.tracker = tracker_cfg(cfg)
    v.updater = updater_cfg(cfg)
    v.imu_rate, v.sigma_g, v.sigma_wg, v.sigma_a, v.sigma_wa = cfg.imu_rate, cfg.sigma_g, cfg.sigma_wg, cfg.sigma_a, cfg.sigma_wa
    v.gravity, v.thr_angle, v.thr_displ, v.enable_alignment = cfg.gravity, cfg.thr_angle, cfg.thr_displ, cfg.enable_alignment
    v.min_dist, v.block_x, v.block_y = float(np.float32(cfg.min_dist)), cfg.block_x, cfg.block_y
    v.qual_lvl = float(np.float32(cfg.qual_lvl))
    return v
