"""GPU parity: CUDA tracker (through the C ABI) vs the CPU oracle on a seeded synthetic stream.
Bar: bit-exact (integer / byte / float32 bit patterns / index sets)."""
import numpy as np
import pytest

import rvio_b200  # noqa: F401
from rvio_b200 import synth, host, capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _run_stream(cfg, n_frames, seed, start, check_pyr=True):
    st = synth.Stream(cfg, n_frames, seed, t_static=0.25)
    det = lambda img, n, s: orc.detect_with_subpix(img, n, s, cfg)
    o = orc.Tracker(cfg, det)
    g = host.Tracker(cfg, 0, det)
    consumed = 0
    stats = dict(frames=0, lost=0, ransac_rej=0, emitted=0, type2=0)
    for i in range(n_frames):
        imu, consumed = st.imu_for_frame(i, consumed)
        if i < start or len(imu) < 2:
            continue
        img = st.frames[i]
        # --- oracle frame (detector on its own equalised image)
        rc_o = o.track(img, imu)
        # --- CUDA frame: same detections are recomputed from ITS equalised image (must be identical bytes)
        rc_g = g.track(img, imu)
        assert rc_g == (capi.FIRST_IMAGE if rc_o == 1 else rc_o), (i, rc_o, rc_g)
        eq_o = orc.clahe(img)
        # after commit the processed image is the "previous" pyramid (which=1)
        eq_g = g.pyramid(1, 0)
        assert np.array_equal(eq_o, eq_g), f"frame {i}: CLAHE differs in {(eq_o != eq_g).sum()} px"
        if check_pyr:
            lv = eq_o
            for l in range(1, 4):
                lv = orc.pyr_down(lv)
                assert np.array_equal(lv, g.pyramid(1, l)), f"frame {i}: pyramid level {l} differs"
        if rc_o == 1:
            continue
        do, dg = o.debug(), g.debug()
        assert do["n"] == dg["n"]
        assert np.array_equal(do["status"], dg["status"]), f"frame {i}: LK status differs at {np.nonzero(do['status'] != dg['status'])[0]}"
        assert np.array_equal(_bits(do["lk"]), _bits(dg["lk"])), f"frame {i}: LK coordinates differ ({(_bits(do['lk']) != _bits(dg['lk'])).any(1).sum()} pts)"
        assert np.array_equal(_bits(do["un"]), _bits(dg["un"])), f"frame {i}: undistorted coordinates differ"
        assert np.array_equal(do["flags"], dg["flags"]), f"frame {i}: inlier flags differ at {np.nonzero(do['flags'] != dg['flags'])[0]}"
        rs = orc.lib().orc_tracker_ransac(o.h).contents
        rg = g.ransac_debug()
        if rg["n_cand"] >= 32:
            assert list(rs.two_points) == list(rg["two_points"]), f"frame {i}: sampled pairs differ (glibc rand stream)"
            assert list(rs.n_inliers) == list(rg["n_inliers"]), f"frame {i}: hypothesis votes differ"
            assert rs.winner == rg["winner"]
            np.testing.assert_allclose(np.array(rs.hyp).reshape(16, 3, 3), rg["hyp"], rtol=0, atol=1e-13)
        to, oo, xo = o.update_lists()
        tg, og, xg = g.update_lists()
        assert np.array_equal(to, tg) and np.array_equal(oo, og), f"frame {i}: update lists differ"
        assert np.array_equal(_bits(xo), _bits(xg)), f"frame {i}: update measurements differ"
        stats["frames"] += 1
        stats["lost"] += int((do["status"] == 0).sum())
        stats["ransac_rej"] += int(((do["status"] != 0) & (do["flags"] == 0)).sum())
        stats["emitted"] += len(to)
        stats["type2"] += int((to == ord('2')).sum())
    return stats


def test_tracker_stream_config2():
    cfg = synth.baseline_config(1)            # 752x480, 200 feats, 11-clone window
    s = _run_stream(cfg, 45, 20260923, 0)
    print("config2 stream stats:", s)
    assert s["frames"] >= 40 and s["emitted"] > 0 and s["type2"] > 0


def test_tracker_stream_config1_short_tracks():
    cfg = synth.baseline_config(0)            # 150 feats, 10-clone window
    cfg.max_track_len = 6                     # exercise the type-'2' path and trimming often
    cfg.inlier_thr = 3e-9                     # clean synthetic data: tighten the gate so RANSAC rejects tracks
    s = _run_stream(cfg, 30, 20260922, 0, check_pyr=False)
    print("config1 stream stats:", s)
    assert s["type2"] > 0 and s["ransac_rej"] > 0


def test_tracker_stress_1280x720():
    cfg = synth.baseline_config(2)            # 1280x720, 600 feats
    cfg.max_track_len = 8
    s = _run_stream(cfg, 14, 20260924, 0, check_pyr=True)
    print("config3 stream stats:", s)
    assert s["frames"] >= 10


def test_tracker_stream_fisheye_model():
    """Camera.Fisheye: 1 (Tracker.cc:119): same stream, the equidistant undistortion in every stage that normalises pixels
    (LK epilogue, seeding, refill) -- undistorted coordinates, RANSAC votes and update lists stay bit-identical."""
    cfg = synth.baseline_config(1)
    cfg.fisheye = 1
    cfg.k1, cfg.k2, cfg.p1, cfg.p2, cfg.k3 = -0.0127, 0.0154, -0.0201, 0.0072, 0.0
    s = _run_stream(cfg, 30, 20260925, 0, check_pyr=False)
    print("fisheye stream stats:", s)
    assert s["frames"] >= 25 and s["emitted"] > 0


def test_tracker_feature_sharding_matches_unsharded():
    """Multi-GPU form on one device: two tracker handles play ranks 0 and 1 of a 2-way feature-sharded stream
    (LK for half of the feature indices each), the per-feature LK arrays are exchanged slice by slice (what the NCCL
    all-gather does across GPUs, tests/dist_sharded_tracker.py), RANSAC + bookkeeping run on both: every observable of both
    ranks must equal the unsharded tracker's, bit for bit."""
    import ctypes as C
    cfg = synth.baseline_config(1)
    n_frames = 20
    st = synth.Stream(cfg, n_frames, 20260927, t_static=0.25)
    det = lambda img, n, s: orc.detect_with_subpix(img, n, s, cfg)
    ref = host.Tracker(cfg, 0, det)
    ranks = [host.Tracker(cfg, 0, det) for _ in range(2)]
    L = capi.lib()
    cudart = C.cdll.LoadLibrary("libcudart.so.12")
    world = 2
    consumed = 0
    checked = 0
    for i in range(n_frames):
        imu, consumed = st.imu_for_frame(i, consumed)
        if len(imu) < 2:
            continue
        im = st.frames[i]
        rc_ref = ref.track(im, imu)
        imu_c = np.ascontiguousarray(imu, np.float64).reshape(-1, 8)
        rcs, ptrs = [], []
        for r, t in enumerate(ranks):
            rc = capi.check(L.rvio_tracker_track_begin(t.h, im.reshape(-1), im.shape[1], im.shape[0], im.strides[0], 1,
                                                       imu_c.ctypes.data, len(imu_c), r, world))
            rcs.append(rc)
            lk, un, stt, S = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
            capi.check(L.rvio_tracker_lk_results(t.h, world, C.byref(lk), C.byref(un), C.byref(stt), C.byref(S)))
            ptrs.append((lk.value, un.value, stt.value, S.value))
        assert rcs[0] == rcs[1] == (capi.FIRST_IMAGE if rc_ref == 1 else rc_ref)
        if rcs[0] == capi.OK:
            S = ptrs[0][3]
            assert S == host.shard_range(cfg.n_features, 0, world)[2]
            for src in range(world):                       # rank src's slice -> the other rank
                dst = 1 - src
                for k, item in ((0, 8), (1, 8), (2, 1)):
                    cudart.cudaMemcpy(C.c_void_p(ptrs[dst][k] + src * S * item), C.c_void_p(ptrs[src][k] + src * S * item),
                                      C.c_size_t(S * item), 3)
            for t in ranks:
                capi.check(L.rvio_tracker_track_finish(t.h))
        # the remaining host flow (seed / refill / commit) is the ordinary one, replicated
        for t in ranks:
            if rcs[0] == capi.FIRST_IMAGE:
                pts = np.ascontiguousarray(det(t.equalized_image(), cfg.n_features, 1), np.float32).reshape(-1, 2)
                capi.check(L.rvio_tracker_seed(t.h, pts, len(pts)))
            elif rcs[0] == capi.OK and t.n_free() > 0:
                newer = host.find_newer(cfg, det(t.equalized_image(), cfg.n_features, 2), t.tracked_px())
                if len(newer):
                    used = C.c_int()
                    capi.check(L.rvio_tracker_refill(t.h, np.ascontiguousarray(newer), len(newer), C.byref(used)))
            if rcs[0] != capi.NO_FEATURES:
                capi.check(L.rvio_tracker_commit(t.h))
        if rcs[0] == capi.OK:
            d0 = ref.debug()
            t0, o0, x0 = ref.update_lists()
            for t in ranks:
                d1 = t.debug()
                assert np.array_equal(d0["status"], d1["status"]) and np.array_equal(d0["flags"], d1["flags"]), i
                assert np.array_equal(_bits(d0["lk"]), _bits(d1["lk"])) and np.array_equal(_bits(d0["un"]), _bits(d1["un"])), i
                t1, o1, x1 = t.update_lists()
                assert np.array_equal(t0, t1) and np.array_equal(o0, o1) and np.array_equal(_bits(x0), _bits(x1)), i
            checked += 1
    assert checked >= 12

