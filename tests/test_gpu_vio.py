"""GPU parity of the fused device-resident pipeline (rvio_vio_step) against the oracle's System::MonoVIO loop.

Both sides are free running on the same seeded stream and the same detector output per frame, the oracle in its DEFAULT
rule (the reference's first-small-row cut, Updater.cc:515-524), which the device reproduces (compress.cu).
Tolerance: BASELINE.json's bar is 1e-5 m / 1e-4 rad per frame; measured agreement is ~1e-9 after tens of frames of feedback.
"""
import numpy as np
import pytest

import rvio_b200  # noqa: F401
from rvio_b200 import synth, host
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _quat_angle(q1, q2):
    # angle of q1^-1 * q2 from its vector part (well conditioned near zero, unlike arccos of the dot product)
    x1, y1, z1, w1 = q1
    x2, y2, z2, w2 = q2
    v = np.array([w1 * x2 - x1 * w2 - y1 * z2 + z1 * y2,
                  w1 * y2 + x1 * z2 - y1 * w2 - z1 * x2,
                  w1 * z2 - x1 * y2 + y1 * x2 - z1 * w2])
    return 2 * np.arcsin(min(1.0, float(np.linalg.norm(v))))


def _run(cfg, n_frames, seed, full_info):
    st = synth.Stream(cfg, n_frames, seed, t_static=0.5)
    cache = {}
    cur = [0]

    def det(img, n, s):
        pts = orc.detect_with_subpix(img, n, s, cfg)
        cache[cur[0]] = pts
        return pts

    orc.lib().orc_updater_set_rank_rule(1 if full_info else 0)
    try:
        v = orc.VioOracle(cfg, det)
        consumed = 0
        poses_o, imus, infos = [], [], []
        for i in range(st.n_frames):
            cur[0] = i
            imu, consumed = st.imu_for_frame(i, consumed)
            imus.append(imu)
            poses_o.append(v.step(st.frames[i], imu))
            infos.append(None if v.last_info is None else (v.last_info.n_feat, v.last_info.n_good, v.last_info.rank, v.last_info.rank_full))
        xo, Po = v.state()
    finally:
        orc.lib().orc_updater_set_rank_rule(0)
    g = host.Vio(cfg)
    g.set_rank_rule(full_info)
    poses_g, ginfo = [], []
    for i in range(st.n_frames):
        poses_g.append(g.step(st.frames[i], imus[i], cache.get(i)))
        gi = g.update_info()
        ginfo.append((gi.updated, gi.rank, gi.rank_flags))
    xg, Pg = g.state()
    return poses_o, poses_g, (xo, Po), (xg, Pg), infos, ginfo


def _worst(po, pg):
    n_valid = 0
    worst_p = worst_a = 0.0
    for i, (a, b) in enumerate(zip(po, pg)):
        assert (a is None) == (b is None), f"frame {i}: init state differs"
        if a is None:
            continue
        n_valid += 1
        worst_p = max(worst_p, float(np.abs(a[:3] - b[:3]).max()))
        worst_a = max(worst_a, _quat_angle(a[3:], b[3:]))
    return n_valid, worst_p, worst_a


def test_vio_stream_matches_oracle_config2():
    """120 frames of the config-2 stream against the oracle in the REFERENCE rule; frame 47 is the one where the
    reference's cut discards the '1' features' rows (rank 29 of 38)."""
    cfg = synth.baseline_config(1)
    po, pg, (xo, Po), (xg, Pg), infos, ginfo = _run(cfg, 120, 20260923, full_info=False)
    n_valid, worst_p, worst_a = _worst(po, pg)
    cut_frames = [i for i, inf in enumerate(infos) if inf is not None and inf[1] > 2 and inf[2] < inf[3]]
    dev_cut = [i for i, g in enumerate(ginfo) if g[0] and (g[2] & 1)]
    sweeps = sum(1 for g in ginfo if g[0] and (g[2] & 2))
    print(f"vio config2 (reference rule): {n_valid} poses, worst |dp| = {worst_p:.3e} m, worst angle = {worst_a:.3e} rad; "
          f"cut discards rows on frames {cut_frames} (device: {dev_cut}); {sweeps} frames decided by the Givens sweep")
    assert n_valid >= 95
    assert cut_frames and cut_frames == dev_cut
    for i, inf in enumerate(infos):
        if inf is not None and inf[1] > 2:
            if ginfo[i][2] & 16:      # dependent columns inside the active set: all rows kept, nRank also counts dependent rows
                assert inf[2] == inf[3] and ginfo[i][1] <= inf[2], (i, ginfo[i], inf)
            else:
                assert ginfo[i][1] == inf[2], (i, ginfo[i], inf)      # rows kept == the oracle's nRank
    assert worst_p < 1e-5 and worst_a < 1e-4          # BASELINE.json bar
    assert worst_p < 1e-7 and worst_a < 1e-7          # what float64 on both sides actually gives
    assert xo.shape == xg.shape
    np.testing.assert_allclose(xg, xo, rtol=0, atol=1e-7)
    np.testing.assert_allclose(Pg, Po, rtol=0, atol=1e-7 * np.abs(Po).max())


def test_vio_full_information_mode_and_size_of_the_rank_cut():
    """The other mode (RVIO_RANK_RULE_FULL_INFORMATION) equals the oracle with the cut disabled; the two modes differ from the
    first frame where the reference's cut discards rows by the size of the discarded information (4.3e-3 m on this stream)."""
    cfg = synth.baseline_config(1)
    po_f, pg_f, _, _, _, _ = _run(cfg, 80, 20260923, full_info=True)
    n_valid, worst_p, worst_a = _worst(po_f, pg_f)
    assert n_valid >= 55 and worst_p < 1e-7 and worst_a < 1e-7
    po_r, pg_r, _, _, infos, _ = _run(cfg, 80, 20260923, full_info=False)
    first_cut = next(i for i, inf in enumerate(infos) if inf is not None and inf[1] > 2 and inf[2] < inf[3])
    before = max(float(np.abs(a[:3] - b[:3]).max()) for a, b in zip(pg_f[:first_cut], pg_r[:first_cut]) if a is not None)
    after = max(float(np.abs(a[:3] - b[:3]).max()) for a, b in zip(pg_f[first_cut:], pg_r[first_cut:]) if a is not None)
    print(f"modes agree to {before:.2e} m before frame {first_cut}, differ by up to {after:.2e} m after it")
    assert before < 1e-7
    assert 1e-4 < after < 2e-2


def test_run_asl_tool_matches_oracle_loop(tmp_path):
    """tools/run_asl.py (EuRoC ASL tree in, stamped_pose_ests.dat out: System.cc:369-380) on a small synthetic ASL tree, device
    detector; the file is diffed against the oracle's MonoVIO loop fed by the same reader (detector = the C restatement)."""
    import subprocess
    import sys
    import os
    from rvio_b200 import io_formats
    cfg = synth.baseline_config(0)
    st = synth.Stream(cfg, 44, 31, t_static=0.5)
    io_formats.write_asl(str(tmp_path / "asl"), st.frame_t, st.frames, st.imu)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ycfg = tmp_path / "cfg.yaml"
    ycfg.write_text("%YAML:1.0\nTracker.nFeatures: 150\nTracker.nMaxTrackingLength: 11\n")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "run_asl.py"), str(tmp_path / "asl"), str(tmp_path / "out"),
                        "--config", str(ycfg)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    got = io_formats.read_pose_file(str(tmp_path / "out" / "stamped_pose_ests.dat"))
    cfg2 = synth.Config.from_yaml(str(ycfg))
    o = orc.VioOracle(cfg2, lambda img, k, s: orc.detect_restated(img, k, s, cfg2))
    want = []
    for t, im, imu in io_formats.EurocAslReader(str(tmp_path / "asl"), cfg2.time_offset):
        p = o.step(im, imu)
        if p is not None:
            want.append(np.concatenate([[t], p]))
    want = np.array(want)
    assert got.shape == want.shape and len(got) >= 15, (got.shape, want.shape)
    assert np.array_equal(got[:, 0], want[:, 0])                      # stamps
    assert np.abs(got[:, 1:4] - want[:, 1:4]).max() < 1e-7            # positions
    assert np.abs(np.abs(np.sum(got[:, 4:] * want[:, 4:], axis=1)) - 1).max() < 1e-12     # quaternions (sign-free)


def test_prefetch_gives_identical_poses():
    """rvio_vio_prefetch (frame announced one step ahead, uploaded on the copy stream) must not change a single bit of the
    pose stream; steady frames must actually consume the uploaded copies; an announcement that is not used within two
    steps expires (the buffer may have been recycled by the host)."""
    import torch
    cfg = synth.baseline_config(1)
    st = synth.Stream(cfg, 70, 20260925, t_static=0.5)
    consumed, imus = 0, []
    for i in range(st.n_frames):
        imu, consumed = st.imu_for_frame(i, consumed)
        imus.append(imu)
    keep = [torch.from_numpy(np.ascontiguousarray(f)).pin_memory() for f in st.frames]
    frames = [k.numpy() for k in keep]

    def run(prefetch):
        g = host.Vio(cfg)
        out = []
        if prefetch:
            g.prefetch(frames[0])
        for i in range(st.n_frames):
            if prefetch and i + 1 < st.n_frames:
                g.prefetch(frames[i + 1])
            out.append(g.step(frames[i], imus[i], None, device_detector=True))
        hits = g.prefetch_fence()
        x, P = g.state()
        g.close()
        return out, hits, x, P

    p0, h0, x0, P0 = run(False)
    p1, h1, x1, P1 = run(True)
    assert h0 == 0 and h1 >= st.n_frames - 12, (h0, h1)          # frames before the IMU has two samples are not steps
    assert sum(p is not None for p in p0) > 40
    for a, b in zip(p0, p1):
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(a, b)
    assert np.array_equal(x0, x1) and np.array_equal(P0, P1)

    # expiry: announce a buffer, step three OTHER frames, then hand the announced buffer in with new contents
    g = host.Vio(cfg)
    ref = host.Vio(cfg)
    for i in range(40):
        g.step(frames[i], imus[i], None, device_detector=True)
        ref.step(frames[i], imus[i], None, device_detector=True)
    scratch = torch.from_numpy(frames[5].copy()).pin_memory()
    sbuf = scratch.numpy()
    g.prefetch(sbuf)
    for i in range(40, 43):
        g.step(frames[i], imus[i], None, device_detector=True)
        ref.step(frames[i], imus[i], None, device_detector=True)
    sbuf[:] = frames[43]
    a = g.step(sbuf, imus[43], None, device_detector=True)
    b = ref.step(frames[43], imus[43], None, device_detector=True)
    assert a is not None and np.array_equal(a, b)
    for i in range(44, 50):
        a = g.step(frames[i], imus[i], None, device_detector=True)
        b = ref.step(frames[i], imus[i], None, device_detector=True)
        assert np.array_equal(a, b)
    assert g.prefetch_fence() == 0
    g.close(); ref.close()


def test_programmatic_dependent_launch_gives_identical_poses():
    """rvio_b200_pdl: the short dependent kernels of a frame (CLAHE -> pyramid -> LK -> RANSAC -> per-feature -> normal terms) launched
    as programmatic dependents (resident early, blocked in griddepcontrol.wait until the predecessor completed) must reproduce the
    plain stream order bit for bit, eagerly and as replayed frame graphs."""
    from rvio_b200 import capi
    L = capi.lib()
    cfg = synth.baseline_config(1)
    st = synth.Stream(cfg, 80, 20260923, t_static=0.5)
    consumed, imus = 0, []
    for i in range(st.n_frames):
        imu, consumed = st.imu_for_frame(i, consumed)
        imus.append(imu)

    def run(pdl, graphs):
        prev = L.rvio_b200_pdl(1 if pdl else 0)
        try:
            g = host.Vio(cfg)
            if not graphs:
                capi.check(L.rvio_vio_graphs(g.h, 0, None))
            out = [g.step(st.frames[i], imus[i], None, device_detector=True) for i in range(st.n_frames)]
            n_graph = g.graph_launches()
            x, P = g.state()
            g.close()
        finally:
            L.rvio_b200_pdl(prev)
        return out, x, P, n_graph

    base, x0, P0, _ = run(False, True)
    assert sum(p is not None for p in base) > 50
    for pdl, graphs in ((True, True), (True, False)):
        out, x, P, n_graph = run(pdl, graphs)
        assert (n_graph > 30) == graphs
        for a, b in zip(base, out):
            assert (a is None) == (b is None)
            if a is not None:
                assert np.array_equal(a, b), (pdl, graphs)
        assert np.array_equal(x0, x) and np.array_equal(P0, P)
