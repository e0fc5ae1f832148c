"""GPU parity: CUDA MSCKF update (through the C ABI) vs the CPU oracle (float64).

Tolerances (float64 device arithmetic, different but algebraically identical factorisations):
  per-feature inverse depth 1e-9, Mahalanobis distance 1e-8 relative, normal terms 1e-9 relative,
  state 1e-9 absolute (bar in BASELINE.json: 1e-5 m / 1e-4 rad), covariance 1e-9 relative to max|P|.
The reference's first-small-row rank cut (Updater.cc:515-524) can discard informative rows of the compressed system.
The device reproduces it (compress.cu: certificate on the normal terms, or a replay of the reference's Givens sweep);
every case here is compared against the oracle in its DEFAULT (reference) rule, the kept-row count included, and the
frames where the cut discards information are counted.  The full-information mode is compared against the oracle with
the cut disabled.
"""
import numpy as np
import pytest

import rvio_b200  # noqa: F401
from rvio_b200 import synth, host
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _collect_cases(cfg, n_frames, seed):
    st = synth.Stream(cfg, n_frames, seed, t_static=0.5)
    v = orc.VioOracle(cfg)
    consumed = 0
    cases = []
    for i in range(st.n_frames):
        imu, consumed = st.imu_for_frame(i, consumed)
        p = v.step(st.frames[i], imu)
        if p is not None and v.last_info is not None and getattr(v, "last_update_in", None) is not None:
            x, Pc, types, off, xy = v.last_update_in
            cases.append((x.copy(), Pc.copy(), types.copy(), off.copy(), xy.copy()))
            v.last_update_in = None
    return cases


def _check_case(cfg, upd, x, Pc, types, off, xy, stats):
    d = int(round(np.sqrt(len(Pc))))
    P = Pc.reshape(d, d).T.copy()
    n = d - 24
    L = orc.lib()
    L.orc_updater_set_rank_rule(0)
    xo, Po, info, dbg = orc.updater_update(cfg, x, P, types, off, xy, debug=True)          # the reference's rule
    cut = info.updated and info.rank < info.rank_full
    xg, Pg = upd.update(x, P, types, (off, xy))
    gi = upd.info
    if info.updated:
        if gi.rank_flags & 16:       # dependent columns inside the active set: every row kept; the reference's nRank also counts dependent rows
            assert not cut and gi.rank <= info.rank
        else:
            assert gi.rank == info.rank, (gi.rank, info.rank, info.rank_full, gi.rank_flags)
        assert bool(gi.rank_flags & 1) == bool(cut), (gi.rank_flags, info.rank, info.rank_full)
        stats["by_sweep"] = stats.get("by_sweep", 0) + int(bool(gi.rank_flags & 2))
    if cut:
        stats["rank_cut_frames"] += 1
        # the other mode: every row kept == the oracle with the cut disabled
        L.orc_updater_set_rank_rule(1)
        xf, Pf, _, _ = orc.updater_update(cfg, x, P, types, off, xy, debug=True)
        L.orc_updater_set_rank_rule(0)
        upd.set_rank_rule(True)
        xg2, Pg2 = upd.update(x, P, types, (off, xy))
        upd.set_rank_rule(False)
        np.testing.assert_allclose(xg2, xf, rtol=0, atol=1e-9)
        np.testing.assert_allclose(Pg2, Pf, rtol=0, atol=1e-9 * np.abs(Pf).max())
        stats["cut_effect"] = max(stats.get("cut_effect", 0.0), float(np.abs(xf - xo).max()))
    nf = len(types)
    assert gi.n_feat == nf
    if nf:
        gd = upd.debug(nf)
        assert np.array_equal(gd["status"], dbg["status"]), (gd["status"], dbg["status"])
        ok = dbg["status"] != 1
        np.testing.assert_allclose(gd["pfinv"][ok], dbg["pfinv"][ok], rtol=0, atol=1e-9)
        gok = dbg["status"] != 1
        gok &= dbg["status"] != 2
        np.testing.assert_allclose(gd["gamma"][gok], dbg["gamma"][gok], rtol=1e-8, atol=1e-10)
        assert gi.n_good == info.n_good and gi.rows_stacked == info.rows_stacked
        assert (gi.n_reject_init, gi.n_reject_lm, gi.n_reject_gate) == (info.n_reject_init, info.n_reject_lm, info.n_reject_gate)
        if info.rows_stacked > 0 and not (gi.rank_flags & 8):
            G, z = upd.normal_terms(n)
            Go = dbg["H"].T @ dbg["H"]; zo = dbg["H"].T @ dbg["r"]
            np.testing.assert_allclose(G, Go, rtol=0, atol=1e-9 * max(1.0, np.abs(Go).max()))
            np.testing.assert_allclose(z, zo, rtol=0, atol=1e-9 * max(1.0, np.abs(zo).max()))
    assert gi.updated == info.updated
    np.testing.assert_allclose(xg, xo, rtol=0, atol=1e-9)
    np.testing.assert_allclose(Pg, Po, rtol=0, atol=1e-9 * np.abs(Po).max())
    assert np.array_equal(Pg, Pg.T)
    stats["cases"] += 1
    stats["updated"] += int(info.updated)
    stats["max_dx"] = max(stats["max_dx"], float(np.abs(xo - x).max()))
    stats["max_err_x"] = max(stats["max_err_x"], float(np.abs(xg - xo).max()))


def test_updater_stream_config2():
    cfg = synth.baseline_config(1)
    cases = _collect_cases(cfg, 100, 20260923)
    assert len(cases) >= 40
    upd = host.Updater(cfg)
    stats = dict(cases=0, updated=0, rank_cut_frames=0, max_dx=0.0, max_err_x=0.0)
    for c in cases:
        _check_case(cfg, upd, *c, stats)
    print("updater config2:", stats)
    assert stats["updated"] >= 30
    assert stats["rank_cut_frames"] >= 1                  # frame 47 of this stream: '2' and '1' supports are disjoint


def test_updater_stream_config1():
    cfg = synth.baseline_config(0)
    cases = _collect_cases(cfg, 45, 20260922)
    upd = host.Updater(cfg)
    stats = dict(cases=0, updated=0, rank_cut_frames=0, max_dx=0.0, max_err_x=0.0)
    for c in cases:
        _check_case(cfg, upd, *c, stats)
    print("updater config1:", stats)
    assert stats["updated"] >= 10


def test_updater_passthrough_and_empty():
    cfg = synth.baseline_config(1)
    upd = host.Updater(cfg)
    N = 5
    x = np.zeros(26 + 7 * N); x[3] = 1; x[13] = 1; x[9] = 1
    for c in range(N):
        x[26 + 7 * c + 3] = 1
    d = 24 + 6 * N
    P = np.eye(d) * 1e-4
    xg, Pg = upd.update(x, P, np.zeros(0, np.uint8), (np.zeros(1, np.int32), np.zeros((0, 2), np.float32)))
    assert upd.info.updated == 0
    assert np.array_equal(xg, x) and np.array_equal(Pg, P)


@pytest.mark.parametrize("idx,n_feat,n_clones", [(1, 100, None), (2, 96, None), (4, 64, None), (4, 40, 17), (2, 50, 13), (2, 50, 14)])
def test_updater_worstcase_shapes(idx, n_feat, n_clones):
    """SURVEY 8d micro-benchmark shapes: maximum-length tracks, 50/50 type mix; N = 11 / 25 / 30 exercise both the
    single-CTA solve (N <= 13) and the general GEMM + Gauss-Jordan path."""
    cfg = synth.baseline_config(idx)
    x, P, types, off, xy = synth.make_update_case(cfg, n_feat, 100 + idx, n_clones=n_clones)
    upd = host.Updater(cfg)
    stats = dict(cases=0, updated=0, rank_cut_frames=0, max_dx=0.0, max_err_x=0.0)
    d = P.shape[0]
    _check_case(cfg, upd, x, np.ascontiguousarray(P.T).reshape(-1), types, off, xy, stats)   # column-major bytes, as the stream cases
    print(f"config[{idx}] N={(len(x) - 26) // 7} feats={n_feat}:", stats)
    assert stats["updated"] == 1


def test_updater_feature_sharding_matches_unsharded():
    """Multi-GPU form on one device: each 'rank' forms the normal terms of its share (f % world == rank); the shares are
    summed (what ncclAllReduce does across GPUs) and the finish step must reproduce the unsharded update."""
    import ctypes as C
    import torch
    from rvio_b200 import capi
    cfg = synth.baseline_config(1)
    x, P, types, off, xy = synth.make_update_case(cfg, 90, 7)
    upd = host.Updater(cfg)
    xr, Pr = upd.update(x, P, types, (off, xy))
    L = capi.lib()
    d = P.shape[0]; n = d - 24
    Pc = np.ascontiguousarray(P.T)
    world = 4
    total = None
    count = n * n + n + 8 + n + 1          # [G | z | counters | per-class information]
    for rank in range(world):
        capi.check(L.rvio_updater_update_begin(upd.h, x, len(x), Pc, d, types, off, np.ascontiguousarray(xy).reshape(-1), len(types), rank, world))
        ptr, cnt = C.c_void_p(), C.c_int()
        capi.check(L.rvio_updater_reduce_buffer(upd.h, C.byref(ptr), C.byref(cnt)))
        assert cnt.value == count
        stream = torch.cuda.ExternalStream(L.rvio_updater_stream(upd.h))
        stream.synchronize()
        buf = torch.empty(count, dtype=torch.float64, device="cuda")
        C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(buf.data_ptr()), ptr, C.c_size_t(8 * count), 3)
        total = buf.clone() if total is None else total + buf
    # write the reduced terms back (stand-in for the all-reduce result) and finish
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(ptr, C.c_void_p(total.data_ptr()), C.c_size_t(8 * count), 3)
    torch.cuda.synchronize()
    xo = np.empty_like(x); Po = np.empty_like(Pc); info = capi.UpdateInfo()
    capi.check(L.rvio_updater_update_finish(upd.h, xo, Po, C.byref(info)))
    assert info.n_good == upd.info.n_good and info.updated == 1
    np.testing.assert_allclose(xo, xr, rtol=0, atol=1e-10)
    np.testing.assert_allclose(Po.T, Pr, rtol=0, atol=1e-10 * np.abs(Pr).max())
