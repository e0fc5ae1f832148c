"""CPU: the C restatement of Updater::update (oracle/updater.c) cross-checked against the independent NumPy/LAPACK
restatement (oracle/np_updater.py, different factorisations), Numerics helpers, chi^2 table."""
import numpy as np
import pytest

import rvio_b200  # noqa: F401
from rvio_b200 import synth
from oracle import oracle as orc, np_updater as npu


def test_numerics_helpers():
    L = orc.lib()
    r = np.random.default_rng(0)
    for _ in range(50):
        q1 = r.standard_normal(4); q1 /= np.linalg.norm(q1)
        q2 = r.standard_normal(4); q2 /= np.linalg.norm(q2)
        o = np.zeros(4); L.orc_quat_mul(q1, q2, o)
        np.testing.assert_allclose(o, npu.quat_mul(q1, q2), atol=1e-15)
        R = np.zeros(9); L.orc_quat_to_rot(q1, R)
        np.testing.assert_allclose(R.reshape(3, 3), npu.quat_to_rot(q1), atol=1e-15)
        np.testing.assert_allclose(R.reshape(3, 3) @ R.reshape(3, 3).T, np.eye(3), atol=1e-13)
        qb = np.zeros(4); L.orc_rot_to_quat(R, qb)
        np.testing.assert_allclose(qb, q1 if q1[3] >= 0 else -q1, atol=1e-12)
        # JPL convention: R(q1 (x) q2) = R(q1) R(q2)
        np.testing.assert_allclose(npu.quat_to_rot(npu.quat_mul(q1, q2)), npu.quat_to_rot(q1) @ npu.quat_to_rot(q2), atol=1e-13)


def test_chi2_table():
    from scipy.stats import chi2
    L = orc.lib()
    for dof in (1, 2, 9, 21, 59, 200, 500):
        assert abs(L.orc_chi2_95(dof) - chi2.ppf(0.95, dof)) < 6e-7
    assert L.orc_chi2_95(1) == 3.841459          # Numerics.h:174 first entry


@pytest.fixture(scope="module")
def cases():
    cfg = synth.baseline_config(0)
    cfg.max_track_len = 8
    st = synth.Stream(cfg, 48, 11, t_static=0.25)
    v = orc.VioOracle(cfg)
    consumed, out = 0, []
    for i in range(st.n_frames):
        imu, consumed = st.imu_for_frame(i, consumed)
        p = v.step(st.frames[i], imu)
        if p is not None and getattr(v, "last_update_in", None) is not None:
            out.append(tuple(a.copy() for a in v.last_update_in))
            v.last_update_in = None
    assert len(out) >= 10
    return cfg, out


def test_c_vs_numpy_updater(cases):
    cfg, cs = cases
    sig = float(max(np.float32(cfg.sigma_px), np.float32(cfg.sigma_py)))
    L = orc.lib()
    n_upd = n_cut = 0
    for (x, Pc, types, off, xy) in cs:
        d = int(round(np.sqrt(len(Pc)))); P = Pc.reshape(d, d).T.copy()
        xo, Po, info, dbg = orc.updater_update(cfg, x, P, types, off, xy, debug=True)
        if info.updated and info.rank < info.rank_full:       # reference's first-small-row cut dropped rows: compare full-info
            n_cut += 1
            L.orc_updater_set_rank_rule(1)
            try:
                xo, Po, _, _ = orc.updater_update(cfg, x, P, types, off, xy, debug=True)
            finally:
                L.orc_updater_set_rank_rule(0)
        xn, Pn, inf = npu.update(sig, cfg.T_BC0, x, P, types, off, xy)
        assert np.array_equal(dbg["status"], inf["status"])
        ok = dbg["status"] == 0
        ok3 = (dbg["status"] == 0) | (dbg["status"] == 3)
        np.testing.assert_allclose(dbg["gamma"][ok3], inf["gamma"][ok3], rtol=1e-9)
        np.testing.assert_allclose(dbg["pfinv"][ok], inf["pfinv"][ok], atol=1e-10)
        G = dbg["H"].T @ dbg["H"]
        np.testing.assert_allclose(G, inf["G"], atol=1e-9 * max(1, np.abs(G).max()))
        np.testing.assert_allclose(xo, xn, atol=1e-9)
        np.testing.assert_allclose(Po, Pn, atol=1e-9 * np.abs(Po).max())
        n_upd += int(info.updated)
    assert n_upd >= 5
    print("updates", n_upd, "rank-cut frames", n_cut)


def test_update_passthrough_with_two_features(cases):
    cfg, cs = cases
    x, Pc, types, off, xy = max(cs, key=lambda c: len(c[2]))
    d = int(round(np.sqrt(len(Pc)))); P = Pc.reshape(d, d).T.copy()
    xo, Po, info = orc.updater_update(cfg, x, P, types[:2], off[:3], xy[:off[2]])
    assert info.updated == 0 and np.array_equal(xo, x) and np.array_equal(Po, P)     # Updater.cc:460,621-627
