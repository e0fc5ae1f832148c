"""CPU: glibc rand() restatement pinned against libc; RANSAC restatement sanity (Ransac.cc:50-266)."""
import ctypes as C

import numpy as np

import rvio_b200  # noqa: F401
from rvio_b200 import synth
from oracle import oracle as orc


def test_glibc_rand_stream_matches_libc():
    libc = C.CDLL("libc.so.6")
    L = orc.lib()
    st = orc.RandState()
    for seed in (1, 12345):
        libc.srand(seed)
        L.orc_rand_seed(C.byref(st), seed)
        ref = [libc.rand() for _ in range(2000)]
        got = [L.orc_rand_next(C.byref(st)) for _ in range(2000)]
        assert ref == got
    libc.srand(1)
    assert [libc.rand() for _ in range(3)] == [1804289383, 846930886, 1681692777]      # never-seeded stream head


def _two_view(n, seed, n_out):
    r = np.random.default_rng(seed)
    cfg = synth.Config()
    T = np.array(cfg.T_BC0).reshape(4, 4)
    Ric = T[:3, :3]
    w = np.array([0.3, -0.2, 0.25]); dt = 0.005; n_imu = 10
    imu = np.zeros((n_imu, 8)); imu[:, 0:3] = w; imu[:, 7] = dt
    ang = np.linalg.norm(w) * dt * n_imu; k = w / np.linalg.norm(w)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R_imu = np.eye(3) - np.sin(ang) * K + (1 - np.cos(ang)) * K @ K           # JPL: I - sin[k]x + (1-cos)[k]x^2
    R = Ric.T @ R_imu @ Ric
    t = np.array([0.04, -0.01, 0.02])
    X1 = np.stack([r.uniform(-2, 2, n), r.uniform(-1.5, 1.5, n), r.uniform(3, 8, n)], 1)
    X2 = (R @ X1.T).T + t
    p1 = np.stack([X1[:, 0] / X1[:, 2], X1[:, 1] / X1[:, 2], np.ones(n)], 1).astype(np.float32).astype(np.float64)
    p2 = np.stack([X2[:, 0] / X2[:, 2], X2[:, 1] / X2[:, 2], np.ones(n)], 1)
    out = r.choice(n, n_out, replace=False)
    p2[out, :2] += r.uniform(0.02, 0.06, (n_out, 2)) * r.choice([-1, 1], (n_out, 2))
    p2 = p2.astype(np.float32).astype(np.float64)
    return cfg, imu, np.ascontiguousarray(p1), np.ascontiguousarray(p2), set(out.tolist())


def test_ransac_finds_planted_outliers():
    cfg, imu, p1, p2, outliers = _two_view(120, 3, 15)
    L = orc.lib()
    rs = orc.RansacState()
    L.orc_ransac_init(C.byref(rs), 1, 1e-5, cfg.small_angle, np.array(cfg.T_BC0))
    flags = np.ones(120, np.uint8); flags[[5, 17]] = 0
    n_in = L.orc_ransac_find_inliers(C.byref(rs), p1, p2, 120, imu, len(imu), flags)
    rejected = set(np.nonzero(flags == 0)[0].tolist()) - {5, 17}
    assert outliers <= rejected                  # every planted outlier is rejected
    assert len(rejected - outliers) <= 6         # and few inliers are lost (float32-rounded coordinates)
    assert n_in == int(flags.sum())
    tp = np.array(rs.two_points).reshape(16, 2)
    assert len(set(tp.ravel().tolist())) == 32   # 16 disjoint pairs
    assert not ({5, 17} & set(tp.ravel().tolist()))


def test_ransac_few_candidates_leaves_flags_untouched():
    cfg, imu, p1, p2, _ = _two_view(40, 4, 5)
    L = orc.lib()
    rs = orc.RansacState()
    L.orc_ransac_init(C.byref(rs), 1, 1e-5, cfg.small_angle, np.array(cfg.T_BC0))
    for n_cand, expect in ((16, 0), (12, 0), (20, -1), (31, -1)):     # <=16: Ransac.cc:201-205 ; 17..31: defined as untouched
        flags = np.zeros(40, np.uint8); flags[:n_cand] = 1
        before = flags.copy()
        rc = L.orc_ransac_find_inliers(C.byref(rs), p1, p2, 40, imu, len(imu), flags)
        assert rc == expect and np.array_equal(flags, before)
