"""Independent NumPy/LAPACK float64 restatement of Updater::update (src/rvio/Updater.cc:72-628).

TEST INFRASTRUCTURE ONLY.  Purpose: cross-check oracle/updater.c (Eigen is absent, so the C restatement cannot
be pinned against the reference binary).  This version deliberately uses DIFFERENT numerical routes:
  * left-nullspace projection by a Householder QR of H_f (numpy.linalg.qr) instead of Givens sweeps,
  * chi^2 gate through numpy.linalg.solve,
  * compression + EKF through the normal terms G = H^T H, z = H^T r:
        dx = P[:,c] (G Pcc + s^2 I)^-1 z ,  P+ = P - P[:,c] (G Pcc + s^2 I)^-1 G P[c,:]
    which is algebraically identical to Givens-QR compression + K = P H^T S^-1 + Joseph form (the update depends on
    H, r only through H^T H and H^T r when the noise is s^2 I).  This is also the algorithm the CUDA path uses.
Agreement with oracle/updater.c to ~1e-9 is asserted in tests/test_oracle_updater.py.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.stats import chi2 as _chi2


def quat_mul(q1, q2):     # Numerics.h:30-63
    x1, y1, z1, w1 = q1
    L = np.array([[w1, z1, -y1, x1], [-z1, w1, x1, y1], [y1, -x1, w1, z1], [-x1, -y1, -z1, w1]])
    q = L @ q2
    q = q / np.linalg.norm(q)
    return -q if q[3] < 0 else q


def skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])


def quat_to_rot(q):       # Numerics.h:111-120
    qx = skew(q[:3])
    return np.eye(3) - 2 * q[3] * qx + 2 * qx @ qx


def rot_to_quat(R):       # Numerics.h:126-167
    T = np.trace(R)
    q = np.zeros(4)
    if R[0, 0] > T and R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        q[0] = math.sqrt((1 + 2 * R[0, 0] - T) / 4)
        q[1] = (R[0, 1] + R[1, 0]) / (4 * q[0]); q[2] = (R[0, 2] + R[2, 0]) / (4 * q[0]); q[3] = (R[1, 2] - R[2, 1]) / (4 * q[0])
    elif R[1, 1] > T and R[1, 1] > R[0, 0] and R[1, 1] > R[2, 2]:
        q[1] = math.sqrt((1 + 2 * R[1, 1] - T) / 4)
        q[0] = (R[0, 1] + R[1, 0]) / (4 * q[1]); q[2] = (R[1, 2] + R[2, 1]) / (4 * q[1]); q[3] = (R[2, 0] - R[0, 2]) / (4 * q[1])
    elif R[2, 2] > T and R[2, 2] > R[0, 0] and R[2, 2] > R[1, 1]:
        q[2] = math.sqrt((1 + 2 * R[2, 2] - T) / 4)
        q[0] = (R[0, 2] + R[2, 0]) / (4 * q[2]); q[1] = (R[1, 2] + R[2, 1]) / (4 * q[2]); q[3] = (R[0, 1] - R[1, 0]) / (4 * q[2])
    else:
        q[3] = math.sqrt((1 + T) / 4)
        q[0] = (R[1, 2] - R[2, 1]) / (4 * q[3]); q[1] = (R[2, 0] - R[0, 2]) / (4 * q[3]); q[2] = (R[0, 1] - R[1, 0]) / (4 * q[3])
    q = q / np.linalg.norm(q)
    return -q if q[3] < 0 else q


def _dir(phi, psi):
    e = np.array([math.cos(phi) * math.sin(psi), math.sin(phi), math.cos(phi) * math.cos(psi)])
    J = np.array([[-math.sin(phi) * math.sin(psi), math.cos(phi) * math.cos(psi)],
                  [math.cos(phi), 0.0],
                  [-math.sin(phi) * math.cos(psi), -math.cos(phi) * math.sin(psi)]])
    return e, J


def _hproj(h):
    return np.array([[1 / h[2], 0, -h[0] / h[2] ** 2], [0, 1 / h[2], -h[1] / h[2] ** 2]])


def feature_block(cfg_sigma, Ric, tic, x, N, ftype, meas):
    """Returns (status, pfinv, Hx (2L' x 6N), Hf (2L' x 3), r) for one feature (Updater.cc:109-368)."""
    Rci = Ric.T
    tci = -Rci @ tic
    L = len(meas)
    ph = L - 1
    rel = x[len(x) - 7 * ph:] if ftype == ord('1') else x[26:26 + 7 * ph]
    relI = np.zeros((ph, 7))
    relI[0, :4] = rel[0:4]
    relI[0, 4:] = -quat_to_rot(rel[0:4]) @ rel[4:7]
    for i in range(1, ph):
        relI[i, :4] = quat_mul(rel[7 * i:7 * i + 4], relI[i - 1, :4])
        relI[i, 4:] = quat_to_rot(rel[7 * i:7 * i + 4]) @ (relI[i - 1, 4:] - rel[7 * i + 4:7 * i + 7])
    RI = [quat_to_rot(relI[i, :4]) for i in range(ph)]
    RC, tC = [], []
    for i in range(ph):
        qC = rot_to_quat(Rci @ RI[i] @ Ric)
        RC.append(quat_to_rot(qC))
        tC.append(Rci @ RI[i] @ tic + Rci @ relI[i, 4:] + tci)
    f32 = np.float32
    m0 = meas[0]
    phi = math.atan2(float(m0[1]), math.sqrt(float(m0[0]) ** 2 + 1))
    psi = math.atan2(float(m0[0]), 1)
    rho = 0.0
    if abs(phi) > .5 * 3.14 or abs(psi) > .5 * 3.14:
        return 1, (phi, psi, rho), None, None, None
    e, Jang = _dir(phi, psi)
    rinv = 1. / cfg_sigma ** 2
    lam, last = 0.01, math.inf
    for _ in range(10):
        A = np.zeros((3, 3)); g = np.zeros(3); cost = 0.0
        for i in range(L):
            if i == 0:
                h = e
                Hm = np.hstack([_hproj(h) @ Jang, np.zeros((2, 1))])
            else:
                h = RC[i - 1] @ e + rho * tC[i - 1]
                Hp = _hproj(h)
                Hm = np.hstack([Hp @ RC[i - 1] @ Jang, (Hp @ tC[i - 1]).reshape(2, 1)])
            pt = np.array([f32(h[0] / h[2]), f32(h[1] / h[2])], f32)
            err = (meas[i].astype(f32) - pt).astype(np.float64)
            cost += rinv * err @ err
            A += rinv * Hm.T @ Hm
            g += rinv * Hm.T @ err
        if cost <= last:
            A[np.diag_indices(3)] += lam * np.diag(A)
            dp = np.linalg.solve(A, g)
            phi += dp[0]; psi += dp[1]; rho += dp[2]
            e, Jang = _dir(phi, psi)
            if abs(last - cost) < 1e-6 and dp[2] < 1e-6:
                break
            lam *= .1
            last = cost
        else:
            lam *= 10
            last = cost
    if abs(phi) > .5 * 3.14 or abs(psi) > .5 * 3.14 or math.isinf(rho) or rho < 0:
        return 2, (phi, psi, rho), None, None, None
    ph_full = ph
    if ftype == ord('2'):
        L = int(math.ceil(.5 * L)); ph = L - 1
    Hx = np.zeros((2 * L, 6 * N)); Hf = np.zeros((2 * L, 3)); r = np.zeros(2 * L)
    c0 = 6 * (N - ph_full) if ftype == ord('1') else 0
    pt = np.array([f32(e[0] / e[2]), f32(e[1] / e[2])], f32)
    r[0:2] = (meas[0].astype(f32) - pt).astype(np.float64)
    Hf[0:2, 0:2] = _hproj(e) @ Jang
    for i in range(1, L):
        R = RI[i - 1]; Rc = RC[i - 1]; tc = tC[i - 1]
        h = Rc @ e + rho * tc
        pt = np.array([f32(h[0] / h[2]), f32(h[1] / h[2])], f32)
        Hp = _hproj(h)
        r[2 * i:2 * i + 2] = (meas[i].astype(f32) - pt).astype(np.float64)
        for j in range(i):
            RjT = RI[j].T
            dpx = skew(Ric @ e + rho * tic + rho * RjT @ relI[j, 4:])
            right = -rho * (np.eye(3) if j == 0 else RI[j - 1].T)
            sub = np.hstack([dpx @ RjT, right])
            Hx[2 * i:2 * i + 2, c0 + 6 * j:c0 + 6 * j + 6] = Hp @ Rci @ R @ sub
        Hf[2 * i:2 * i + 2, :] = np.hstack([Hp @ Rc @ Jang, (Hp @ tc).reshape(2, 1)])
    return 0, (phi, psi, rho), Hx, Hf, r


def update(sigma, T_BC0, x, P, types, offsets, xy):
    """Returns (x_out, P_out, info dict).  P: (d,d)."""
    T = np.array(T_BC0, np.float64).reshape(4, 4)
    Ric, tic = T[:3, :3], T[:3, 3]
    N = (len(x) - 26) // 7
    n = 6 * N
    d = 24 + n
    Pcc = P[24:, 24:]
    G = np.zeros((n, n)); z = np.zeros(n)
    status, gammas, pf = [], [], []
    n_good = 0; rows = 0
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    for f in range(len(types)):
        meas = xy[offsets[f]:offsets[f + 1]]
        st, pfinv, Hx, Hf, r = feature_block(sigma, Ric, tic, x, N, int(types[f]), meas)
        pf.append(pfinv)
        if st != 0:
            status.append(st); gammas.append(np.nan)
            continue
        ncol = 3 if np.linalg.norm(Hf[:, 2]) >= 1e-4 else 2
        Q, _ = np.linalg.qr(Hf[:, :ncol], mode="complete")
        Nl = Q[:, ncol:]                                    # left null space
        Hn = Nl.T @ Hx; rn = Nl.T @ r
        dof = Hn.shape[0]
        S = Hn @ Pcc @ Hn.T + sigma ** 2 * np.eye(dof)
        S = .5 * (S + S.T)
        gamma = abs(rn @ np.linalg.solve(S, rn))
        gammas.append(gamma)
        if gamma < round(float(_chi2.ppf(0.95, dof)), 6):
            G += Hn.T @ Hn; z += Hn.T @ rn
            n_good += 1; rows += dof
            status.append(0)
        else:
            status.append(3)
    info = dict(n_good=n_good, rows=rows, status=np.array(status), gamma=np.array(gammas), pfinv=np.array(pf), G=G, z=z)
    if n_good <= 2:
        return x.copy(), P.copy(), info
    M = G @ Pcc + sigma ** 2 * np.eye(n)
    Pc_ = P[:, 24:]                                         # d x n
    dx = Pc_ @ np.linalg.solve(M, z)
    Pn = P - Pc_ @ np.linalg.solve(M, G @ P[24:, :])
    Pn = .5 * (Pn + Pn.T)
    xo = x.copy()

    def dq(v):
        q = np.zeros(4); q[:3] = .5 * v
        nrm = np.linalg.norm(q[:3])
        if nrm < 1:
            q[3] = math.sqrt(1 - nrm ** 2)
        else:
            q[:3] *= 1 / math.sqrt(1 + nrm ** 2); q[3] = 1 / math.sqrt(1 + nrm ** 2)
        return q
    xo[0:4] = quat_mul(dq(dx[0:3]), x[0:4])
    xo[4:10] = dx[3:9] + x[4:10]
    xo[7:10] /= np.linalg.norm(xo[7:10])
    xo[10:14] = quat_mul(dq(dx[9:12]), x[10:14])
    xo[14:26] = dx[12:24] + x[14:26]
    for c in range(N):
        xo[26 + 7 * c:30 + 7 * c] = quat_mul(dq(dx[24 + 6 * c:27 + 6 * c]), x[26 + 7 * c:30 + 7 * c])
        xo[30 + 7 * c:33 + 7 * c] = dx[27 + 6 * c:30 + 6 * c] + x[30 + 7 * c:33 + 7 * c]
    return xo, Pn, info
