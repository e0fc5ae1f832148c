"""Independent NumPy/SciPy restatement of the reference's model compression + EKF step (src/rvio/Updater.cc:460-619).

TEST INFRASTRUCTURE ONLY -- a second opinion on the part of oracle/updater.c that decides parity with the reference:
the Givens sweep order (Updater.cc:494-512), the trailing-zero-column drop (:480-489), the first-small-row cut
(:515-524) and the gain / Joseph form (:540-619).  Eigen is absent from this image, so neither can be pinned against
the reference binary; the two restatements share no code and use different schedules:

  * updater.c            column by column, rows bottom-up (the reference's loop nest), hand-written LU inverse
  * compress_reference() the same rotations in WAVEFRONT order (rotation (n, m) at step (M-1-m) + 2n; rotations of one
                         step act on disjoint row pairs) -- any topological order of the rotation DAG must give the
                         same trapezoid; numpy.linalg for S^-1
  * unique_rows_by_qr()  scipy.linalg.qr (Householder, no pivoting): the rows of R before the first dependent column are
                         unique up to sign for ANY orthogonal triangularisation, so they must match the Givens trapezoid

Eigen semantics used (public documentation): JacobiRotation::makeGivens(p, q) -> q == 0: c = sign(p),
s = 0; p == 0: c = 0, s = -sign(q); otherwise c = p / hypot, s = -q / hypot (both of Eigen's branches reduce to that);
applyOnTheLeft(0, 1, G.adjoint()): row0' = c row0 - s row1, row1' = s row0 + c row1.
"""
from __future__ import annotations

import numpy as np


def make_givens(p: float, q: float):
    if q == 0:
        return (-1.0 if p < 0 else 1.0), 0.0
    if p == 0:
        return 0.0, (1.0 if q < 0 else -1.0)
    h = float(np.hypot(p, q))
    return p / h, -q / h


def compress_reference(H: np.ndarray, r: np.ndarray):
    """H: R x n stacked clone-column Jacobian, r: R.  Returns (Hn, rn, info) as Updater.cc:474-536 leaves them."""
    M, n = H.shape
    if M <= n:                                            # fat: no compression (:531-536)
        return H.copy(), r.copy(), dict(compressed=False, Np=n, rank=M, rank_full=M)
    Np = n
    while Np > 0 and np.linalg.norm(H[:, Np - 1]) == 0:   # :480-489
        Np -= 1
    A = np.concatenate([H[:, :Np], r[:, None]], axis=1).astype(np.float64, copy=True)
    T = M + Np - 2
    for t in range(T):                                    # wavefront over the rotation DAG
        n_lo, n_hi = max(0, t - M + 2), min(Np - 1, t // 2)
        for c0 in range(n_lo, n_hi + 1):
            m = M - 1 - t + 2 * c0
            c, s = make_givens(A[m - 1, c0], A[m, c0])
            if c == 1.0 and s == 0.0:
                continue
            x = A[m - 1, c0:].copy(); y = A[m, c0:].copy()
            A[m - 1, c0:] = c * x - s * y
            A[m, c0:] = s * x + c * y
    norms = np.linalg.norm(A[:, :Np], axis=1)
    rank = 0
    while rank < M and norms[rank] >= 1e-4:               # :515-524
        rank += 1
    Hn = np.zeros((rank, n)); Hn[:, :Np] = A[:rank, :Np]
    return Hn, A[:rank, Np].copy(), dict(compressed=True, Np=Np, rank=rank, rank_full=int((norms >= 1e-4).sum()),
                                         trapezoid=A[:, :Np], norms=norms)


def unique_rows_by_qr(H: np.ndarray, tol: float = 1e-9):
    """Householder QR of the non-zero columns; returns (|R| rows before the first dependent column, j*)."""
    import scipy.linalg as sl
    n = H.shape[1]
    Np = n
    while Np > 0 and np.linalg.norm(H[:, Np - 1]) == 0:
        Np -= 1
    R = sl.qr(H[:, :Np], mode="r")[0][:Np]
    dg = np.abs(np.diag(R))
    dep = np.nonzero(dg < tol * max(1.0, dg.max()))[0]
    j = int(dep[0]) if len(dep) else Np
    return R[:j], j


def ekf_reference_form(x, P, Hn, rn, sigma):
    """Updater.cc:540-619 on the kept rows: S = H P H^T + s^2 I (symmetrised), K = P H^T S^-1, dx = K r, Joseph form."""
    from . import np_updater as npu
    d = P.shape[0]
    N = (len(x) - 26) // 7
    Hf = np.zeros((Hn.shape[0], d)); Hf[:, 24:] = Hn
    S = Hf @ P @ Hf.T + sigma * sigma * np.eye(len(rn))
    S = .5 * (S + S.T)
    K = P @ Hf.T @ np.linalg.inv(S)
    dx = K @ rn
    xo = x.copy()

    def dq(v):
        q = np.zeros(4); q[:3] = .5 * v
        nv = np.linalg.norm(q[:3])
        if nv < 1:
            q[3] = np.sqrt(1 - nv * nv)
        else:
            q[:3] /= np.sqrt(1 + nv * nv); q[3] = 1 / np.sqrt(1 + nv * nv)
        return q
    xo[0:4] = npu.quat_mul(dq(dx[0:3]), x[0:4])
    xo[4:10] = dx[3:9] + x[4:10]
    xo[7:10] /= np.linalg.norm(xo[7:10])
    xo[10:14] = npu.quat_mul(dq(dx[9:12]), x[10:14])
    xo[14:26] = dx[12:24] + x[14:26]
    for c in range(N):
        xo[26 + 7 * c:30 + 7 * c] = npu.quat_mul(dq(dx[24 + 6 * c:27 + 6 * c]), x[26 + 7 * c:30 + 7 * c])
        xo[30 + 7 * c:33 + 7 * c] = dx[27 + 6 * c:30 + 6 * c] + x[30 + 7 * c:33 + 7 * c]
    A = np.eye(d) - K @ Hf
    Pn = A @ P @ A.T + sigma * sigma * (K @ K.T)
    return xo, .5 * (Pn + Pn.T)
