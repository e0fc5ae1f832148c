#!/usr/bin/env python
"""How much accuracy does the normal-term route (G = H^T H, Cholesky) give away against an orthogonal triangularisation of the
stacked Jacobian (SURVEY hard part 3: "do not square the condition number")?  CPU study on the oracle's own stacked H of real
update frames (the oracle is test infrastructure; this script is analysis, not product):

  for every update of a seeded stream that compresses (rows > columns)
    H, r   = stacked, gated Jacobian and residual as Updater.cc:424-458 leaves them (oracle debug output)
    cols   = columns with information (the trailing / dependent ones are dropped by a rank-revealing QR)
    route Q: Householder QR of H[:, cols]            -> R_q, y_q = Q^T r
    route G: fp64 G = H^T H, z = H^T r, Cholesky     -> R_g, y_g = R_g^-T z
    dx_*   = P_c R^T (R P_cc R^T + sigma^2 I)^-1 y   (the state correction either route feeds)
  reported: cond_2(H[:, cols]), |R_g - R_q| / |R_q|, |y_g - y_q| / |y_q|, |dx_g - dx_q| (metres / radians of the state).

    python tools/gram_vs_qr.py [config_index] [n_frames] [seed]
"""
import os
import sys

import numpy as np
import scipy.linalg as sl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvio_b200  # noqa: E402,F401
from rvio_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def study(idx, n_frames, seed):
    cfg = synth.baseline_config(idx)
    st = synth.Stream(cfg, n_frames, seed, t_static=0.5)
    v = orc.VioOracle(cfg, lambda img, n, s: orc.detect_with_subpix(img, n, s, cfg))
    sig2 = float(max(np.float32(cfg.sigma_px), np.float32(cfg.sigma_py))) ** 2
    consumed = 0
    rows = []
    for i in range(st.n_frames):
        imu, consumed = st.imu_for_frame(i, consumed)
        v.last_info = None
        v.step(st.frames[i], imu)
        li = v.last_info
        if li is None or li.n_good < 3:
            continue
        x, Pcm, types, off, xy = v.last_update_in
        d = 24 + 6 * ((len(x) - 26) // 7)
        P = Pcm.reshape(d, d).T
        _, _, info, dbg = orc.updater_update(cfg, x, P, types, off, xy, debug=True)
        H, r = dbg["H"], dbg["r"]
        n = H.shape[1]
        if H.shape[0] <= n:
            continue
        # columns that carry information: rank-revealing QR, tolerance far above rounding and far below the reference's 1e-4 row test
        _, Rp, piv = sl.qr(H, mode="economic", pivoting=True)
        k = int(np.sum(np.abs(np.diag(Rp)) > 1e-7 * abs(Rp[0, 0])))
        cols = np.sort(piv[:k])
        A = H[:, cols]
        s = np.linalg.svd(A, compute_uv=False)
        Q, Rq = np.linalg.qr(A)
        sg = np.sign(np.diag(Rq)); sg[sg == 0] = 1
        Rq = Rq * sg[:, None]; yq = (Q * sg[None, :]).T @ r
        G = A.T @ A; z = A.T @ r
        Rg = np.linalg.cholesky(G).T
        yg = sl.solve_triangular(Rg, z, trans="T", lower=False)
        Pc = P[:, 24:][:, cols]; Pcc = P[24:, 24:][np.ix_(cols, cols)]

        def dx(R, y):
            S = R @ Pcc @ R.T + sig2 * np.eye(len(y))
            return Pc @ R.T @ np.linalg.solve(S, y)
        dq, dg = dx(Rq, yq), dx(Rg, yg)
        rows.append((i, H.shape[0], n, k, s[0] / s[-1], np.linalg.norm(Rg - Rq) / np.linalg.norm(Rq),
                     np.linalg.norm(yg - yq) / max(np.linalg.norm(yq), 1e-300), float(np.abs(dg - dq).max()), float(np.abs(dq).max())))
    return rows


def main():
    idx = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 20260923
    rows = study(idx, n_frames, seed)
    print(f"configs[{idx}] stream, seed {seed}, {n_frames} frames: {len(rows)} compressing updates")
    print("frame  rows  n  kept  cond2(H)   |Rg-Rq|/|Rq|  |yg-yq|/|yq|  max|dx_g-dx_q|  max|dx|")
    for r in rows:
        print(f"{r[0]:5d} {r[1]:5d} {r[2]:3d} {r[3]:4d}  {r[4]:9.3e}  {r[5]:11.3e}  {r[6]:11.3e}  {r[7]:13.3e}  {r[8]:9.3e}")
    a = np.array([r[4:] for r in rows])
    print(f"worst: cond2 {a[:, 0].max():.3e}, |dR|/|R| {a[:, 1].max():.3e}, |dy|/|y| {a[:, 2].max():.3e}, |d dx| {a[:, 3].max():.3e} "
          f"(largest correction {a[:, 4].max():.3e})")


if __name__ == "__main__":
    main()
