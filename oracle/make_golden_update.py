"""Generates tests/golden/update_golden.npz: frozen Updater::update cases (inputs -> x+, P+, nRank, per-feature status /
Mahalanobis distance) at the shapes of BASELINE.json configs 1, 2, 3 and 5.

The reference ships no vectors for this stage and cannot be built here (no Eigen): the outputs come from oracle/updater.c
in the REFERENCE rule and are written only if the independent restatement (oracle/np_compress.py: wavefront-ordered
Givens sweep, SciPy QR, numpy gain) reproduces them (rank exactly, x+/P+ to 1e-10).  The GPU box has no /root/reference
and needs none: tests/test_gpu_golden.py compares the device path with these arrays.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:] = [p for p in sys.path if os.path.abspath(p or '.') != os.path.dirname(os.path.abspath(__file__))]
sys.path.insert(0, ROOT)
import rvio_b200  # noqa: E402,F401
from rvio_b200 import synth  # noqa: E402
from oracle import oracle as orc, np_compress as npc  # noqa: E402


def stream_cases(cfg, n_frames, seed, want):
    st = synth.Stream(cfg, n_frames, seed, t_static=0.5)
    v = orc.VioOracle(cfg)
    consumed, out = 0, {}
    for i in range(st.n_frames):
        imu, consumed = st.imu_for_frame(i, consumed)
        v.step(st.frames[i], imu)
        if i in want and v.last_info is not None and v.last_info.updated:
            out[i] = tuple(a.copy() for a in v.last_update_in)
        v.last_info = None
    return out


def freeze(out, name, cfg_idx, cfg, x, Pc, types, off, xy):
    d = int(round(np.sqrt(len(Pc)))); P = Pc.reshape(d, d).T.copy()
    xo, Po, info, dbg = orc.updater_update(cfg, x, P, types, off, xy, debug=True)
    sigma = float(max(np.float32(cfg.sigma_px), np.float32(cfg.sigma_py)))
    Hn, rn, ci = npc.compress_reference(dbg["H"], dbg["r"])
    assert ci["rank"] == info.rank and ci["rank_full"] == info.rank_full, (name, ci["rank"], info.rank)
    xn, Pn = npc.ekf_reference_form(x, P, Hn, rn, sigma)
    assert np.abs(xo - xn).max() < 1e-10 and np.abs(Po - Pn).max() < 1e-10 * np.abs(Po).max(), name
    for k, v in dict(cfg=np.array(cfg_idx), x=x, P=P, types=types, off=off, xy=xy, x_out=xo, P_out=Po,
                     rank=np.array(info.rank), rank_full=np.array(info.rank_full), n_good=np.array(info.n_good),
                     rows=np.array(info.rows_stacked), compressed=np.array(info.compressed),
                     status=dbg["status"], gamma=dbg["gamma"], pfinv=dbg["pfinv"]).items():
        out[f"{name}/{k}"] = np.asarray(v)
    print(f"{name}: N={(len(x) - 26) // 7} feats={len(types)} good={info.n_good} rows={info.rows_stacked} "
          f"rank={info.rank}/{info.rank_full} |dx|max={np.abs(xo - x).max():.3e}")


def main():
    out = {}
    names = []
    c1 = synth.baseline_config(0)
    for i, c in stream_cases(c1, 46, 20260922, {36, 41, 44}).items():
        freeze(out, f"cfg1_f{i}", 0, c1, *c); names.append(f"cfg1_f{i}")
    c2 = synth.baseline_config(1)
    # 43: plain compressed frame; 47: the cut discards the '1' rows (rank 29 of 38); 54: dependent column in the middle,
    # nothing discarded (decided by the sweep); 52: fat-ish mixed frame
    for i, c in stream_cases(c2, 60, 20260923, {35, 43, 47, 52, 54}).items():
        freeze(out, f"cfg2_f{i}", 1, c2, *c); names.append(f"cfg2_f{i}")
    for idx, nf, seed in ((2, 48, 302), (4, 40, 304)):
        cfg = synth.baseline_config(idx)
        x, P, types, off, xy = synth.make_update_case(cfg, nf, seed)
        freeze(out, f"cfg{idx + 1}_worst", idx, cfg, x, np.ascontiguousarray(P.T).reshape(-1), types, off, xy)
        names.append(f"cfg{idx + 1}_worst")
    xa, Pa, ta, oa, xya = synth.make_update_case(c2, 24, 5, track_len=4, mix_types=True)
    freeze(out, "cfg2_disjoint", 1, c2, xa, np.ascontiguousarray(Pa.T).reshape(-1), ta, oa, xya); names.append("cfg2_disjoint")
    out["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "update_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
