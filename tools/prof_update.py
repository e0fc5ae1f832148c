#!/usr/bin/env python
"""Small driver for ncu: runs Updater::update on a worst-case shaped case a few times.
    ncu --set full --import-source on -k regex:'k_chol_S|k_rank_rule|k_solve_small_R' -c 6 -o gpurun_out/prof python tools/prof_update.py 1 2
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rvio_b200  # noqa: E402,F401
from rvio_b200 import synth, host, capi  # noqa: E402

for idx in [int(a) for a in sys.argv[1:]] or [1]:
    cfg = synth.baseline_config(idx)
    nf = min((cfg.n_features + 1) // 2, 256)
    x, P, types, off, xy = synth.make_update_case(cfg, nf, 900 + idx, mix_types=False)
    upd = host.Updater(cfg)
    for _ in range(3):
        upd.update(x, P, types, (off, xy))
    print(idx, upd.info.n_good, upd.info.rows_stacked, upd.info.rank, upd.info.rank_flags)
    L = capi.lib()
    if hasattr(L, "rvio_b200_phase_clocks"):          # profiling variant (make PHASES=1; RVIO_B200_LIB=.../librvio_b200_phases.so)
        import ctypes
        buf = (ctypes.c_longlong * 64)()
        L.rvio_b200_phase_clocks(buf, 64)
        c = list(buf)
        def seg(name, ks):
            v = [c[k] for k in ks]
            print(f"  {name}: " + " ".join(f"{b - a}" for a, b in zip(v, v[1:])) + f"  (cycles; total {v[-1] - v[0]})")
        seg("k_rank_rule   [setup | fill | cholesky | test | verdict | emit]", [16, 17, 18, 19, 20, 21])
        if c[6] > c[0] > 0:
            seg("k_solve_small_R [stage | W | S | cholesky | P+ | correction]", [0, 1, 2, 3, 4, 5, 6])
        seg("  panel 1 of the rank rule [A | barrier | walk | diag tile | factor | my rows | barrier]", [48, 49, 50, 51, 52, 53, 54, 55])
        seg("  panel 1 of the other Cholesky [A | barrier | - | diag tile | factor | my rows | barrier]", [40, 41, 42, 43, 44, 45, 46, 47])
        seg("    diagonal tile 2, rank rule [columns | norms + pivots | X]", [56, 57, 58, 59])
        seg("    diagonal tile 2, other     [columns | norms + pivots | X]", [60, 61, 62, 63])
        if c[34] > c[32] > 0:
            seg("k_chol_S      [cholesky | inverses + write-out]", [32, 33, 34])
    upd.close()
