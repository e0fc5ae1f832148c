#!/usr/bin/env python
"""Profiling driver for the corner selection kernel (profiling build: make PHASES=1, RVIO_B200_LIB=.../librvio_b200_phases.so):
runs a short synthetic stream through the fused pipeline with the device detector and prints k_det_select's phase clocks."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rvio_b200  # noqa: E402,F401
from rvio_b200 import synth, host, capi  # noqa: E402

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = synth.baseline_config(idx)
n = 40
st = synth.Stream(cfg, n, 11, t_static=0.5)
v = host.Vio(cfg, 0)
consumed = 0
L = capi.lib()
for i in range(n):
    imu, consumed = st.imu_for_frame(i, consumed)
    if hasattr(L, "rvio_b200_gram_ns"):
        gbuf = (ctypes.c_ulonglong * 16)()
        L.rvio_b200_gram_ns(gbuf, 1)
    v.step(st.frames[i], imu, None, device_detector=True)
    if hasattr(L, "rvio_b200_gram_ns") and i >= n - 3:
        gbuf = (ctypes.c_ulonglong * 16)()
        L.rvio_b200_gram_ns(gbuf, 0)
        gg = list(gbuf)
        if gg[4] > gg[0] > 0:
            print(f"frame {i}: k_gram ns from the first CTA's start: products done {gg[1] - gg[0]}, partials + fence {gg[2] - gg[0]}, tile-0 reduce done {gg[3] - gg[0]}, counters + classes done {gg[4] - gg[0]}")
    if hasattr(L, "rvio_b200_det_clocks") and i >= n - 3:
        buf = (ctypes.c_longlong * 64)()
        L.rvio_b200_det_clocks(buf, 64)
        c = list(buf)
        nb = int(c[31])
        print(f"frame {i}: {c[30]} candidates, {nb} block(s), {c[32]} settle iterations, {c[33]} corners; total {c[29] - c[0]} cycles")
        for b in range(min(nb, 6)):
            s = [c[k + 4 * b] for k in (1, 2, 3, 4)]
            prev = c[0] if b == 0 else c[4 * b]
            if b == 0 and c[44] > c[40] > 0:
                print(f"   first settle iteration: masks {c[41] - c[40]}, rounds {c[42] - c[41]}, emit {c[43] - c[42]}, filter + compact {c[44] - c[43]} -> {c[45]} alive")
            print(f"   block {b}: {c[34 + b]} keys: filter+histogram {s[0] - prev}, gather {s[1] - s[0]}, sort {s[2] - s[1]}, settle {s[3] - s[2]}")
if hasattr(L, "rvio_b200_aug_clocks"):
    buf = (ctypes.c_longlong * 16)()
    L.rvio_b200_aug_clocks(buf)
    c = list(buf)
    names = ["augmentation copy", "state shift", "V + P00 (thread 0's pose algebra)", "24x24 products", "cross terms", "final"]
    print("k_augment_compose (cycles): " + ", ".join(f"{nm} {c[k + 1] - c[k]}" for k, nm in enumerate(names)) + f"; total {c[6] - c[0]}")
if hasattr(L, "rvio_b200_feat_clocks"):
    buf = (ctypes.c_longlong * 16)()
    L.rvio_b200_feat_clocks(buf)
    c = list(buf)
    names = ["setup + pose chain", "camera poses", "LM", "Jacobians", "nullspace", "gate product", "symmetrise", "Cholesky + gamma", "publish"]
    print("k_feature, a full-length type-2 track (cycles): " + ", ".join(f"{nm} {c[k + 1] - c[k]}" for k, nm in enumerate(names)) + f"; total {c[9] - c[0]}")
if hasattr(L, "rvio_b200_trk_clocks"):
    buf = (ctypes.c_longlong * 32)()
    L.rvio_b200_trk_clocks(buf, 32)
    c = list(buf)
    names = ["compaction", "(barrier)", "draws", "models", "count inliers", "winner + flags", "bookkeeping: lost", "tracked", "scalars"]
    print("k_ransac_bookkeep (cycles): " + ", ".join(f"{nm} {c[k + 1] - c[k]}" for k, nm in enumerate(names[:8])) + f"; total {c[8] - c[0]}")
v.close()
