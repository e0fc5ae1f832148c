#!/usr/bin/env python
"""Writes a short synthetic stream as an EuRoC ASL directory (tools/run_asl.py input) -- used to smoke-test the I/O path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rvio_b200  # noqa: E402,F401
from rvio_b200 import synth, io_formats  # noqa: E402

out = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cfg = synth.baseline_config(1)
st = synth.Stream(cfg, n, 77, t_static=1.0)
io_formats.write_asl(out, st.frame_t, st.frames, st.imu)
print("wrote", out, n, "frames")
